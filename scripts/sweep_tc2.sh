#!/bin/bash
# tuning aid: L2 row prefetch of the CTA-pair a_max kernel off / on (MRG_TC2_PREFETCH)
for pf in 0 1 0 1; do
  echo "MRG_TC2_PREFETCH=$pf: $(MRG_TC2_PREFETCH=$pf timeout 60 python scripts/time_amax.py child 2>&1 | grep 'fp32')"
done
