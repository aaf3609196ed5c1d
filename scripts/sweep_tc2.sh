#!/bin/bash
# tuning aid: L2 row prefetch of the CTA-pair a_max kernel on / off
for pf in 0 1 0 1; do
  echo "prefetch=$pf: $(MRG_TC2_PREFETCH=$pf timeout 60 python scripts/time_amax.py child 2>&1 | grep 'fp32')"
done
