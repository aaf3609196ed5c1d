"""Dump the judged metrics of an `ncu --set full` report (.ncu-rep) as one markdown table row per launch.

usage: python profiles/extract_ncu.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/rNN_x.md
Also writes, next to the markdown (when -j FILE is given), a JSON map kernel-name -> per-launch DRAM traffic that
bench.py reads for `roofline.traffic`.
"""
import csv
import io
import json
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("lts__t_sector_hit_rate.pct", "L2hit%"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main(argv):
    jpath = None
    if "-j" in argv:
        i = argv.index("-j")
        jpath = argv[i + 1]
        argv = argv[:i] + argv[i + 2:]
    traffic = {}
    for rep in argv:
        hdr, units, rows = rows_of(rep)
        print(f"### {rep.split('/')[-1]}  (ncu --set full --clock-control none; per launch, cold cache)\n")
        cols = [(hdr.index(k), lab) for k, lab in WANT if k in hdr]
        print("| kernel | " + " | ".join(lab for _, lab in cols) + " | DRAM GB/s |")
        print("|---|" + "---|" * (len(cols) + 1))
        kn = hdr.index("Kernel Name")
        for r in rows:
            name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[kn]))
            cells, t_us, byts = [], None, 0.0
            for i, lab in cols:
                v, u = r[i], units[i]
                if lab == "time":
                    t_us = to_us(v, u)
                    cells.append(f"{t_us:.1f} us")
                elif lab in ("dram_rd", "dram_wr"):
                    b = to_bytes(v, u)
                    byts += b
                    cells.append(f"{b / 1e6:.1f} MB")
                elif lab in ("grid", "block", "regs", "tensor_inst"):
                    cells.append(v.replace(",", "").split(".")[0])
                else:
                    cells.append(f"{float(v.replace(',', '')):.1f}" if v else "")
            gbs = byts / (t_us * 1e-6) / 1e9 if t_us else 0
            print(f"| `{name[:60]}` | " + " | ".join(cells) + f" | {gbs:.0f} |")
            traffic.setdefault(name, []).append(byts)
        print()
    if jpath:
        json.dump({k: {"launches": len(v), "dram_bytes_per_launch": max(v)} for k, v in traffic.items()},
                  open(jpath, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1:])
