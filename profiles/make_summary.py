"""Regenerate a round's ncu summary from the committed ncu exports (profiles/<tag>_*_ncu_raw.csv, <tag>_launches.csv).

    python profiles/make_summary.py            > profiles/r01_ncu_summary.md
    python profiles/make_summary.py r02g profiles/r02_preface.md > profiles/r02_ncu_summary.md
"""
import sys
import collections
import csv
import glob
import os

import summarize

HERE = os.path.dirname(os.path.abspath(__file__))
PREFACE = """# Round 1 — ncu evidence (B200, `gpurun -- 'bash profiles/capture.sh'`)

All captures: `ncu --set full --clock-control none --import-source on` on one EAGER bench step per kernel family
(`bench.py --steps 2 --warmup 1 --eager`; the CUDA graph of the timed bench replays exactly these launches).  ncu
serialises launches and flushes caches between replay passes, so absolute times here are cold-cache: compare the
kernels' SHARE of a step with `bench.py --profile-kernels` (CUDA events, warm), not the absolute durations.

How to read it (details in DESIGN.md section 6):

* `hash_bwd_kernel<2,0,1>` (main grids, with dL/dx): L2 throughput ~72 % of peak, DRAM ~157 MB per launch against
  231 MB of algorithmic bytes -> bound by L2 atomic (RED) throughput; 35 % warps active at 72 registers.
* `hash_fwd_kernel`: ~62 us per main-grid launch, L2 45 %, DRAM 54 MB.
* `prop_{fwd,bwd}_kernel`: issue-bound (41-54 % issue-active at 18-23 % occupancy).
* `mlp_tc_*`: tensor pipe 12-18 % active, 8 warps per 128-point tile, 2-3 CTAs per SM (1 for the head backward).
* `adam_kernel`: DRAM throughput 74 % of peak under the profiler; 5.97 TB/s (91 % of the measured copy peak) when
  timed with CUDA events (`tools/bench_adam.py`).
"""


def launch_table(path):
    rows = list(csv.reader(line for line in open(path) if line.startswith('"')))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    names = [r[ik] for r in rows[1:]]
    t = [float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}[r[iu]] for r in rows[1:]]
    marks = [i for i, n in enumerate(names) if "loss_sum_kernel" in n]  # launched once per step
    a, b = marks[-2], marks[-1]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in zip(names[a:b], t[a:b]):
        key = n.split("(")[0].replace("void ", "")
        if "tn::" not in n:
            key = "torch glue (aten elementwise / reduce / fill / copy kernels)"
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"## One train step, launch by launch ({b - a} launches, {tot:.0f} us serialised, cold cache)", "",
           "| kernel | launches | us | share |", "|---|---|---|---|"]
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k[:80]}` | {n} | {v:.1f} | {100 * v / tot:.1f} % |")
    return "\n".join(out)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    print(open(sys.argv[2]).read() if len(sys.argv) > 2 else PREFACE)
    print(launch_table(os.path.join(HERE, f"{tag}_launches.csv")))
    print()
    for path in sorted(glob.glob(os.path.join(HERE, f"{tag}_*_ncu_raw.csv"))):
        rows = summarize.load(path)
        cols = ["kernel"] + [s for _, s in summarize.KEEP if any(s in r for r in rows)]
        print(f"## {os.path.basename(path)}")
        print("| " + " | ".join(cols) + " |")
        print("|" + "---|" * len(cols))
        for r in rows:
            print("| " + " | ".join(str(r.get(c, "")) for c in cols) + " |")
        print()
