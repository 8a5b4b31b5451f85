"""L2 gather / RED throughput of this GPU with the hash-grid kernels' access pattern (tn_l2_probe).
usage: python tools/probe_l2.py [log2_rows ...]   -> one JSON line per table size"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfstudio_thermal_b200 import _lib  # noqa: E402
from nerfstudio_thermal_b200._lib import call, ptr, stream  # noqa: E402

MODES = {0: "gather8", 1: "gather8_pair", 6: "gather8_quad", 2: "red_v2", 3: "red_v2_pair", 4: "red_v4", 5: "red_f32",
         7: "gather8_pair_sector", 8: "gather8_pair_line", 9: "red_v2_pair_sector", 10: "red_v2_pair_line"}


def probe(log2_rows=23, iters=64, ctas=148 * 8, reps=5):
    """{name: giga-operations per second (per-lane ops)} on a table of 2^log2_rows 8-byte rows."""
    dev = torch.device("cuda")
    table = torch.zeros((1 << log2_rows, 2), device=dev)
    sink = torch.zeros(1, device=dev)
    out = {}
    for mode, name in MODES.items():
        best = float("inf")
        for _ in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call("tn_l2_probe", mode, ptr(table), log2_rows, iters, ctas, ptr(sink), stream())
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name] = ctas * 256 * iters / (best * 1e-3) / 1e9
    return out


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [23, 25]
    for lr in sizes:
        res = probe(lr)
        print(json.dumps({"log2_rows": lr, "table_MB": (8 << lr) / 1e6, "Gops_per_s": {k: round(v, 2) for k, v in res.items()}}))
