"""Encode kernels alone on the sample positions of the bench batch: fwd / bwd time per launch with the tables cold
(L2 flushed between launches, as inside a train step where ~300 MB pass through L2 between two uses of a table) and
hot (back-to-back launches).   python tools/bench_encode.py [--rays 4096] [--log2-T 19]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nerfstudio_thermal_b200 as tn  # noqa: E402
from nerfstudio_thermal_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--log2-T", type=int, default=19)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda")
    args = argparse.Namespace(density_mode="separate", log2_hashmap_size=a.log2_T, init="trained")
    model = bench.build_model(args).to(dev).train()
    b = {k: v.to(dev) for k, v in bench.make_batch(a.rays, 42).items()}
    rb = tn.RayBundle(origins=b["origins"], directions=b["directions"], pixel_area=b["pixel_area"],
                      camera_indices=b["camera_indices"])
    with torch.no_grad():
        model(rb)
    x = model.field._sample_locations.detach().contiguous()  # [R, S, 3]
    R, S = x.shape[:2]
    enc = model.field.mlp_base.model[0]
    table, spec = enc.hash_table.detach(), enc.spec
    flat = x.view(-1, 3)
    n = flat.shape[0]
    dy = torch.randn(n, spec.out_dim, device=dev)
    dtable = torch.zeros_like(table)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, cold):
        ts = []
        for _ in range(a.reps):
            if cold:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    out = torch.empty(n, spec.out_dim, device=dev)
    dx = torch.empty_like(flat)
    from nerfstudio_thermal_b200._lib import call, ptr, stream

    def fwd(spr):
        call("tn_hash_encode_fwd", ptr(flat), ptr(table), 0, spec._c_scales, n, spec.num_levels, spec.features,
             spec.log2_T, spr, ptr(out), None, None, stream())

    def bwd(spr, want_dx=True):
        call("tn_hash_encode_bwd", ptr(flat), ptr(table), 0, spec._c_scales, ptr(dy), n, spec.num_levels, spec.features,
             spec.log2_T, spr, ptr(dtable), ptr(dx) if want_dx else None, None, stream())

    for name, fn in (("fwd patch", lambda: fwd(S)), ("fwd plain", lambda: fwd(0)), ("bwd patch dx", lambda: bwd(S)),
                     ("bwd plain dx", lambda: bwd(0)), ("bwd patch nodx", lambda: bwd(S, False))):
        for _ in range(3):
            fn()
        print(f"{name:16s} N={n} T=2^{a.log2_T}: cold {timeit(fn, True):7.1f} us   hot {timeit(fn, False):7.1f} us")


if __name__ == "__main__":
    main()
