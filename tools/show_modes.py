"""print ms/step of gpurun_out/modes_n<N>_<mode>.json (tools/bench_modes.sh)"""
import glob, json, sys
for f in sorted(glob.glob("gpurun_out/modes_n*_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][0])
        print(f"{f:44s} {d['ms_per_step']:.3f} ms/step  value {d['value']:.0f}  e2e {d['e2e']['ms_per_step']:.3f} ms")
    except Exception as e:  # noqa: BLE001
        print(f, "FAILED", e)
