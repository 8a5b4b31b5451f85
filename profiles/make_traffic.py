"""profiles/traffic.json from the round's `ncu --set full` raw pages: per launch tag of bench.py, the DRAM bytes of
one launch (`roofline.traffic`) and the L1/L2 request counts behind the L2 rooflines.
usage: python profiles/make_traffic.py gpurun_out/r02_*.raw.csv > profiles/traffic.json"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3}
M = {
    "dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum", "time": "gpu__time_duration.sum",
    "l1_ld_sectors": "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1_red_sectors": "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l2_read_sectors_from_l1": "lts__t_sectors_srcunit_tex_op_read.sum",
    "l2_red_sectors": "lts__t_sectors_op_red.sum", "insts": "smsp__inst_executed.sum",
}


def tag_of(name, row):
    if "hash_bwd_kernel<2, 0, 1" in name or "hash_bwd_kernel<(int)2, (bool)0, (bool)1" in name:
        return "tn_hash_encode_bwd[L16,T2^19,dx]"
    if "hash_fwd_kernel<2, 0, 0" in name or "hash_fwd_kernel<(int)2, (bool)0, (bool)0" in name:
        jac = "hash_fwd_kernel<2, 0, 0, 1>" in name or "(bool)0, (bool)0, (bool)1>" in name
        return "tn_hash_encode_fwd[L16,T2^19,jac]" if jac else "tn_hash_encode_fwd[L16,T2^19]"
    if "prop_bwd_kernel" in name:
        return "tn_prop_density_bwd[L5,S256,dx]" if row["insts"] > 8e7 else "tn_prop_density_bwd[L5,S96,dx]"
    if "prop_fwd_kernel" in name:
        return "tn_prop_density_fwd[L5,S256]" if row["insts"] > 2e7 else "tn_prop_density_fwd[L5,S96]"
    return None


def main(paths):
    acc = {}
    for path in paths:
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            d = {}
            for k, m in M.items():
                if m not in idx:
                    d[k] = 0.0
                    continue
                try:
                    v = float(r[idx[m]].replace(",", ""))
                except ValueError:
                    v = 0.0
                u = units[idx[m]]
                d[k] = v * (UNIT.get(u, 1.0) if k.startswith("dram") else TIME.get(u, 1.0) if k == "time" else 1.0)
            tag = tag_of(r[idx["Kernel Name"]], d)
            if tag:
                acc.setdefault(tag, []).append(d)
    out = {"_source": "ncu --set full --clock-control none, one eager train step of the round's final code (profiles/capture_r02.sh, r02g_*); "
                      "per launch, averaged over the captured launches of the tag", "_detail": {}}
    for tag, ds in sorted(acc.items()):
        n = len(ds)
        avg = {k: sum(d[k] for d in ds) / n for k in M}
        out[tag] = avg["dram_rd"] + avg["dram_wr"]
        out["_detail"][tag] = {"launches_captured": n, "time_us_under_ncu": round(avg["time"], 2),
                               **{k: round(avg[k]) for k in M if k not in ("time",)}}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1:])
