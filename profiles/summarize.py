"""Condense `ncu --page raw --csv` dumps (gpurun_out/*.raw.csv) into profiles/<name>_summary.csv + a markdown table."""
import csv
import sys

KEEP = [
    ("gpu__time_duration.sum", "time_us"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_bytes.sum", "l2_MB"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("lts__t_sectors_op_red.sum", "l2_red_sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1_ld_sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "l1_red_sectors"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wavefronts_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_hmma_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
]


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        d = {"kernel": r[idx["Kernel Name"]].split("(")[0][:48]}
        for m, short in KEEP:
            if m not in idx:
                continue
            v, u = to_float(r[idx[m]]), units[idx[m]]
            if v is None:
                continue
            if short.endswith("_MB"):
                v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
            if short == "time_us":
                v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
            d[short] = round(v, 3)
        out.append(d)
    return out


if __name__ == "__main__":
    for path in sys.argv[1:]:
        rows = load(path)
        cols = ["kernel"] + [s for _, s in KEEP if any(s in r for r in rows)]
        print(f"### {path}")
        print("| " + " | ".join(cols) + " |")
        print("|" + "---|" * len(cols))
        for r in rows:
            print("| " + " | ".join(str(r.get(c, "")) for c in cols) + " |")
        print()
