"""Condenses an `ncu --page raw --csv` export into a per-kernel markdown table for profiles/."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue %"),
    ("sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "ADU %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp insts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_sectors_srcunit_tex_op_write.sum", "L2 write sectors"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
]


def main(path, title):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    print("Source: `ncu --set full --clock-control none --import-source on` on one step of "
          "`python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (workload c2, 4x19x512x1024, D_conf), "
          "exported with `ncu -i ... --page raw --csv`. Per-launch values; cold-cache and serialised, so "
          "compare shares, not absolutes.\n")
    names = []
    for d in data:
        n = d[idx["Kernel Name"]]
        n = n.split("(")[0].replace("void ", "")
        names.append(n)
    print("| metric | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for key, label in KEYS:
        if key not in idx:
            continue
        vals = []
        for d in data:
            v = d[idx[key]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3g}" if abs(f) < 1e6 else f"{f:.4g}"
            except ValueError:
                pass
            vals.append(v)
        print(f"| {label} ({units[idx[key]]}) | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
