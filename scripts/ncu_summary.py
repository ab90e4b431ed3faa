"""Summarise an `ncu --page raw --csv` export: one block per kernel launch with the metrics the roofline needs
(duration, DRAM / L2 traffic, tensor-pipe and issue utilisation, occupancy limits, warp-stall mix).

    python scripts/ncu_summary.py gpurun_out/x_raw.csv [--md] > profiles/rNN_x.txt
"""
import csv
import sys

KEYS = [
    ("time_us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("regs", "launch__registers_per_thread"),
    ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic"),
    ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_MB", "lts__t_bytes.sum"), ("l2_pct", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed"),
    ("l1_pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor_pct2", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
    ("fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("inst_M", "smsp__inst_executed.sum"),
    ("sm_busy_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("smem_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
]
STALL = "smsp__average_warps_issue_stalled_"


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        print(f"== {name[:110]}")
        line = []
        for label, key in KEYS:
            if key in col:
                v, u = num(r[col[key]]), units[col[key]]
                if v is None:
                    continue
                if u == "byte":
                    v, u = v / 1e6, "MB"
                if u == "Kbyte":
                    v, u = v / 1e3, "MB"
                if u == "Gbyte":
                    v, u = v * 1e3, "MB"
                if u == "Mbyte":
                    u = "MB"
                if label == "inst_M":
                    v /= 1e6
                if u == "ms":
                    v, u = v * 1e3, "us"
                if u == "ns":
                    v, u = v / 1e3, "us"
                line.append(f"{label}={v:.4g}{'' if label.endswith(('pct', 'MB', 'us', '_M')) else ''}")
        print("   " + "  ".join(line))
        st = []
        for h, i in col.items():
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                v = num(r[i])
                if v is not None and v >= 0.05:
                    st.append((v, h[len(STALL):-len("_per_issue_active.ratio")]))
        st.sort(reverse=True)
        print("   stalls(warps per issue): " + "  ".join(f"{n}={v:.2f}" for v, n in st[:8]))


if __name__ == "__main__":
    main()
