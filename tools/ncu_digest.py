#!/usr/bin/env python
"""Condense an .ncu-rep into the handful of numbers the roofline discussion needs (one block per captured launch).

    python tools/ncu_digest.py gpurun_out/prof_x.ncu-rep [algorithmic_bytes_or_flops] > profiles/rNN_x.txt
"""
import csv, io, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg.per_second", "sm clock"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe active %"),
    ("sm__inst_executed_pipe_tensor", "tensor inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data pipe, LSU %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "L1 data pipe, tensor-core operand reads %"),
    ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor-core unit busy %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "alu pipe active %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem)"),
]
STALLS = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:110])
        for key, label in WANT:
            for i, h in enumerate(hdr):
                if h == key or (h.startswith(key) and h[len(key):] in ("", ".sum", ".avg")):
                    print(f"  {label:28s} {r[i]} {units[i]}   [{h}]")
                    break
        st = sorted(((float(r[i] or 0), h[len(STALLS):].replace("_per_issue_active.ratio", "")) for i, h in enumerate(hdr)
                     if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio")), reverse=True)[:6]
        print("  top stalls (warps per issue):", ", ".join(f"{n} {v:.2f}" for v, n in st))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
