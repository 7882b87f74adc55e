#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per profiled launch the duration,
DRAM bytes, throughput percentages, occupancy, registers and the top stall reasons."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64cyc%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf"),
    # shared-memory pipe: wavefronts (128 bytes each at full width) and their share of the pipe's peak
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_pipe%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed", "smem_ld%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed", "smem_st%"),
    ("lts__t_sectors.sum", "l2_sectors"),
    ("lts__t_sector_hit_rate.pct", "l2_hit%"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit%"),
    ("smsp__inst_executed.sum", "inst"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    if not stall_cols:
        stall_cols = [h for h in hdr if "issue_stalled" in h and h.endswith(".ratio")]
    for r in data:
        name = r[idx["Kernel Name"]]
        print("==", name[:100], "grid", r[idx["Grid Size"]], "block", r[idx["Block Size"]])
        parts = []
        for k, short in KEYS:
            if k in idx and r[idx[k]] != "":
                parts.append(f"{short}={r[idx[k]]}{units[idx[k]] if short in ('dur','dram_rd','dram_wr') else ''}")
        print("   ", "  ".join(parts))
        try:        # derived: shared-memory and L2 throughput in bytes per second
            dur_ns = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
            if units[idx["gpu__time_duration.sum"]].strip() in ("us", "usecond"):
                dur_ns *= 1e3
            elif units[idx["gpu__time_duration.sum"]].strip() in ("ms", "msecond"):
                dur_ns *= 1e6
            wf = float(r[idx["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]].replace(",", ""))
            l2 = float(r[idx["lts__t_sectors.sum"]].replace(",", ""))
            print(f"    shared memory {wf * 128 / dur_ns:.0f} GB/s at 128 B per wavefront "
                  f"(peak 148 SMs x 128 B x 1.965 GHz = 37 230 GB/s), L2 {l2 * 32 / dur_ns:.0f} GB/s of sectors")
        except Exception:  # noqa: BLE001
            pass
        st = []
        for h in stall_cols:
            try:
                st.append((float(r[idx[h]].replace(",", "")), h.split("issue_stalled_")[1].split("_per")[0]))
            except Exception:
                pass
        st.sort(reverse=True)
        print("    stalls:", ", ".join(f"{n}={v:.2f}" for v, n in st[:6]))


if __name__ == "__main__":
    main(sys.argv[1])
