"""`ncu --set full` report (.ncu-rep) -> the handful of metrics DESIGN.md / profiles/ quote, as text.
   python tools/ncu_summary.py gpurun_out/x/vq.ncu-rep [header line] > profiles/rN_ncu_full_vq.txt"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum"]


def main(path, header=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    if header:
        print(header)
    for r in rows[2:]:
        print("---- " + r[hdr.index("Kernel Name")][:90])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:82s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
