"""ncu report -> the per-launch summary CSV kept under profiles/.  usage: ncu_summarize.py <report.ncu-rep> <out.csv>"""
import csv, subprocess, sys

COLS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_inst_executed_op_tmem_ldt.sum", "smsp__sass_inst_executed_op_tmem_stt.sum"]
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = [c for c in COLS if c in hdr]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(keep)
    w.writerow([units[hdr.index(c)] for c in keep])
    for r in rows[2:]:
        w.writerow([r[hdr.index(c)] for c in keep])
print(out, len(rows) - 2, "launches,", len(keep), "columns")
