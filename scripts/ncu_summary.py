"""Summarise an ncu capture (read on the CPU box): headline metrics per kernel + stall mix.
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep | gpurun_out/x_raw.csv"""
import csv, io, subprocess, sys
rep = sys.argv[1]   # an .ncu-rep, or the CSV of `ncu -i x.ncu-rep --page raw --csv` (made on the GPU box: reports of many kernels exceed what gpurun brings back)
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("=====", name[:110])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("  %-82s %s %s" % (w, r[i], units[i]))
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    print("  stalls per issue:", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:8]))
