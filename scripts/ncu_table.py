"""A one-line-per-launch markdown table from the CSV of `ncu -i x.ncu-rep --page raw --csv`.
    python scripts/ncu_table.py gpurun_out/x_raw.csv > profiles/rNN_ncu_kernels.md"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
def col(r, name, fmt="%.1f", scale=1.0):
    if name not in hdr:
        return "-"
    v = r[hdr.index(name)]
    try:
        return fmt % (float(v.replace(",", "")) * scale)
    except ValueError:
        return v or "-"
units = rows[1]
def bytes_mb(r, name):
    if name not in hdr:
        return "-"
    i = hdr.index(name)
    try:
        v = float(r[i])
    except ValueError:
        return "-"
    u = units[i]
    return "%.1f" % (v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0))
def dur_us(r):
    i = hdr.index("gpu__time_duration.sum")
    v = float(r[i])
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[i], 1.0)
print("| # | kernel | grid x block | regs | duration µs | issue active % | FMA pipe % | XU % | tensor pipe % | L1TEX % | DRAM read / write MB | top stalls (per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for n, r in enumerate(rows[2:]):
    name = r[hdr.index("Kernel Name")]
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("dm::", "")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    top = ", ".join("%s %.2f" % (a, v) for v, a in sorted(st, reverse=True)[:3])
    print("| %d | `%s` | %s x %s | %s | %.1f | %s | %s | %s | %s | %s | %s / %s | %s |" % (
        n, name, col(r, "launch__grid_size", "%d"), col(r, "launch__block_size", "%d"), col(r, "launch__registers_per_thread", "%d"),
        dur_us(r), col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        col(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        col(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        col(r, "l1tex__throughput.avg.pct_of_peak_sustained_active"),
        bytes_mb(r, "dram__bytes_read.sum"), bytes_mb(r, "dram__bytes_write.sum"), top))
