"""Per-instruction stall samples of one kernel in an .ncu-rep: top instructions by samples with their
dominant stall reason, and totals per stall reason.
    python scripts/ncu_hotspots.py rep.ncu-rep <kernel-name-substring> [top]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kern.append(cur)
    elif cur is not None:
        if cur["hdr"] is None:
            cur["hdr"] = r
        else:
            cur["rows"].append(r)
for k in kern:
    if pat not in k["name"]:
        continue
    h = k["hdr"]
    iS, iI, isrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    sc = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    R = k["rows"]
    tot = sum(int(r[iS]) for r in R)
    print(k["name"][:100], "instructions", len(R), "samples", tot)
    agg = {h[i]: sum(int(r[i] or 0) for r in R) for i in sc}
    print("  totals:", ", ".join("%s %.1f%%" % (n.replace("stall_", ""), 100.0 * v / tot) for n, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
    order = sorted(range(len(R)), key=lambda i: -int(R[i][iS]))[:top]
    for i in sorted(order):
        r = R[i]
        st = sorted(((int(r[j] or 0), h[j].replace("stall_", "")) for j in sc), reverse=True)[:2]
        print("  #%5d %5.2f%% exec %9s  %-70s %s" % (i, 100.0 * int(r[iS]) / tot, r[iI], r[isrc].strip()[:70], st))
    break
