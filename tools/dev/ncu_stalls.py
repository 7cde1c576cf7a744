"""Dev tool: kernel-wide warp-stall sample totals by reason (source page).  usage: ncu_stalls.py rep"""
import csv, io, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; tot = collections.Counter()
for r in rows:
    if len(r) > 8 and r[0] == "Address": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "(Not Issued)" not in h:
            try: tot[h] += int(r[i])
            except ValueError: pass
s = sum(tot.values())
for k, v in tot.most_common(): print("%-26s %7d %5.1f%%" % (k, v, 100.0 * v / s))
