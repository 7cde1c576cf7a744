"""Dev tool: stall-reason totals for SASS address (last 5 hex digits) ranges.  usage: ncu_stalls_range.py rep lo-hi[,lo-hi...]"""
import csv, io, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
for spec in sys.argv[2].split(","):
    lo, hi = [int(x, 16) for x in spec.split("-")]
    hdr = None; tot = collections.Counter(); inst = 0
    for r in rows:
        if len(r) > 8 and r[0] == "Address": hdr = r; ii = hdr.index("Instructions Executed"); continue
        if hdr is None or len(r) < len(hdr): continue
        a = int(r[0][-5:], 16)
        if not (lo <= a <= hi): continue
        try: inst += int(r[ii])
        except ValueError: pass
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "(Not Issued)" not in h:
                try: tot[h] += int(r[i])
                except ValueError: pass
    s = sum(tot.values())
    print("range", spec, "instr", inst, "samples", s, " ".join("%s=%.0f%%" % (k[6:], 100.0 * v / max(s, 1)) for k, v in tot.most_common(8)))
