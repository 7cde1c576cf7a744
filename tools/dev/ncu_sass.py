"""Dev tool: SASS listing with executed counts and stall samples, in address order.  usage: ncu_sass.py rep > out.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
for r in rows:
    if len(r) > 8 and r[0] == "Address": hdr = r; i_inst = hdr.index("Instructions Executed"); i_s = hdr.index("# Samples"); i_t = hdr.index("Avg. Threads Executed"); continue
    if hdr is None or len(r) < 8: continue
    print("%s %9s %5s t%-3s %s" % (r[0][-5:], r[i_inst], r[i_s], r[i_t], r[1].strip()))
