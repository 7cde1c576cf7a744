"""Dev tool: executed warp-instruction histogram by SASS opcode, and by (source line range) buckets.
usage: python tools/dev/ncu_ops.py rep [lo-hi:name,...]"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
ops = collections.Counter(); samp = collections.Counter()
byline = collections.defaultdict(collections.Counter)
cur = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; i_inst = hdr.index("Instructions Executed"); i_s = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if r[0] != "":
        cur = (cur_file, int(r[0])); continue
    if r[2] in ("...", "-"): continue
    try: inst, s = int(r[i_inst]), int(r[i_s])
    except ValueError: continue
    sass = r[3].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
    op = m.group(2).split(".")[0] if m else sass[:10]
    if op in ("SHFL", "LDS", "STS", "LDG", "STG", "ATOMS", "BAR", "REDUX"): op = ".".join(m.group(2).split(".")[:2])
    ops[op] += inst; samp[op] += s
    byline[cur][op] += inst
tot = sum(ops.values()); ts = sum(samp.values())
print("total", tot)
for op, n in ops.most_common(45): print("%-14s %10d %5.1f%%  samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * samp[op] / ts))
if len(sys.argv) > 2:
    for spec in sys.argv[2].split(","):
        rng, name = spec.split(":")
        f = None
        if "@" in rng: rng, f = rng.split("@")
        lo, hi = map(int, rng.split("-"))
        c = collections.Counter()
        for (fn, ln), cc in byline.items():
            if lo <= ln <= hi and (f is None or fn.startswith(f)): c.update(cc)
        print("== %s lines %d-%d: %d (%.1f%%)  " % (name, lo, hi, sum(c.values()), 100.0 * sum(c.values()) / tot), dict(c.most_common(12)))
