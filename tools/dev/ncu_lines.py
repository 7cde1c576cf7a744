"""Dev tool: per-CUDA-source-line totals (warp instructions executed, stall samples) from an ncu report captured with
--import-source on.   usage: python tools/dev/ncu_lines.py gpurun_out/x.ncu-rep [top_n] [kernel-id]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
lines, cur_file, total_inst, total_samp = [], "", 0, 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; i_inst = hdr.index("Instructions Executed"); i_s = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if r[0] == "":
        continue
    try:
        inst, samp = int(r[i_inst]), int(r[i_s])
    except ValueError:
        continue
    lines.append((inst, samp, cur_file, r[0], r[1].strip()[:110]))
    total_inst += inst; total_samp += samp
print("total warp-instructions %d, samples %d" % (total_inst, total_samp))
print("--- by instructions")
for inst, samp, f, ln, src in sorted(lines, reverse=True)[:top]:
    print("%9d %5.1f%%  s=%5.1f%%  %s:%s  %s" % (inst, 100.0 * inst / total_inst, 100.0 * samp / max(1, total_samp), f, ln, src))
print("--- by samples")
for inst, samp, f, ln, src in sorted(lines, key=lambda x: -x[1])[:top]:
    print("%9d %5.1f%%  s=%5.1f%%  %s:%s  %s" % (inst, 100.0 * inst / total_inst, 100.0 * samp / max(1, total_samp), f, ln, src))
