"""gpurun_out/r02_cfg5_launches.csv (ncu launch list of bench.py --config cfg5) -> profiles/r02_cfg5_launches_summary.csv:
per kernel name, launches and mean duration of the LAST step captured, with its share of that step."""
import collections, csv
rows = [r for r in csv.reader(open('gpurun_out/r02_cfg5_launches.csv')) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
h = rows[hdr]; rows = rows[hdr + 1:]
ki = h.index('Kernel Name'); vi = h.index('Metric Value')
names = [r[ki].split('(')[0].replace('void ', '').replace('hvpr::', '') for r in rows]
vals = [float(r[vi].replace(',', '')) / 1e3 for r in rows]
# one detector step = the launches from one vox_init_kernel to the next
starts = [i for i, n in enumerate(names) if n.startswith('vox_init_kernel')]
# bench.py calibrates the head bias after the first step (library kernels, a re-plan): take the LAST complete step
a, b = (starts[-2], starts[-1]) if len(starts) >= 2 else (starts[-1], len(names))
agg = collections.OrderedDict()
skipped = 0
for n, v in zip(names[a:b], vals[a:b]):
    if n.startswith('native::') or n.startswith('at_cuda_detail::'):       # torch.zeros of a (re)plan inside the window: not part of a step
        skipped += 1
        continue
    agg.setdefault(n, []).append(v)
tot = sum(sum(v) for v in agg.values())
lines = ["# r02 - ncu launch list (gpu__time_duration.sum, --clock-control none) of ONE points -> detections step (cfg5): python bench.py --config cfg5 --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline",
         "# 8 frames x 120k points, G2; cold-cache, serialised: compare SHARES.  %d launches of this repo's kernels in the step (%d buffer-allocation fills of the eager plan left out)" % (b - a - skipped, skipped),
         "kernel,launches,total_us,share_of_step_pct"]
for n, v in agg.items():
    lines.append("%s,%d,%.1f,%.1f" % (n, len(v), sum(v), 100 * sum(v) / tot))
lines.append("TOTAL,%d,%.1f,100.0" % (b - a - skipped, tot))
open('profiles/r02_cfg5_launches_summary.csv', 'w').write("\n".join(lines) + "\n")
print("\n".join(lines))
