"""Turn gpurun_out/r01_launches.csv + gpurun_out/r01_full.ncu-rep into the tracked summaries under profiles/."""
import collections, csv, io, json, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = [r for r in csv.reader(open('gpurun_out/%s_launches.csv' % tag)) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
h = rows[hdr]; rows = rows[hdr + 1:]
ki = h.index('Kernel Name'); vi = h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows:
    agg.setdefault(r[ki].split('(')[0].replace('void ', '').replace('hvpr::', ''), []).append(float(r[vi].replace(',', '')))
ours = {k: v for k, v in agg.items() if any(t in k for t in ('vox_', 'pfn_', 'mem_attn', 'bev_fill'))}
# pfn_kernel<1, 0> is the low-register variant the streaming schedule launches next to the canvas fill: listed, not summed
tot = sum(sum(v) / len(v) for k, v in ours.items() if 'pfn_kernel<1, 0>' not in k)
lines = ["# %s - ncu launch list (gpu__time_duration.sum, --clock-control none): python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline" % tag,
         "# workload: G2 432x496, B=8 frames x 120k pts (LiDAR-like), bf16_rescore memory attention.  Cold-cache, serialised: compare SHARES.",
         "kernel,launches,mean_us,share_of_step_pct"]
for k, v in ours.items():
    m = sum(v) / len(v)
    lines.append("%s,%d,%.1f,%.1f" % (k, len(v), m / 1e3, m / tot * 100))
lines.append("TOTAL,,%.1f,100.0" % (tot / 1e3))
open('profiles/%s_launches_summary.csv' % tag, 'w').write("\n".join(lines) + "\n")
print("\n".join(lines))
out = subprocess.run(("ncu -i gpurun_out/%s_full.ncu-rep --page raw --csv" % tag).split(), capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(out)))
h = rr[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum.per_cycle_elapsed']
idx = [h.index(w) for w in want if w in h]
mult = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}
import os
traffic = json.load(open('profiles/traffic.json')) if os.path.exists('profiles/traffic.json') else {}
w = csv.writer(open('profiles/%s_ncu_full_summary.csv' % tag, 'w'))
w.writerow(["# %s ncu --set full --clock-control none --import-source on (one launch per kernel; caches flushed by ncu)" % tag])
w.writerow([h[i] for i in idx]); w.writerow([rr[1][i] for i in idx])
for r in rr[2:]:
    w.writerow([r[i].split('(')[0] if h[i] == 'Kernel Name' else r[i] for i in idx])
    name = r[h.index('Kernel Name')]
    rd = float(r[h.index('dram__bytes_read.sum')]) * mult[rr[1][h.index('dram__bytes_read.sum')]]
    wr = float(r[h.index('dram__bytes_write.sum')]) * mult[rr[1][h.index('dram__bytes_write.sum')]]
    for kk, vv in {'bev_fill': 'bev_fill', 'mem_attn': 'mem_attn', 'pfn': 'pfn', 'vox_gather': 'voxelize_gather', 'vox_hash': 'voxelize_hash'}.items():
        if kk in name: traffic[vv] = rd + wr
json.dump(traffic, open('profiles/traffic.json', 'w'), indent=1)
print(traffic)
