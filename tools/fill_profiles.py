"""Copy one gpu_validate.sh session (gpurun_out/*_<tag>.*) into profiles/ under the round's names.
    python tools/fill_profiles.py <tag> [round]"""
import collections
import csv
import json
import shutil
import sys

tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r01"
rows = list(csv.reader(open('gpurun_out/prof_conv_%s_raw.csv' % tag)))
h, u = rows[0], rows[1]
want = ['Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__time_duration.sum', 'launch__block_size', 'launch__cluster_size', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max',
        'sm__cycles_elapsed.avg.per_second', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second']
idx = [h.index(w) for w in want if w in h]
with open('profiles/%s_conv3x3_shift_ncu_full.csv' % rnd, 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow([h[i] for i in idx]); w.writerow([u[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
for r in rows[2:]:
    print({h[i]: r[i] for i in idx})
scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
rd = sum(float(r[h.index('dram__bytes_read.sum')]) for r in rows[2:]) / len(rows[2:]) * scale[u[h.index('dram__bytes_read.sum')]]
wr = sum(float(r[h.index('dram__bytes_write.sum')]) for r in rows[2:]) / len(rows[2:]) * scale[u[h.index('dram__bytes_write.sum')]]
json.dump({"kernel": "k_conv3x3_pair", "n_positions": 8192, "dram_bytes_per_launch": rd + wr,
           "algorithmic_bytes_per_launch": 8192 * 289 * 256 * 2 * 2.5 + 9 * 256 * 256 * 2,
           "note": "mean of a no-skip and a skip layer (reads input [+ skip], writes output) at 8,192 positions; bench launches carry up to "
                   "16,384 positions (traffic scales linearly); algorithmic = 2.5 activation tensors of 289x256 bf16 per position + the layer's weights",
           "source": "ncu --set full, profiles/%s_conv3x3_shift_ncu_full.csv (gpurun session %s)" % (rnd, tag)},
          open('profiles/conv_traffic.json', 'w'), indent=1)
for a, b in (("a", "bench_modeA"), ("b", "bench_modeB"), ("match", "bench_match_config4"), ("ref", "bench_reference_arm")):
    shutil.copy("gpurun_out/bench_%s_%s.json" % (tag, a), "profiles/%s_%s.json" % (rnd, b))
rows = list(csv.reader(l for l in open('gpurun_out/launches_%s.csv' % tag) if not l.startswith('==')))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]
kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split('(')[0][:44]
    v = float(r[mv].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1}.get(r[mu], 1e-6)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot, n = sum(a[1] for a in agg.values()), sum(a[0] for a in agg.values())
with open('profiles/%s_launches_summary.txt' % rnd, 'w') as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 900: python bench.py --steps 1 --warmup 1 --games 256 --no-cpu  (gpurun session %s)\n" % tag)
    f.write("total launches %d, total %.3f ms (cold-cache, serialised: compare SHARES)\n" % (n, tot))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("%-44s n=%5d total=%10.3f ms avg=%9.1f us share=%5.1f%%\n" % (k, a[0], a[1], 1e3 * a[1] / a[0], 100 * a[1] / tot))
shutil.copy('gpurun_out/launches_%s.csv' % tag, 'profiles/%s_launches_bench_small.csv' % rnd)
print(open('profiles/%s_launches_summary.txt' % rnd).read()[:900])
