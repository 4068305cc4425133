"""Copy one tools/gpu_r02_profiles.sh session (gpurun_out/*_<tag>.*) into profiles/ under the round's names:
the conv kernel's `ncu --set full` rows (two launches at 16,384 positions: a layer without and one with the skip
tensor), conv_traffic.json (bench.py's roofline.traffic), the launch-list summary and raw CSV, and the per-kernel
summaries of the stem / heads / board / tree kernels.
    python tools/fill_profiles_r02.py <tag> [round]"""
import collections
import csv
import json
import shutil
import subprocess
import sys

tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r02"
rows = list(csv.reader(open('gpurun_out/prof_conv_%s_raw.csv' % tag)))
h, u = rows[0], rows[1]
want = ['Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__time_duration.sum', 'launch__block_size', 'launch__cluster_size', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max',
        'sm__cycles_elapsed.avg.per_second', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second']
idx = [h.index(w) for w in want if w in h]
with open('profiles/%s_conv3x3_dense_ncu_full.csv' % rnd, 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow([h[i] for i in idx]); w.writerow([u[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
for r in rows[2:]:
    print({h[i]: r[i] for i in idx})
scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
rd = sum(float(r[h.index('dram__bytes_read.sum')]) for r in rows[2:]) / len(rows[2:]) * scale[u[h.index('dram__bytes_read.sum')]]
wr = sum(float(r[h.index('dram__bytes_write.sum')]) for r in rows[2:]) / len(rows[2:]) * scale[u[h.index('dram__bytes_write.sum')]]
N = 16384
json.dump({"kernel": "k_conv3x3_pair<0>", "n_positions": N, "dram_bytes_per_launch": rd + wr,
           "algorithmic_bytes_per_launch": N * 289 * 256 * 2 * 2.5 + 9 * 256 * 256 * 2,
           "note": "mean of a no-skip and a skip layer (reads input [+ skip], writes output) at 16,384 positions = the bench's launch size "
                   "(its last forward of a search step is smaller; traffic scales linearly); algorithmic = 2.5 activation tensors of "
                   "289x256 bf16 per position + the layer's weights.  Dense layout: no pad rows / pixels and no halo re-reads beyond L2",
           "source": "ncu --set full, profiles/%s_conv3x3_dense_ncu_full.csv (tools/gpu_r02_profiles.sh %s)" % (rnd, tag)},
          open('profiles/conv_traffic.json', 'w'), indent=1)
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
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 1300: python bench.py --steps 1 --warmup 1 --games 256 --no-cpu --no-sub  (%s)\n" % tag)
    f.write("total launches %d, total %.3f ms (cold-cache, serialised: compare SHARES).  k_conv3x3_pair<0> = the 40 tower convs + the stem GEMM of each forward; <1> = the dense-head GEMMs\n" % (n, tot))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("%-44s n=%5d total=%10.3f ms avg=%9.1f us share=%5.2f%%\n" % (k, a[0], a[1], 1e3 * a[1] / a[0], 100 * a[1] / tot))
shutil.copy('gpurun_out/launches_%s.csv' % tag, 'profiles/%s_launches_bench_small.csv' % rnd)
with open('profiles/%s_kernels_ncu_summary.txt' % rnd, 'w') as f:
    for part, title in (("fwd", "one 16,384-position forward of a 2-block tower (tools/bench_tower.py 16384 2): im2col, stem GEMM + 4 convs (<0>), dense-head GEMMs (<1>), k_heads_finish is below the capture window"),
                        ("tree", "board / tree kernels: one real 1,024-game ply per mode (tools/prof_kernels.py 1024; node pool, k_reroot frees instead of copying) + rules kernels on 4,096 positions")):
        out = subprocess.run([sys.executable, "tools/ncu_summary.py"], stdin=open('gpurun_out/prof_%s_%s_raw.csv' % (part, tag)), stdout=subprocess.PIPE, text=True).stdout
        f.write("== %s\n%s\n" % (title, out))
shutil.copy('gpurun_out/prof_tree_%s_raw.csv' % tag, 'profiles/%s_board_tree_kernels_ncu_raw.csv' % rnd)
shutil.copy('gpurun_out/prof_units_%s.json' % tag, 'profiles/%s_board_tree_kernels_units.json' % rnd)
print(open('profiles/%s_launches_summary.txt' % rnd).read()[:900])
print(open('profiles/%s_kernels_ncu_summary.txt' % rnd).read())
