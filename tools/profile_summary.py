"""profiles/ summaries from an `ncu --page raw --csv` export:
     python tools/profile_summary.py gpurun_out/prof_X_raw.csv profiles/r01_vN
   writes <prefix>_ncu_full_summary.csv (one row per launch, the columns the README discusses) and
   profiles/traffic.json (DRAM bytes per launch, mean per kernel kind) for bench.py's roofline.traffic."""
import csv, json, os, sys

raw, prefix = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'launch__grid_size', 'launch__block_size']
idx = [hdr.index(w) for w in want]
ki = hdr.index('Kernel Name')
short = lambda n: n.split('(')[0].replace('void ', '').replace('aslam::', '')
with open(prefix + '_ncu_full_summary.csv', 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['kernel'] + want)
    w.writerow([''] + [units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([short(r[ki])] + [r[i] for i in idx])
num = lambda v: float(v.replace(',', ''))
ri, wi = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
scale = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}
names = {'resize_kernel': 'resize', 'fast_kernel': 'fast_nms', 'blur_kernel': 'blur', 'blur_march_kernel': 'blur', 'select_kernel': 'select_harris',
         'describe_kernel': 'orient_brief', 'knn2_partial_v2_kernel': 'knn2_partial', 'knn2_finish_matches_kernel': 'knn2_finish'}
acc = {}
for r in rows[2:]:
    k = short(r[ki]).split('<')[0]
    acc.setdefault(names.get(k, k), []).append(num(r[ri]) * scale[units[ri]] + num(r[wi]) * scale[units[wi]])
tr = {k: sum(v) / len(v) for k, v in acc.items()}
tr['_source'] = '%s_ncu_full_summary.csv: dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the launches of a kind), ' \
                'ncu --set full, 256 frames of 752x480, 2000 keypoints' % prefix
json.dump(tr, open(os.path.join(os.path.dirname(prefix) or '.', 'traffic.json'), 'w'), indent=1)
print({k: round(v / 1e6, 1) for k, v in tr.items() if k != '_source'})
# pipes.json: what ncu measured on the pipes of each kernel kind (time-weighted mean over the launches of a kind);
# bench.py attaches it to the stage entries as "ncu" -- a kernel far below its HBM roofline shows here which unit holds it
cols = {'alu_pct': 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'xu_pct': 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'fma_pct': 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'lsu_pct': 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'issue_active_pct': 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex_throughput_pct': 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'l2_throughput_pct': 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram_throughput_pct': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'warps_active_pct': 'sm__warps_active.avg.pct_of_peak_sustained_active', 'registers_per_thread': 'launch__registers_per_thread'}
ti = hdr.index('gpu__time_duration.sum')
pipes = {}
for r in rows[2:]:
    k = names.get(short(r[ki]).split('<')[0], short(r[ki]).split('<')[0])
    t = num(r[ti])
    e = pipes.setdefault(k, {'_t': 0.0, **{c: 0.0 for c in cols}})
    e['_t'] += t
    for c, m in cols.items():
        e[c] += t * num(r[hdr.index(m)])
out = {k: {c: round(v[c] / v['_t'], 2) for c in cols} for k, v in pipes.items() if v['_t'] > 0}
out['_source'] = '%s_ncu_full_summary.csv (ncu --set full --clock-control none, one 256-frame step of tools/ncu_target.py)' % prefix
json.dump(out, open(os.path.join(os.path.dirname(prefix) or '.', 'pipes.json'), 'w'), indent=1)
