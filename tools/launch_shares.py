"""Kernel shares of a full 256-frame step: the ncu launch list of `bench.py --steps 2 --warmup 1` (cold-cache,
serialised durations) against the CUDA-event shares of the same command run plain (roofline_stages.*.share).
     python tools/launch_shares.py gpurun_out/r02_v5_launches.csv gpurun_out/r02_v5_bench_steps2.json"""
import csv, json, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, gi, vi = hdr.index('Kernel Name'), hdr.index('Grid Size'), hdr.index('Metric Value')
ui = hdr.index('Metric Unit')
names = {'resize_kernel': 'resize', 'fast_kernel': 'fast_nms', 'blur_march_kernel': 'blur', 'blur_kernel': 'blur', 'select_kernel': 'select_harris',
         'describe_kernel': 'orient_brief', 'knn2_partial_v2_kernel': 'knn2_partial', 'knn2_finish_matches_kernel': 'knn2_finish',
         'fill_track_problems_kernel': 'misc'}
scale = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 'nsecond': 1e-6, 's': 1e3, 'second': 1e3}
# a full-batch step = a run of launches that starts with seven resize launches whose grids have z = 256
steps, cur = [], None
for r in rows[1:]:
    k = r[ki].split('(')[0].replace('void ', '').replace('aslam::', '').split('<')[0]
    full = r[gi].replace(' ', '').rstrip(')').endswith(',256')
    if k == 'resize_kernel' and full and (cur is None or 'knn2_finish' in cur):
        cur = collections.OrderedDict(); steps.append(cur)
    if cur is not None and k in names and ('knn2_finish' not in cur or names[k] == 'knn2_finish'):
        cur[names[k]] = cur.get(names[k], 0.0) + float(r[vi].replace(',', '')) * scale[r[ui]]
steps = [s for s in steps if 'knn2_finish' in s and 'fast_nms' in s]
bench = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
ev = {k: v['share'] for k, v in bench['roofline_stages'].items()}
print('complete 256-frame steps in the launch list:', len(steps))
for s in steps:
    tot = sum(s.values())
    print('  ncu ms %.3f  ' % tot + '  '.join('%s %.3f' % (k, v / tot) for k, v in s.items()))
print('  events       ' + '  '.join('%s %.3f' % (k, ev.get(k, float('nan'))) for k in steps[0]) if steps else 'none')
