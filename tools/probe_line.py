import json, sys
d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/probe.log'))
print(round(d['ms_per_batch256'], 3), round(d['frames_per_s_device']), {k: round(v[0], 3) for k, v in d['stages_ms_per_batch'].items() if v[1]})
