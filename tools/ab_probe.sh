#!/bin/bash
# A/B the library variants under aria_slam_b200/variants/ (plus env-selected kNN variants) with tools/gpu_probe.py
for lib in "" aria_slam_b200/variants/*.so; do
  ASLAM_B200_LIB=${lib:+$PWD/$lib} python tools/gpu_probe.py > /tmp/p.json 2>/dev/null
  python - "$lib" <<'PY'
import json,sys
d=json.load(open('/tmp/p.json'))
print(sys.argv[1] or "default", round(d['ms_per_batch256'],3), {k:round(v[0],3) for k,v in d['stages_ms_per_batch'].items() if v[1]})
PY
done
for c in ${ASLAM_AB_CSA:-}; do ASLAM_KNN_CSA=$c python tools/gpu_probe.py > /tmp/p.json 2>/dev/null; python -c "
import json; d=json.load(open('/tmp/p.json')); print('csa$c', round(d['stages_ms_per_batch']['knn2_partial'][0],3))"; done
