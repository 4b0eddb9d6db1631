#!/bin/bash
# A/B of environment variants of the c2 bench in one GPU-box visit: bash tools/gpu_ab.sh <tag> "<ENV=..>" "<ENV=..>" ...
TAG=${1:-ab}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
i=0
for v in "$@"; do
  i=$((i+1))
  echo "== variant $i: $v" 
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${BENCH_ARGS} > $OUT/v$i.json 2> $OUT/v$i.err
  python - "$OUT/v$i.json" "$v" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get('roofline') or {}
    kb = d.get('kernel_breakdown') or {}
    print(sys.argv[2], '| value %.1f  ms %.3f  e2e %.1f  conv TF %.1f  sum_kernels %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], r.get('achieved', 0), sum(v['ms_per_step'] for v in kb.values())), d.get('stages'), d.get('step_ms'))
    print({k: round(v['ms_per_step'], 3) for k, v in list(kb.items())[:8]})
except Exception as e:
    print('unreadable', e); print(open(sys.argv[1].replace('.json', '.err')).read()[-1500:])
PY
done
