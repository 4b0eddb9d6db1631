#!/bin/bash
# One GPU-box visit: the -m gpu suite, the two bench workloads, the reference arm and A/B runs of opt-in variants.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [ab]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( time IA_TEST_OPTIN=1 timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -150 ) > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_c2.json 2> $OUT/bench_c2.err
timeout 900 python bench.py --workload c3 --steps 5 --warmup 3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err
if [ "$2" = "ab" ]; then
  IA_FIR_NOISE_PREFETCH=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_c2_firnpf.json 2> $OUT/bench_c2_firnpf.err
  IA_RENDER_MLP=fp16 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_c2_mlp16.json 2> $OUT/bench_c2_mlp16.err
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
fi
tail -3 $OUT/pytest.log
grep -E "[0-9]+ (passed|failed)" $OUT/pytest.log | tail -1
for f in $OUT/bench_*.json; do echo $f; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get('roofline') or {}
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', (d.get('e2e') or {}).get('value'), 'frac', r.get('frac'), 'conv TF', r.get('achieved'),
          'parity', d.get('parity'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'stages', d.get('stages'))
    kb = d.get('kernel_breakdown') or {}
    print({k: round(v['ms_per_step'], 3) for k, v in list(kb.items())[:12]})
except Exception as e:
    print('unreadable', e); print(open(sys.argv[1].replace('.json', '.err')).read()[-1500:])
PY
done
