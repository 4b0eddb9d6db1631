O=gpurun_out/s4_2gpu
mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -8 ) > $O/pytest.log 2>&1
cat $O/pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
tail -c 400 $O/bench_2gpu.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/s4_2gpu/bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['e2e'], d['config'].get('gather'), d.get('parity'))
P
