mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
for n in 8 4 2 1; do  # the 1m workload at every N, then config 5 (10m) at 8
  if [ $n -eq 1 ]; then python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_scale_n_n1.json 2> gpurun_out/r2_scale_n_n1.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_scale_n_n$n.json 2> gpurun_out/r2_scale_n_n$n.err; fi
  echo "n=$n rc=$?"
done
python - <<'PY'
import json
for n in (1,2,4,8):
    l=[x for x in open(f'gpurun_out/r2_scale_n_n{n}.json').read().splitlines() if x.startswith('{')][-1]
    d=json.loads(l)
    print(n, round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), d['clocks'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 8 --workload 10m --steps 5 --warmup 3 > gpurun_out/r2_scale_10m_n8.json 2> gpurun_out/r2_scale_10m_n8.err; echo "10m n=8 rc=$?"
python - <<'PY'
import json
d=json.loads([x for x in open('gpurun_out/r2_scale_10m_n8.json').read().splitlines() if x.startswith('{')][-1])
print('10m n=8', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],3), 'ms e2e', round(d['e2e']['value'],1), d['parity'])
PY
