mkdir -p gpurun_out
for n in 8 4 2 1; do
  if [ $n -eq 1 ]; then python bench.py --steps 30 --warmup 3 --no-cpu > gpurun_out/r1n_bench_n1.json 2> gpurun_out/r1n_bench_n1.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 30 --warmup 3 --no-cpu > gpurun_out/r1n_bench_n$n.json 2> gpurun_out/r1n_bench_n$n.err; fi
  echo "n=$n rc=$?"
done
python - <<'PY'
import json
for n in (1,2,4,8):
    l=[x for x in open(f'gpurun_out/r1n_bench_n{n}.json').read().splitlines() if x.startswith('{')][-1]
    d=json.loads(l)
    print(n, round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), d['clocks'])
PY
