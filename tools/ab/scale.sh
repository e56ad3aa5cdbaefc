mkdir -p gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1h_ref_n1.json 2> gpurun_out/r1h_ref_n1.err
python bench.py > gpurun_out/r1h_bench_n1.json 2> gpurun_out/r1h_bench_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r1h_bench_n$n.json 2> gpurun_out/r1h_bench_n$n.err
  echo "n=$n rc=$?"
done
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        l=[x for x in open(f'gpurun_out/r1h_bench_n{n}.json').read().splitlines() if x.startswith('{')][-1]
        d=json.loads(l)
        print(n, round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), 'frac', d['roofline'] and round(d['roofline']['frac'],3), d['clocks'])
    except Exception as e:
        print(n, 'ERR', e)
PY
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
