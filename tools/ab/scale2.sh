# 2-GPU A/B of the frame exchange: peer-memory kernel (default) against ncclAllGather + row placement
mkdir -p gpurun_out
for g in peer nccl; do
  if [ $g = nccl ]; then export MB200_GATHER=nccl; else unset MB200_GATHER; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_x2_$g.json 2> gpurun_out/r2_x2_$g.err; echo "$g rc=$?"
  python - <<PY
import json
d=json.loads([x for x in open('gpurun_out/r2_x2_$g.json').read().splitlines() if x.startswith('{')][-1])
print('$g', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), d['parity']['ranks_equal'], d['config']['parallelism'][-60:])
PY
done
