# 8-GPU box: the 1m frame at N = 8 through both exchanges, N = 4 and 2, and config 5 (10m) at N = 8
mkdir -p gpurun_out
run() { # name, N, extra bench args
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 2954$2 bench.py --gpus $2 $3 > gpurun_out/$1.json 2> gpurun_out/$1.err; echo "$1 rc=$?"
  python - <<PY
import json
d=json.loads([x for x in open('gpurun_out/$1.json').read().splitlines() if x.startswith('{')][-1])
print('$1', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), d['parity']['ranks_equal'], d['parity'].get('equal_to_oracle'), d['config']['parallelism'][-50:])
PY
}
run r2_scale_n8_peer 8 "--steps 20 --warmup 3 --no-cpu"
MB200_GATHER=nccl run r2_scale_n8_nccl 8 "--steps 20 --warmup 3 --no-cpu"
run r2_scale_n4_peer 4 "--steps 20 --warmup 3 --no-cpu"
run r2_scale_n2_peer 2 "--steps 20 --warmup 3 --no-cpu"
run r2_scale_10m_n8_peer 8 "--workload 10m --steps 5 --warmup 3"
