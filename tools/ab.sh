mkdir -p gpurun_out
run() { env "$@" python tools/ab_frame.py 2>&1 | tail -1; }
run MB200_TRACE_VAR=11
run MB200_TRACE_MR=440
run MB200_TRACE_MR=333
run MB200_TRACE_MR=327
run MB200_TRACE_MR=325
run MB200_TRACE_MR=336
run MB200_TRACE_MR=248
