mkdir -p gpurun_out
run() { env "$@" python tools/ab_frame.py 2>&1 | tail -1; }
run MB200_TRACE_VAR=11
run MB200_TRACE_POLICY=0
run MB200_TRACE_POLICY=3
run MB200_TRACE_POLICY=4
run MB200_TRACE_POLICY=6
