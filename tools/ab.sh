mkdir -p gpurun_out
run() { env "$@" python tools/ab_frame.py 2>&1 | tail -1; }
run A=1
run MB200_TRACE_VAR=1100
run MB200_TRACE_VAR=1104
run MB200_TRACE_VAR=2009
run MB200_TRACE_VAR=2129
run MB200_TRACE_VAR=2010
run MB200_TRACE_VAR=1112
