# Round-2 A/B sweep, part 2 (development build): top-of-tree staging and the Woop record
mkdir -p gpurun_out
run() { env "$@" python tools/ab_frame.py 2>&1 | tail -1; }
run A=1
run MB200_TRACE_VAR=300
run MB200_TRACE_VAR=2 MB200_TOP_NODES=180
run MB200_TRACE_VAR=2 MB200_TOP_NODES=100
run MB200_TRACE_VAR=2 MB200_TOP_NODES=48
run MB200_TRI_LAYOUT=woop
run MB200_TRI_LAYOUT=96
