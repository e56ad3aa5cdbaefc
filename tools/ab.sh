mkdir -p gpurun_out
run() { env "$@" python tools/ab_frame.py 2>&1 | tail -1; }
run A=1
for v in 3016 3008 3116 3404; do
run MB200_TRACE_VAR=$v
MB200_TRACE_VAR=$v python tools/ab_small.py | tail -2
done
