mkdir -p gpurun_out
run() { env "$@" python tools/ab_frame.py 2>&1 | tail -1; }
run A=1
python tools/ab_small.py | tail -2
python tools/ab_size.py | head -4
run MB200_TRACE_VAR=1035
MB200_TRACE_VAR=1035 python tools/ab_small.py | tail -2
MB200_TRACE_VAR=1035 python tools/ab_size.py | head -4
