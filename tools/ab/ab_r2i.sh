# phase vote (one step body per warp iteration; development build)
run() { env "$@" python tools/ab_frame.py 2>&1 | tail -1; }
run A=1
run MB200_TRACE_VAR=24
run MB200_TRACE_VAR=56
