# one slot per phase per lane (trace_ds.cuh, development build)
run() { env "$@" python tools/ab_frame.py 2>&1 | tail -1; }
run A=1
run MB200_TRACE_DS=608
run MB200_TRACE_DS=604
run MB200_TRACE_DS=616
run MB200_TRACE_DS=508
run MB200_TRACE_DS=708
run MB200_TRACE_DS=808
