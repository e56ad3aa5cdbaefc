# two rays per lane (development build)
run() { env "$@" python tools/ab_frame.py 2>&1 | tail -1; }
run A=1
run MB200_TRACE_TR=1
run MB200_TRACE_TR=5
run MB200_TRACE_TR=9
run MB200_TRACE_TR=105
run MB200_TRACE_TR=205
