# occupancy / refill / stack sweep on top of the octant node copies (development build)
run() { env "$@" python tools/ab_frame.py 2>&1 | tail -1; }
run A=1
run MB200_TRACE_VAR=409
run MB200_TRACE_VAR=410
run MB200_TRACE_VAR=407
run MB200_TRACE_VAR=406
run MB200_TRACE_VAR=506
run MB200_TRACE_VAR=512
run MB200_TRACE_VAR=308
run MB200_TRACE_VAR=316
run MB200_TRACE_VAR=300
run MB200_TRI_LAYOUT=96
