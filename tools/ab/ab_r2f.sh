# octant copies of the pair nodes (development build)
run() { env "$@" python tools/ab_frame.py 2>&1 | tail -1; }
run A=1
run MB200_NODE_OCT=1 MB200_TRACE_VAR=8
run MB200_NODE_OCT=1 MB200_TRACE_VAR=8 MB200_TRI_LAYOUT=96
