# 64-byte pair nodes (development build)
run() { env "$@" python tools/ab_frame.py 2>&1 | tail -1; }
run A=1
run MB200_NODE_LAYOUT=64 MB200_TRACE_VAR=4
run MB200_NODE_LAYOUT=64 MB200_TRACE_VAR=4 MB200_TRI_LAYOUT=96
