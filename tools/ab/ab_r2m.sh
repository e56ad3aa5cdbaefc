# phase-sorted CTA (trace_ps.cuh, development build): MB200_TRACE_PS = slots per CTA * 100 + lanes below which a round ends
run() { env "$@" timeout 120 python tools/ab_frame.py 2>&1 | tail -1; }
run A=1
run MB200_TRACE_PS=19216
run MB200_TRACE_PS=19208
run MB200_TRACE_PS=19224
run MB200_TRACE_PS=25616
run MB200_TRACE_PS=16016
