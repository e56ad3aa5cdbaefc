# Round-2 A/B sweep (development): make -C mallie_b200/csrc DEV=1, then under gpurun:  bash tools/ab/ab_r2.sh
mkdir -p gpurun_out
run() { env "$@" python tools/ab_frame.py 2>&1 | tail -1; }
run A=1
run MB200_FRAME_FUSED=0
run MB200_TRACE_VAR=1
run MB200_TRI_LAYOUT=96
run MB200_TRI_LAYOUT=96 MB200_TRACE_VAR=1
run MB200_TRACE_VAR=104
run MB200_TRACE_VAR=104 MB200_TRI_LAYOUT=96
run MB200_FRAME_FUSED=0 MB200_TRI_LAYOUT=96
run MB200_FRAME_FUSED=0 MB200_TRACE_VAR=1
run MB200_FRAME_FUSED=0 MB200_TRACE_VAR=1 MB200_TRI_LAYOUT=96
# Woop record (not bit-exact): frame time only, the image differs by construction
run MB200_TRI_LAYOUT=woop
