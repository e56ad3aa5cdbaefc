# A/B harness (development): `make -C mallie_b200/csrc DEV=1`, then e.g.
#   MB200_TRACE_VAR=1100 python tools/ab_frame.py     variant table: launch_sm_variant() in device/kernels.cu
#   python tools/ab_size.py | ab_small.py | ab_band.py  launch-size sweep, fixed launch cost, one rank's share of a frame
mkdir -p gpurun_out
run() { env "$@" python tools/ab_frame.py 2>&1 | tail -1; }
run A=1
run MB200_FRAME_PIPELINE=0
run MB200_FRAME_LPT=0
