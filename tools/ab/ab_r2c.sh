# one rank's share of an 8-GPU frame on one GPU, under different frame-pipeline settings (development)
run() { echo "== $*"; env "$@" AB_N=1,8 python tools/ab_band.py 2>&1 | tail -2; }
run A=1
run MB200_FRAME_PIPELINE=0
run MB200_FRAME_LPT=0
run MB200_FRAME_FUSED=1
run MB200_FRAME_FUSED=1 MB200_FRAME_PIPELINE=0
run MB200_FRAME_BATCH_ITEMS=1100000
run MB200_FRAME_BATCH_ITEMS=600000
