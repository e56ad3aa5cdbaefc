# the randomised frame / parity runs under the non-default code paths (each knob is read once per process)
run() { echo "== $*"; env "$@" python tools/fuzz_frames.py 25 5 2>&1 | tail -1; }
run MB200_FRAME_FUSED=1
run MB200_NODE_OCT=0
run MB200_SORT_BOUNCES=1
run MB200_SORT_BOUNCES=0
run MB200_FRAME_ROWSPLIT=0
run MB200_FRAME_PIPELINE=0
run MB200_FRAME_LPT=0 MB200_HOT_STEPS=6
run MB200_HOT_STEPS=6 MB200_FRAME_BATCH_ITEMS=2000
echo "== MB200_NODE_OCT=0 fuzz_parity"; MB200_NODE_OCT=0 python tools/fuzz_parity.py 30 9 2>&1 | tail -1
echo "== fuzz_parity seed 3"; python tools/fuzz_parity.py 60 3 2>&1 | tail -1
