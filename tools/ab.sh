for hs in 48 96 128 160 224 320; do echo "HOT_STEPS=$hs"; MB200_HOT_STEPS=$hs python tools/ab_band.py | grep "N=1\|N=8"; done
