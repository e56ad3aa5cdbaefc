# bounce-queue regrouping (production build): 0 off, 2 by octant inside the shade kernels' CTAs, 1 full counting sort, 3 both
for m in 0 2 1 3; do MB200_SORT_BOUNCES=$m python tools/ab_path.py 2>&1 | grep "^\["; done
