# canonical pair nodes (no octant copies) for the path-traced frames, regrouping off / on
MB200_NODE_OCT=0 MB200_SORT_BOUNCES=0 python tools/ab_path.py 2>&1 | grep "^\["
MB200_NODE_OCT=0 MB200_SORT_BOUNCES=1 python tools/ab_path.py 2>&1 | grep "^\["
