for b in 1 2 3 4 6 8 16; do GRB_MIRROR_BLOCKS_PER_SM=$b python scripts/mirror_tune.py 2>&1 | tail -1; done
