for d in 0 64 1 2 4 8 16 24 25 3 7 6; do echo -n "dbg=$d: "; SRK_TC_DBG=$d ONLY=fprop timeout 120 python scratch/prof_conv.py; done
