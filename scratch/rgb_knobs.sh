for d in 0 1 2 4 8 16 17 3 7 15 31; do echo -n "dbg=$d: "; SRK_RGB_DBG=$d python scratch/prof_rgb.py 2>/dev/null | grep "outconv bwd" | tr '\n' ' '; echo; done
