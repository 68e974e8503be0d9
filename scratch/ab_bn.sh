for rep in 1 2; do for b in 4 8 3 6; do echo -n "bn blocks/SM=$b: "; SRK_EW_BN_PER_SM=$b python bench.py --no-cpu-baseline --steps 40 2>/dev/null | grep -o '"ms_per_step": [0-9.]*'; done; done
