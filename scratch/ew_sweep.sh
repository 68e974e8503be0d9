for b in 2 4 6 8 12 16; do echo "== blocks/SM $b"; SRK_EW_BLOCKS_PER_SM=$b timeout 200 python scratch/prof_elem.py 2>&1 | grep -v "^$"; done
