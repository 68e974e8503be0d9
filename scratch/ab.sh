# A/B of scheduling knobs inside ONE box (boxes differ by 2-3 %)
for rep in 1 2; do
for cfg in "1 1" "1 0" "0 1" "0 0"; do set -- $cfg
  echo -n "overlap=$1 fuse=$2: "; SRK_OVERLAP_WGRAD=$1 SRK_FUSE_BN_REDUCE=$2 python bench.py --no-cpu-baseline --steps 40 2>/dev/null | grep -o '"ms_per_step": [0-9.]*'
done; done
