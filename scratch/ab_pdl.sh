# programmatic dependent launch per kernel class (SRK_PDL bits: 1 conv, 2 wgrad, 4 BN) inside ONE box
mkdir -p gpurun_out
for m in 0 1 2 4 3 5 6 7 0; do
  echo -n "C2 pdl=$m: "; SRK_PDL=$m timeout 300 python bench.py --no-cpu-baseline --steps 40 2>gpurun_out/ab_pdl_err.log | grep -o '"ms_per_step": [0-9.]*'
done
for m in 0 1 2 4 7 0; do
  echo -n "C3 pdl=$m: "; SRK_PDL=$m timeout 300 python bench.py --config C3 --no-cpu-baseline --steps 20 2>>gpurun_out/ab_pdl_err.log | grep -o '"ms_per_step": [0-9.]*'
done
