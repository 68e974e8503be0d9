"""Runs every kernel of bench.kernel_rooflines eagerly (two launches each, distinct operand sets) so that
`ncu --set full` can capture them:   ncu --set full --clock-control none -o gpurun_out/r2_kernels python tools/prof_kernels.py
tools/ncu_summary.py turns the report into profiles/r2_ncu_kernels.json (DRAM bytes, throughput, pipe utilisation)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "food101-super-resolution_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import srk  # noqa: E402

srk.set_compute_dtype("bf16")
dev = torch.device("cuda:0")


def eager(launches):
    torch.cuda.nvtx.range_push("kernel")
    for f in launches[:2]:
        f()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    return 1.0


rows = bench.kernel_rooflines(int(os.environ.get("B", 64)), dev, bench.peaks(), timer=eager)
print([r["kernel"] for r in rows])
