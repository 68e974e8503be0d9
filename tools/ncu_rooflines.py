"""profiles/r2_ncu_kernels.json (the per-roofline-entry DRAM traffic bench.py reports as `traffic`) from the raw ncu CSV
of tools/prof_kernels.py: the launches appear in the order of bench.kernel_rooflines, two per case.
    python tools/ncu_rooflines.py gpurun_out/r2_ncu_all_raw.csv profiles/r2_ncu_kernels.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v = r[col[name]]
    if v in ("", "n/a"):
        return 0.0
    x = float(v.replace(",", ""))
    u = units[col[name]]
    if name.startswith("dram__bytes"):
        x *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    if name == "gpu__time_duration.sum":
        x *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    return x


L = [(r[col["Kernel Name"]], r) for r in rows[2:]]
L = [(n, r) for n, r in L if "at::" not in n and "pack_weights" not in n and "zero_border" not in n]
pos = 0


def take(substr, count):
    """next `count` launches whose kernel name contains substr (skipping helpers in between)"""
    global pos
    out = []
    while len(out) < count and pos < len(L):
        if substr in L[pos][0]:
            out.append(L[pos][1])
        pos += 1
    return out


def rec(launches, note=None, per_launch_of=None):
    n = per_launch_of or len(launches)
    d = {"launches": len(launches),
         "us": round(sum(val(r, "gpu__time_duration.sum") for r in launches) / n, 3),
         "dram_read": round(sum(val(r, "dram__bytes_read.sum") for r in launches) / n),
         "dram_write": round(sum(val(r, "dram__bytes_write.sum") for r in launches) / n),
         "tensor_pct_active": round(sum(val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") for r in launches) / len(launches), 3),
         "dram_pct": round(sum(val(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed") for r in launches) / len(launches), 3),
         "ncu_kernels": sorted({r[col["Kernel Name"]][:60] for r in launches})}
    d["dram_bytes"] = d["dram_read"] + d["dram_write"]
    if note:
        d["note"] = note
    return d


out = {}
out["conv3x3_c64_fprop+bn_stats"] = rec(take("conv3x3_fold_tc_kernel<0, 1, 1", 2))
out["conv3x3_c64_dgrad+bn_bwd_reduce"] = rec(take("conv3x3_fold_tc_kernel<0, 1, 1", 2), "reads the extra Z tile")
out["conv3x3_c64_dgrad+residual+bn_bwd_reduce"] = rec(take("conv3x3_fold_tc_kernel<0, 1, 1", 2), "reads the residual and the Z tile")
out["conv3x3_c64_dgrad+residual"] = rec(take("conv3x3_fold_tc_kernel<0, 1, 0", 2))
w = take("wgrad3x3_tc_kernel", 1) + take("wgrad_fold_kernel", 1) + take("wgrad3x3_tc_kernel", 1) + take("wgrad_fold_kernel", 1)
out["conv3x3_c64_wgrad"] = rec(w, "wgrad3x3_tc_kernel + wgrad_fold_kernel per launch", per_launch_of=2)
out["conv3x3_64to256_pixelshuffle_prelu@128x128"] = rec(take("conv3x3_tc_kernel<1, 0>", 8), "four 64-channel passes per conv", per_launch_of=2)
out["bn_apply_train+prelu"] = rec(take("bn_apply_kernel", 2))
out["bn_apply_train+residual"] = rec(take("bn_apply_kernel", 2))
out["bn_bwd_apply"] = rec(take("bn_bwd_apply_kernel", 2))
rb = take("bn_bwd_reduce_kernel", 1) + take("bn_bwd_apply_kernel", 1) + take("bn_bwd_reduce_kernel", 1) + take("bn_bwd_apply_kernel", 1)
out["bn_bwd_reduce+bn_bwd_apply"] = rec(rb, "reduce + apply per launch", per_launch_of=2)
start = pos
nl = [r for n, r in L[pos:] if "nlpd" in n]
out["nlpd_fwd+bwd"] = rec(nl, "all nlpd_* launches of one forward + backward", per_launch_of=2)
pl = [r for n, r in L[pos:] if "pixel_loss" in n or "scale_to_float" in n]
out["l1_fwd+bwd"] = rec(pl, "pixel_loss_fwd + pixel_loss_bwd", per_launch_of=2)
out["psnr_sse"] = rec([r for n, r in L[pos:] if "psnr_sse" in n])
out["ssim"] = rec([r for n, r in L[pos:] if "ssim_kernel" in n])
json.dump({"source": "ncu --metrics (time, dram bytes, throughputs, tensor pipe) --clock-control none on tools/prof_kernels.py "
                     "(B=64, C2 layer shapes; cold operands), per roofline entry of bench.py; tools/ncu_rooflines.py",
           "kernels": out}, open(sys.argv[2], "w"), indent=1)
for k, v in out.items():
    print("%-46s %8.1f us  %6.1f MB  tensor %5.1f%%  dram %5.1f%%" % (k, v["us"], v["dram_bytes"] / 1e6, v["tensor_pct_active"], v["dram_pct"]))
