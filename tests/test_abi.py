"""CPU-side checks of the drop-in boundary: libsrk.so loads and exports every symbol include/srk.h
declares, the ctypes table mirrors the header, the nn.Modules keep the reference's state_dict layout,
and nothing silently runs on the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "srk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(srk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from srk import _lib
    syms = _header_symbols()
    assert len(syms) >= 30
    lib = ctypes.CDLL(os.path.abspath(_lib.LIB_PATH))
    for s in syms:
        assert hasattr(lib, s), "libsrk.so does not export %s" % s
    assert sorted(_lib.SIGNATURES) == syms, "ctypes table and include/srk.h disagree"
    assert _lib.cdll.srk_version() >= 100


def test_ctypes_arity_matches_header():
    from srk import _lib
    text = open(os.path.join(ROOT, "include", "srk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, args in re.findall(r"\b(srk_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = args.strip()
        n = 0 if args in ("", "void") else args.count(",") + 1
        assert len(_lib.SIGNATURES[name][1]) == n, name


@pytest.mark.parametrize("arch", ["SRCNN", "RESNET", "AttentionSR"])
def test_state_dict_layout_and_seeded_init_match_reference(arch):
    from src.models import get_model
    fix = load_golden("state_dicts")
    torch.manual_seed(0)
    sd = get_model(arch, scale_factor=4).state_dict()
    assert list(sd.keys()) == [str(k) for k in fix[arch + "/keys"]]
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in fix[arch + "/shapes"]]
    assert [str(v.dtype) for v in sd.values()] == [str(s) for s in fix[arch + "/dtypes"]]
    # same RNG consumption as the reference constructors -> identical seeded weights
    got = np.array([float(v.double().sum()) for v in sd.values()])
    np.testing.assert_allclose(got, fix[arch + "/checksum"], rtol=0, atol=1e-9)


def test_get_model_and_loss_factory_errors():
    from src.loss import get_loss_function
    from src.models import get_model
    with pytest.raises(ValueError, match="Unknown architecture"):
        get_model("VDSR")
    with pytest.raises(ValueError, match="Unknown loss function"):
        get_loss_function("huber", "cpu")
    assert get_loss_function("NLPD", "cpu").kernel.shape == (3, 1, 5, 5)
    assert "kernel" in get_loss_function("nlpd", "cpu").state_dict()


def test_cpu_tensors_are_rejected_not_silently_computed():
    from src.loss import get_loss_function
    from src.metrics import MetricsCalculator
    from src.models import get_model
    x = torch.rand(1, 3, 8, 8)
    for arch in ("SRCNN", "RESNET"):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            get_model(arch)(x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        get_loss_function("mae", "cpu")(x, x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        get_loss_function("nlpd", "cpu")(x, x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MetricsCalculator("cpu").compute(torch.rand(1, 3, 16, 16), torch.rand(1, 3, 16, 16))


def test_missing_library_fails_loudly(tmp_path):
    import subprocess
    import sys
    env = dict(os.environ, SRK_LIB=str(tmp_path / "nope.so"), PYTHONPATH=os.path.join(ROOT, "food101-super-resolution_b200"))
    r = subprocess.run([sys.executable, "-c", "import srk"], env=env, capture_output=True, text=True)
    assert r.returncode != 0 and "libsrk.so not found" in r.stderr


def test_product_code_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "food101-super-resolution_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "sr_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f


def test_nlpd_workspace_plan_is_host_only_and_aligned():
    """srk_nlpd_workspace_bytes is pure host arithmetic (no GPU needed): it rejects level counts outside [1, 6], grows
    with every dimension, and holds at least the pyramid of the difference image (sum of the level sizes, fp32), the
    int8 sign maps, the two gradient ping-pong buffers and the per-block fp64 partial sums of the loss terms
    (8 terms x 148 * 8 blocks: the loss value is a fixed-order sum, not an atomic accumulation); every region is padded to
    16 bytes (the exact-2x kernels use float2 / char2 accesses), so odd sizes cost at most a few bytes per region."""
    from srk import _lib
    f = _lib.cdll.srk_nlpd_workspace_bytes
    assert f(1, 3, 16, 16, 0) == -1 and f(1, 3, 16, 16, 7) == -1
    for (n, c, h, w, L) in [(1, 3, 4, 4, 4), (2, 3, 25, 37, 4), (1, 1, 11, 11, 3), (64, 3, 256, 256, 4)]:
        hs, ws = [h], [w]
        for _ in range(L):
            hs.append((hs[-1] + 1) // 2)
            ws.append((ws[-1] + 1) // 2)
        nc = n * c
        floats = sum(nc * a * b for a, b in zip(hs, ws)) + nc * hs[0] * ws[0] + nc * hs[1] * ws[1]
        signs = sum(nc * hs[l] * ws[l] for l in range(L))
        need = 4 * floats + signs + 8 * 8 * (148 * 8)
        got = f(n, c, h, w, L)
        assert got % 16 == 0 and need <= got <= need + 16 * (2 * L + 6), (n, c, h, w, L, got, need)
        assert f(n + 1, c, h, w, L) > got and f(n, c, h + 2, w, L) > got


SWEEP_SHA256 = {   # sha256 of /root/reference/configs/* (taken where the reference exists)
    "sweep_attentionSR.yaml": "a3b23415136e75d8d328fb9e32eb546c00d889a5838872486aab8d81e2883de2",
    "sweep_resnet.yaml": "52dfb82a12c41095c46110a2f856ce353e82d42c48cc3b347a9a3a6da3040eb7",
    "sweep_srcnn.yaml": "b35af37966fe55b398b4a9e13bb91ffa55887acd22bc808ad7612a0b20d4d046",
    "sweep_tuning.yaml": "b9e4dfcd451cc3c135a4bc84f68820447b1e470d28e1a0584d79344f7637d5e1",
    "sweep_winners.txt": "c8a059ae4d9437976f23d65db5502305799d2d5d833ce66507d41e2e42258b5f",
}


def test_sweep_configs_are_the_reference_files():
    """The sweep entry points (reference configs/sweep_*.yaml, sweep_winners.txt) ship next to train.py, byte for byte
    (wandb sweeps name `program: train.py` and the flags train.py parses)."""
    import hashlib
    want = SWEEP_SHA256
    cfg_dir = os.path.join(ROOT, "food101-super-resolution_b200", "configs")
    for name in want:
        path = os.path.join(cfg_dir, name)
        assert os.path.exists(path), name
        digest = hashlib.sha256(open(path, "rb").read()).hexdigest()
        assert digest == SWEEP_SHA256[name], (name, digest)
    import yaml
    for name in want:
        if name.endswith(".yaml"):
            doc = yaml.safe_load(open(os.path.join(cfg_dir, name)))
            assert doc["program"] == "train.py" and doc["metric"]["name"] == "val_psnr"
            assert set(doc["parameters"]) <= {"architecture", "loss_function", "lr", "batch_size", "subset", "epochs",
                                              "patience", "pretrained_weights", "save_name"}
