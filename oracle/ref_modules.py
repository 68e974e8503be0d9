"""Loader of the unmodified reference modules (oracle/_ref, built by oracle/build_ref.py) - TEST INFRASTRUCTURE ONLY.

The reference package is called `src`, like the drop-in package of this repository, so its two files are loaded by
path under private module names instead of being put on sys.path:

    ref = load()              # -> namespace with .models (reference src/models.py) and .loss (reference src/loss.py)
    m = ref.models.get_model("RESNET", 4, "cuda")
    crit = ref.loss.get_loss_function("nlpd", "cuda")

available() tells whether oracle/_ref exists (it does wherever __graft_entry__.build() ran with /root/reference
present, and on every box that received that tree).  Callers that find it missing fall back to the pinned port
(oracle/sr_oracle.py) and say so."""
import importlib.util
import os
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.path.join(HERE, "_ref", "ref_src")
_cache = None


def available():
    return os.path.exists(os.path.join(REF_SRC, "models.py")) and os.path.exists(os.path.join(REF_SRC, "loss.py"))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    global _cache
    if _cache is None:
        if not available():
            raise FileNotFoundError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
        _cache = types.SimpleNamespace(models=_load("_sr_reference_models", os.path.join(REF_SRC, "models.py")),
                                       loss=_load("_sr_reference_loss", os.path.join(REF_SRC, "loss.py")))
    return _cache
