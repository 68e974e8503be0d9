"""Recipe for oracle/_ref: the UNMODIFIED reference modules, made available where /root/reference does not exist.

    python oracle/build_ref.py            (run in the build container; __graft_entry__.build() calls it)

Copies /root/reference/src/{__init__,models,loss}.py byte for byte into oracle/_ref/ref_src/ and records their
SHA-256 in oracle/_ref/MANIFEST.json.  oracle/_ref/ is git-ignored (reference sources never enter the history)
but NOT gpurun-ignored, so it travels to the GPU box with the snapshot - like the built libsrk.so.  There it is
  * the checker of the full-size GPU parity tests (tests/test_gpu_fullsize.py: the reference's own nn.Modules
    in fp32 on the same GPU), and
  * the thing timed by `bench.py --impl reference` / the cpu_baseline leg (`kind: "reference"`).
src/metrics.py is not copied: it needs torchmetrics / lpips, which are not installable here.
TEST / MEASUREMENT INFRASTRUCTURE ONLY - nothing under food101-super-resolution_b200/ may import it."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ("__init__.py", "models.py", "loss.py")


def build(reference=None, quiet=False):
    """-> path of oracle/_ref/ref_src, or None when neither the reference nor an earlier copy is available."""
    reference = reference or os.environ.get("SR_REFERENCE", "/root/reference")
    src_dir = os.path.join(reference, "src")
    out_dir = os.path.join(DST, "ref_src")
    if not os.path.isdir(src_dir):
        return out_dir if os.path.exists(os.path.join(out_dir, "models.py")) else None
    os.makedirs(out_dir, exist_ok=True)
    manifest = {"source": src_dir, "files": {}}
    for f in FILES:
        s, d = os.path.join(src_dir, f), os.path.join(out_dir, f)
        shutil.copyfile(s, d)
        manifest["files"][f] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if not quiet:
        print("oracle/_ref: %d reference files copied from %s" % (len(FILES), src_dir))
    return out_dir


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
