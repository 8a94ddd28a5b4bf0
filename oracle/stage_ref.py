"""Stage the UNMODIFIED reference modules of the path under oracle/_ref/ (git-ignored, NOT gpurun-ignored: it travels to
the GPU box like a built .so) so that bench.py's reference arm and its `torch_gpu_baseline` leg time the reference's own code
instead of the oracle port.

TEST / MEASUREMENT INFRASTRUCTURE - never imported by the product package.  Run by __graft_entry__.build() whenever
/root/reference is present (build container); on the GPU box only the staged files are used.  Nothing is edited: the files are
copied byte for byte from where they lie (models/*.py: inpaint_networks, inpaint_tools, pix2pix_model, base_model, networks,
edge_operator, UnetG_CT_mask, __init__), and never enter the git history.

    python -m oracle.stage_ref        # -> oracle/_ref/models/*.py + oracle/_ref/MANIFEST.json (sha256 of every file)
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")


def staged():
    return os.path.isfile(os.path.join(DST, "models", "inpaint_networks.py"))


def stage():
    src = os.path.join(REF, "models")
    if not os.path.isdir(src):
        return False
    os.makedirs(os.path.join(DST, "models"), exist_ok=True)
    manifest = {}
    for name in sorted(os.listdir(src)):
        if not name.endswith(".py"):
            continue
        shutil.copyfile(os.path.join(src, name), os.path.join(DST, "models", name))
        with open(os.path.join(src, name), "rb") as fh:
            manifest["models/" + name] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1)
    return True


def import_reference():
    """Put oracle/_ref on sys.path with the import shims the reference needs on this image (stub modules for the
    packages that are absent; a no-op .cuda() when there is no GPU).  Returns the staged `models` package."""
    import types
    if not staged():
        raise RuntimeError("oracle/_ref is not staged (python -m oracle.stage_ref in the build container)")
    sys.dont_write_bytecode = True
    for name in ("nibabel", "matplotlib", "matplotlib.pyplot", "skimage", "skimage.metrics", "skimage.morphology",
                 "skimage.transform", "skimage.measure", "tensorboardX", "openpyxl"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if DST not in sys.path:
        sys.path.insert(0, DST)
    import torch
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import models  # noqa: F401  (the staged reference package)
        import models.inpaint_networks  # noqa: F401
    return sys.modules["models"]


if __name__ == "__main__":
    print("staged" if stage() else "reference mount absent: nothing staged")
