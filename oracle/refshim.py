"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box, so nothing in the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this at run time; it is used by
``oracle/make_golden.py`` and by the CPU-side ``tests/test_oracle_vs_reference.py``
(which skips when the mount is absent).

Shims: empty stub modules for the packages the container lacks, and a
no-op ``.cuda()`` on CPU-only hosts.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def install():
    if not available():
        raise RuntimeError("reference mount /root/reference is absent")
    sys.dont_write_bytecode = True
    for name in ("nibabel", "matplotlib", "matplotlib.pyplot", "skimage", "skimage.metrics",
                 "skimage.morphology", "skimage.transform", "skimage.measure", "tensorboardX",
                 "openpyxl", "cv2_stub"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    for p in (REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "evaluation")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self


def reference_generator(state_dict=None, use_cuda=False):
    install()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from models.inpaint_networks import Generator
    g = Generator({"input_dim": 1, "ngf": 16}, use_cuda)
    if state_dict is not None:
        g.load_state_dict(state_dict)
    g.eval()
    return g
