"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, the Python mirror keeps the reference's module API and checkpoint keys, and the
product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import healthivert_gan_b200 as hv
from healthivert_gan_b200 import _lib
from oracle import generator_ref as gr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "hv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hv_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 20
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/hv_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    assert _lib.lib().hv_version() >= 100
    # ... and nothing else: every extern "C" hv_* symbol of the library is declared (debug hooks included)
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if l.split()[-1].startswith("hv_") and " T " in l}
    assert exported == set(syms), exported ^ set(syms)


def test_layer_table_matches_reference_architecture():
    L = _lib.lib()
    assert L.hv_generator_num_layers() == 47
    name = ctypes.create_string_buffer(64)
    vals = [ctypes.c_int() for _ in range(7)]
    acts = {0: "none", 1: "elu", 2: "relu", 3: "sigmoid"}
    for i, (net, lname, cin, cout, k, stride, pad, dil, act) in enumerate(gr.all_layers()):
        assert L.hv_generator_layer_info(i, name, *[ctypes.byref(v) for v in vals]) == 0
        assert name.value.decode() == f"{net}.{lname}"
        assert [v.value for v in vals[:6]] == [cin, cout, k, stride, pad, dil]
        assert acts[vals[6].value] == act
    assert L.hv_generator_layer_info(47, name, *[None] * 7) < 0
    assert b"out of range" in L.hv_last_error()


def test_state_dict_keys_and_shapes_match_reference_checkpoint_format(synthetic_sd):
    g = hv.Generator({"input_dim": 1, "ngf": 16}, False)
    sd = g.state_dict()
    assert len(sd) == 192
    assert set(sd) == set(synthetic_sd)
    for k, v in synthetic_sd.items():
        assert sd[k].shape == v.shape, k
    g.load_state_dict(synthetic_sd)  # strict
    assert torch.equal(g.coarse_generator.conv5.conv.weight_u, synthetic_sd["coarse_generator.conv5.conv.weight_u"])
    assert hasattr(g.fine_generator, "contextul_attention")  # [sic], reference attribute name


def test_no_cpu_fallback(synthetic_sd):
    g = hv.Generator({"input_dim": 1, "ngf": 16}, False)
    x = torch.zeros(1, 1, 256, 256)
    with pytest.raises(_lib.HvError):
        g(x, x, x, torch.zeros(1))
    with pytest.raises(_lib.HvError):
        hv.Sobel()(x)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "healthivert-gan_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text, f"{f} mentions the oracle"
