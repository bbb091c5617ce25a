"""CPU-only checks of the drop-in boundary: the library loads, exports every symbol declared in include/cmpc.h,
validates arguments, and refuses to run without a GPU (there is no CPU path)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cmpc.h")).read()
    return sorted(set(re.findall(r"\b(cmpc_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol(pkg):
    so = pkg.build_library()
    assert os.path.exists(so)
    L = ctypes.CDLL(so)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "missing export " + n


def test_default_config_matches_reference_constants(pkg):
    from cmpc_b200 import _lib
    c = _lib.default_config(20)
    assert c["N"] == 20 and c["delta"] == 0.01 and c["grav"] == 9.81 and c["mu_fric"] == 0.5      # :11,:18,:41
    assert (c["foot_half_len"], c["foot_half_wid"]) == (0.125, 0.065)                             # :51-52
    assert (c["w_h"], c["w_xy"], c["w_zc"], c["w_foot"], c["w_sym"], c["w_swing"], c["w_rate"]) == (1000, 1, 2000, 1000, 10, 10, 1)
    assert c["pz_max"] == 0.76 and c["box"] == [0.01, 0.005, 0.00005]                             # :230,:259-271
    assert c["relax"] == 1e-8


def test_bad_arguments_are_rejected(pkg):
    from cmpc_b200 import _lib
    L = _lib.load()
    cfg = _lib._Config()
    assert L.cmpc_default_config(0, ctypes.byref(cfg)) != 0
    assert L.cmpc_default_config(65, ctypes.byref(cfg)) != 0
    assert b"N" in L.cmpc_last_error()
    assert L.cmpc_default_config(10, ctypes.byref(cfg)) == 0
    h = ctypes.c_void_p()
    assert L.cmpc_create(ctypes.byref(cfg), 0, 0, ctypes.byref(h)) != 0
    cfg.threads = 48
    assert L.cmpc_create(ctypes.byref(cfg), 4, 0, ctypes.byref(h)) != 0


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.CmpcError) as e:
        pkg.BatchSolver(10, 4)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "online-non-linear-centroidal-mpc-with-stability-guarantees-for-robust-locomotion-of-legged-robots-_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".h", ".cu", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "hostsim" not in src.replace("tests/hostsim", ""), f
