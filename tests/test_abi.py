"""The C-ABI boundary: libsubstrata_b200.so loads without a GPU and exports every symbol include/sbm_audio.h declares;
the ctypes table (substrata_b200/abi.py) names exactly those symbols.  No compute calls here (no GPU)."""
import ctypes as C
import os
import re

import pytest

import substrata_b200 as sb
from substrata_b200 import abi
from substrata_b200.build import build_library
from harness import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sbm_audio.h")).read()
    return sorted(set(re.findall(r"SBM_API\s+[\w\s\*]+?\b(sbm_\w+)\s*\(", src)))


def test_header_and_ctypes_table_agree():
    assert header_symbols() == sorted(abi.SIGNATURES)


def test_library_builds_and_exports_every_symbol():
    path = build_library()
    lib = C.CDLL(path)
    for name in header_symbols():
        assert hasattr(lib, name), name
    abi.bind(lib, require_all=True)


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device the product refuses to create an engine instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sb.SbmError) as ei:
        sb.MixEngine(512)
    assert ei.value.code == -2


def test_exact_host_helpers_exported():
    lib = sb.lib()
    import numpy as np
    rng = np.random.default_rng(0)
    y = rng.uniform(-50, 50, 20000).astype(np.float32)
    x = rng.uniform(-50, 50, 20000).astype(np.float32)
    ours = np.array([lib.sbm_debug_atan2f(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    libm = C.CDLL("libm.so.6")
    libm.atan2f.restype, libm.atan2f.argtypes = C.c_float, [C.c_float, C.c_float]
    ref = np.array([libm.atan2f(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    assert (ours.view(np.uint32) == ref.view(np.uint32)).all()     # bit-exact restatement of glibc's atan2f
