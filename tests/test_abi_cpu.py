"""The C-ABI library loads without a GPU and exports every symbol include/hs_b200.h declares."""
import os
import re

import pytest

from conftest import ROOT
from hyperscanning_signal_analysis_b200 import _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.hs_version() >= 100
    assert lib.hs_launch_count() == 0 or lib.hs_launch_count() > 0


def test_workspace_queries_need_no_gpu():
    lib = _lib.load()
    assert lib.hs_transfer_ws_bytes(599, 38, 8, 256) > 8 * 256 * 16
    assert lib.hs_filtfilt_ws_bytes(38, 15360) >= 0


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from hyperscanning_signal_analysis_b200 import mtmvar
    with pytest.raises(_lib.HsError):
        mtmvar.ar_coeff(np.zeros((4, 100)), 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hyperscanning_signal_analysis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f
