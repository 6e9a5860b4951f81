"""CPU: the C-ABI library builds/loads and exports every symbol include/pifu_b200.h declares."""
import ctypes
import os
import re

from helpers import ROOT

from pifu_b200 import _lib, build


def _declared():
    src = open(os.path.join(ROOT, "include", "pifu_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pifu_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    lib.pifu_abi_version.restype = ctypes.c_int
    assert lib.pifu_abi_version() >= 1


def test_no_cpu_fallback():
    """Without a CUDA device the product path refuses to run instead of computing on the host."""
    import pytest
    import torch
    from pifu_b200 import PIFuNetwNML, config
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    net = PIFuNetwNML(config.coarse_opt())
    net.im_feat_list = [torch.zeros(1, 256, 8, 8)]
    with pytest.raises(Exception):
        net.query(torch.zeros(1, 3, 16), torch.eye(4)[None])
