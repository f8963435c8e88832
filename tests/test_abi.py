"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
that include/scn_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from sparse_rcnn_b200 import _lib, build


@pytest.fixture(scope="module")
def dll():
    build.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_parses_every_prototype():
    protos = _lib.parse_header()
    src = open(_lib.HEADER).read()
    declared = set(re.findall(r"\b(scn_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", src, flags=re.S)))
    assert declared == set(protos), declared ^ set(protos)
    assert len(protos) >= 40


def test_library_exports_every_symbol(dll):
    for name in _lib.parse_header():
        assert hasattr(dll, name), "missing symbol %s" % name


def test_error_reporting_without_gpu(dll):
    _lib.LIB.load()
    # invalid argument -> status code + message, never an abort
    rc = _lib.raw("scn_pack_coords")(0, 4, 7, 0, 0, 0)
    assert rc == 1
    assert b"ncol" in _lib.raw("scn_last_error")()
    with pytest.raises(RuntimeError, match="power of two"):
        _lib.call("scn_hash_clear", 0, 0, 100, 0)
    assert _lib.raw("scn_scan_tmp_elems")(10) >= 2
    assert _lib.raw("scn_conv_weight_image_bytes")(27, 32, 32) == 27 * 1 * 32 * 128
    assert _lib.raw("scn_conv_weight_image_bytes")(8, 44, 22) == 8 * 2 * 32 * 128


def test_no_cpu_fallback_in_product():
    import torch
    from sparse_rcnn_b200 import scn
    with pytest.raises(RuntimeError, match="CUDA"):
        scn.ReLU()(scn.SparseConvNetTensor(torch.zeros(2, 2), None, None))


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "sparse_rcnn_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                assert "scn_oracle" not in src and "import oracle" not in src, os.path.join(d, f)
