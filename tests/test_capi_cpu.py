"""CPU tests of the boundary: the library loads, exports every symbol the header declares,
host utilities and error tables are right, and there is no silent CPU fallback."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "gofeature_b200.h")).read()
    return sorted(set(re.findall(r"GF_API\s+[^;(]*?\b(gf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from gofeature_b200 import _capi
    L = ctypes.CDLL(built)
    names = _header_symbols()
    assert len(names) >= 60
    for n in names:
        assert hasattr(L, n), "missing export " + n
    assert sorted(_capi.SIGNATURES) == names, "ctypes table and header drifted apart"
    _capi.lib()


def test_error_table_matches_error_go(built):
    """error.go:9-49: names and messages, in file order."""
    from gofeature_b200 import api
    L = api.lib()
    want = [("ErrBufferWriteOutofRange", "try to write too much data into buffer"),
            ("ErrBufferCopyOutofRange", "try to copy too much data from source buffer"),
            ("ErrBufferSliceOutofRange", "slice buffer out of range"),
            ("ErrInvalidBufferData", "invalid buffer data"),
            ("ErrAllocateGPUBuffer", "failed to allocate gpu buffer"),
            ("ErrInvalidDeviceID", "invalid device id"),
            ("ErrInvalidFeautres", "invalid features in request"),
            ("ErrFeatureSetNotFound", "feature set not found"),
            ("ErrInvalidSetState", "invalid set state"),
            ("ErrFeatureSetExist", "feature set is already exist"),
            ("ErrTooMuchGPUMemory", "use too much gpu memory"),
            ("ErrAllocatGPUMemory", "fail to allocate gpu memory"),
            ("ErrSliceGPUBuffer", "fail to slice gpu buffer"),
            ("ErrNotEnoughBlocks", "cache does not have enough blocks"),
            ("ErrOutOfBatch", "requests out of batch limit"),
            ("ErrMismatchDimension", "feature with mismatch dimension"),
            ("ErrWriteInputBuffer", "failed to write input buffer"),
            ("ErrWriteOutputBuffer", "failed to write output buffer"),
            ("ErrBlockIsFull", "block is full"),
            ("ErrBlockUsed", "block is used"),
            ("ErrWriteCudaBuffer", "write to cuda buffer error"),
            ("ErrSliceBuffer", "fail to slice cuda buffer"),
            ("ErrClearCudaBuffer", "clear cuda buffer failed"),
            ("ErrAllDevice", "fail to list all cuda device"),
            ("ErrCreateContext", "fail to create cuda context"),
            ("ErrBadTransposeValue", "invalid transpose value to transpose")]
    for code, (name, msg) in enumerate(want, start=1):
        assert L.gf_status_name(code).decode() == name
        assert L.gf_strerror(code).decode() == msg
        assert getattr(api, name) == code


def test_host_utilities_match_oracle(built):
    """util.go helpers served by the library agree with the oracle's restatement."""
    from gofeature_b200 import api
    from oracle import lib as ol
    rng = np.random.default_rng(1)
    for n in (1, 5, 512):
        v = rng.uniform(-1, 1, n).astype(np.float32)
        assert api.NoramlizeFloat32(v).tobytes() == ol.normalize(v).tobytes()
    assert api.NoramlizeFloat32(np.zeros(4)).tolist() == [0, 0, 0, 0]
    feats = [rng.uniform(-1, 1, 6).astype(np.float32).tobytes() for _ in range(3)]
    assert api.FeatureValueTranspose1D(4, *feats) == ol.transpose1d(4, feats)
    assert api.FeatureValueTranspose1D(4) == b""
    with pytest.raises(api.GoFeatureError) as e:
        api.FeatureValueTranspose1D(4, b"12345")
    assert e.value.name == "ErrBadTransposeValue"
    assert api.TFeatureValue(np.arange(3, dtype=np.float32)) == np.arange(3, dtype=np.float32).tobytes()
    assert api.FeatureToFeatureValue1D(api.Feature(b"ab", "x"), api.Feature(b"cd", "y")) == b"abcd"


def test_keys_order_like_the_oracle(built):
    from gofeature_b200 import api
    from oracle import lib as ol
    vals = [1.0, 0.5, 0.0, -0.0, -0.5, float("inf"), float("-inf"), float("nan"), 1e-30, -1e-30]
    for v in vals:
        k = api.key_make(v, 7)
        assert (k >> 32) == ol.lib().gf_oracle_score_image(ctypes.c_float(v))
        assert api.key_slot(k) == 7
        if v == v:
            assert float(api.key_score(k)) == float(np.float32(v)) + 0.0
    assert api.key_make(0.0, 1) == api.key_make(-0.0, 1)              # -0.0 folds into +0.0
    assert api.key_make(0.5, 3) > api.key_make(0.5, 4) > api.key_make(0.25, 0) > api.key_make(float("nan"), 0)


def test_no_cpu_fallback_without_a_gpu(built):
    """Without a CUDA device context creation must fail loudly (ErrCreateContext / ErrAllDevice)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gofeature_b200 import api
    with pytest.raises(api.GoFeatureError) as e:
        api.NewContext(0)
    assert e.value.name in ("ErrCreateContext", "ErrAllDevice", "ErrInvalidDeviceID")


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "gofeature_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), fn
                assert not re.search(r"gf_oracle_[a-z0-9_]+\s*\(", text), fn   # no call into the oracle
                assert "libgf_oracle" not in text, fn
