"""ctypes loader for the C oracle (oracle/gf_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of gf_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by
gofeature_b200/.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgf_oracle.so")

MODE_CANONICAL = 0
MODE_SEQ32 = 1
MODE_F64 = 2

_lib = None


def build(force=False):
    """Compile libgf_oracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "gf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libgf_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        L.gf_oracle_dot.restype = ctypes.c_double
        L.gf_oracle_dot.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int]
        L.gf_oracle_dot_canonical_scalar.restype = ctypes.c_float
        L.gf_oracle_dot_canonical_scalar.argtypes = [fp, fp, ctypes.c_int]
        L.gf_oracle_score_image.restype = ctypes.c_uint32
        L.gf_oracle_score_image.argtypes = [ctypes.c_float]
        L.gf_oracle_normalize.restype = None
        L.gf_oracle_normalize.argtypes = [fp, fp, ctypes.c_int]
        L.gf_oracle_transpose1d.restype = ctypes.c_int
        L.gf_oracle_maxn_float32.restype = ctypes.c_int
        L.gf_oracle_maxn_float32.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), fp]
        L.gf_oracle_set_search.restype = ctypes.c_int
        L.gf_oracle_scan_topk.restype = ctypes.c_int
        L.gf_oracle_synth_rows.restype = None
        L.gf_oracle_synth_rows.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                           ctypes.c_int, fp]
        L.gf_oracle_synth_rows_mt.restype = None
        L.gf_oracle_synth_rows_mt.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                              ctypes.c_int, fp, ctypes.c_int]
        L.gf_oracle_synth_rows_ex_mt.restype = None
        L.gf_oracle_synth_rows_ex_mt.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_float, ctypes.c_uint32,
                                                 ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, fp,
                                                 ctypes.c_int]
        L.gf_oracle_have_avx2.restype = ctypes.c_int
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def dot(x, q, mode=MODE_CANONICAL):
    x = np.ascontiguousarray(x, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    assert x.shape == q.shape and x.ndim == 1
    return lib().gf_oracle_dot(_fp(x), _fp(q), x.shape[0], mode)


def dot_canonical_scalar(x, q):
    x = np.ascontiguousarray(x, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    return lib().gf_oracle_dot_canonical_scalar(_fp(x), _fp(q), x.shape[0])


def normalize(v):
    """util.go:125-139 NoramlizeFloat32."""
    v = np.ascontiguousarray(v, dtype=np.float32)
    out = np.empty_like(v)
    lib().gf_oracle_normalize(_fp(v), _fp(out), v.shape[0])
    return out


def transpose1d(precision, features):
    """util.go:21-39 FeatureValueTranspose1D on a list of bytes objects."""
    row = len(features)
    if row == 0:
        return b""
    len0 = len(features[0])
    bufs = [ctypes.create_string_buffer(bytes(f), len(f)) for f in features]
    arr = (ctypes.c_void_p * row)(*[ctypes.cast(b, ctypes.c_void_p) for b in bufs])
    out = ctypes.create_string_buffer(max(1, (len0 // max(precision, 1)) * row * precision))
    rc = lib().gf_oracle_transpose1d(ctypes.c_int(precision), arr, ctypes.c_int(row), ctypes.c_int(len0), out)
    if rc != 0:
        raise ValueError("ErrBadTransposeValue")
    return out.raw[: (len0 // precision) * row * precision]


def maxn_float32(vector, limit):
    """util.go:76-93 MaxNFloat32 with the canonical tie order (index ascending)."""
    v = np.ascontiguousarray(vector, dtype=np.float32)
    n = v.shape[0]
    m = max(0, min(limit, n))
    idx = np.empty(max(m, 1), dtype=np.int32)
    val = np.empty(max(m, 1), dtype=np.float32)
    got = lib().gf_oracle_maxn_float32(_fp(v), n, limit, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _fp(val))
    return idx[:got].copy(), val[:got].copy()


def set_search(block_rows, heights, live, queries, limit, threshold, mode=MODE_CANONICAL):
    """FeatureSet.Search (set.go:118-159) over explicit block arrays.

    block_rows[b]: float32 [>=heights[b], dims]; live[b]: uint8 [>=heights[b]].
    Returns per query a list of (score, block, row).
    """
    nb = len(block_rows)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    Q, dims = queries.shape
    rows_c = [np.ascontiguousarray(r, dtype=np.float32) for r in block_rows]
    live_c = [np.ascontiguousarray(l, dtype=np.uint8) for l in live]
    rows_p = (ctypes.c_void_p * max(nb, 1))(*[r.ctypes.data for r in rows_c])
    live_p = (ctypes.c_void_p * max(nb, 1))(*[l.ctypes.data for l in live_c])
    h = np.ascontiguousarray(heights, dtype=np.int32)
    lim = max(limit, 0)
    counts = np.zeros(max(Q, 1), dtype=np.int32)
    scores = np.zeros(max(Q * lim, 1), dtype=np.float64)
    oblock = np.zeros(max(Q * lim, 1), dtype=np.int32)
    orow = np.zeros(max(Q * lim, 1), dtype=np.int32)
    lib().gf_oracle_set_search(rows_p, h.ctypes.data_as(ctypes.c_void_p), live_p, ctypes.c_int(nb),
                               ctypes.c_int(dims), _fp(queries), ctypes.c_int(Q), ctypes.c_int(limit),
                               ctypes.c_float(threshold), ctypes.c_int(mode),
                               counts.ctypes.data_as(ctypes.c_void_p), scores.ctypes.data_as(ctypes.c_void_p),
                               oblock.ctypes.data_as(ctypes.c_void_p), orow.ctypes.data_as(ctypes.c_void_p))
    out = []
    for q in range(Q):
        out.append([(float(scores[q * lim + i]), int(oblock[q * lim + i]), int(orow[q * lim + i]))
                    for i in range(counts[q])])
    return out


def scan_topk(rows, queries, limit, mode=MODE_CANONICAL, nthreads=1):
    """Plain brute-force scan of one contiguous matrix (the CPU baseline)."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    n, dims = rows.shape
    Q = queries.shape[0]
    lim = max(limit, 0)
    counts = np.zeros(max(Q, 1), dtype=np.int32)
    scores = np.zeros(max(Q * lim, 1), dtype=np.float64)
    index = np.zeros(max(Q * lim, 1), dtype=np.int64)
    lib().gf_oracle_scan_topk(_fp(rows), ctypes.c_int64(n), ctypes.c_int(dims), _fp(queries), ctypes.c_int(Q),
                              ctypes.c_int(limit), ctypes.c_int(mode), ctypes.c_int(nthreads),
                              counts.ctypes.data_as(ctypes.c_void_p), scores.ctypes.data_as(ctypes.c_void_p),
                              index.ctypes.data_as(ctypes.c_void_p))
    return [(scores[q * lim: q * lim + counts[q]].copy(), index[q * lim: q * lim + counts[q]].copy())
            for q in range(Q)]


def synth_rows(seed, row0, nrows, dims, normalize=True, nthreads=0):
    """Counter-based synthetic rows, bit-identical to the CUDA fill kernel (large requests use every host core)."""
    out = np.empty((nrows, dims), dtype=np.float32)
    if nthreads <= 0:
        nthreads = (os.cpu_count() or 1) if nrows >= 4096 else 1
    lib().gf_oracle_synth_rows_mt(ctypes.c_uint64(seed), ctypes.c_int64(row0), ctypes.c_int64(nrows),
                                  ctypes.c_int(dims), ctypes.c_int(1 if normalize else 0), _fp(out),
                                  ctypes.c_int(nthreads))
    return out


def synth_rows_ex(seed, row0, nrows, dims, centres=0, spread=0.0, dup_ppm=0, normalize=True, nthreads=0):
    """Clustered / duplicated synthetic rows (gf_oracle_synth_rows_ex), bit-identical to gf_set_add_synthetic_ex."""
    out = np.empty((nrows, dims), dtype=np.float32)
    if nthreads <= 0:
        nthreads = (os.cpu_count() or 1) if nrows >= 4096 else 1
    lib().gf_oracle_synth_rows_ex_mt(ctypes.c_uint64(seed), ctypes.c_uint64(centres), ctypes.c_float(spread),
                                     ctypes.c_uint32(dup_ppm), ctypes.c_int64(row0), ctypes.c_int64(nrows),
                                     ctypes.c_int(dims), ctypes.c_int(1 if normalize else 0), _fp(out),
                                     ctypes.c_int(nthreads))
    return out


def synth_centre(seed, centre, dims):
    """The (un-normalised) centre vector of cluster `centre` of a clustered synthetic set."""
    # a set of one centre whose single row has spread 0 would go through the ordinal hash: restate the hash instead
    def mix(x):
        x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return x ^ (x >> 31)
    s2 = seed ^ 0xC3A7E25
    out = np.empty(dims, dtype=np.float32)
    for c in range(dims):
        h = mix(s2 ^ mix(centre * dims + c))
        out[c] = np.float32(h >> 40) * np.float32(1.0 / 8388608.0) - np.float32(1.0)
    return out
