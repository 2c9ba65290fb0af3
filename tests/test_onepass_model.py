"""CPU model of the one-pass chain's SELECTION and CERTIFICATE logic (gf_gemm.cu: GROUPTOP epilogue mode +
grouptop_tail_kernel), restated in numpy, against the oracle's exact answer.

What it pins (no GPU needed): whenever the tail certifies a query, the K keys it emits are exactly the canonical
float32 top-K -- on random data, on rows planted inside one 32-row group (only a group's best two keys are reported),
on near-duplicates, un-normalised rows and tombstones -- and the certificate is not vacuous (random data is
certified).  The arithmetic of the int8 shadow (one scale per row / per query, exact int32 accumulation), the key
packing (score image with the lane in its low 5 bits), the cut (histogram of per-thread maxima), the group
expansion and the error bound delta follow the kernels line by line; the GPU tests compare the kernels themselves
with the oracle.
"""
import numpy as np
import pytest

from oracle import lib as ol

K_BINS_SHIFT = 20
CLUSTER, WARPS, REGION = 8, 16, 128


def _image(v):
    u = np.ascontiguousarray(v, dtype=np.float32).view(np.uint32)
    mask = (u.view(np.int32) >> 31).view(np.uint32)
    return u ^ (mask | np.uint32(0x80000000))


def _image_score(im):
    im = np.asarray(im, dtype=np.uint32)
    u = np.where(im & np.uint32(0x80000000), im & np.uint32(0x7FFFFFFF), ~im)
    return u.astype(np.uint32).view(np.float32)


def _quantise(m):
    """shadow_rows_i8_kernel / prep_queries_kernel: scale = max|x| / 127, round to nearest, clamp."""
    m = np.ascontiguousarray(m, dtype=np.float32)
    scale = (np.abs(m).max(axis=1) / np.float32(127.0)).astype(np.float32)
    inv = np.where(scale > 0, np.float32(1.0) / np.where(scale > 0, scale, 1), 0).astype(np.float32)
    q8 = np.clip(np.rint((m * inv[:, None]).astype(np.float32)), -127, 127).astype(np.int32)
    res2 = (((q8.astype(np.float32) * scale[:, None]).astype(np.float32) - m).astype(np.float64) ** 2).sum(axis=1)
    return q8, scale, res2


def _grouptop(rows8, rscale, q8):
    """GROUPTOP epilogue: per 32-row group the best two keys of acc * row_scale (the query's scale comes later)."""
    acc = rows8.astype(np.int64) @ q8.astype(np.int64)
    v = (acc.astype(np.float32) * rscale).astype(np.float32)
    n = len(v)
    lane = (np.arange(n) % 32).astype(np.uint32)
    key = (_image(v) & np.uint32(0xFFFFFFE0)) | lane
    pad = (-n) % 32
    key = np.concatenate([key, np.zeros(pad, dtype=np.uint32)]).reshape(-1, 32)
    srt = np.sort(key, axis=1)
    return srt[:, -1].copy(), srt[:, -2].copy()


def _tail(rows, q, k1, k2, qscale, K, kp, delta):
    """grouptop_tail_kernel: cut, candidates, exact re-score, expansion, certificate.  -> (keys, certified)"""
    groups = len(k1)
    e = np.arange(groups)
    chunk = e // 32
    thread = (chunk % CLUSTER) * (WARPS * 32) + ((chunk // CLUSTER) % WARPS) * 32 + e % 32
    tbest = np.zeros(CLUSTER * WARPS * 32, dtype=np.uint32)
    np.maximum.at(tbest, thread, k1)
    bins = np.bincount((tbest[tbest != 0] >> K_BINS_SHIFT).astype(np.int64), minlength=4096)
    suffix = np.cumsum(bins[::-1])[::-1]
    above = np.nonzero(suffix >= kp)[0]
    if len(above) == 0 or above[-1] == 0:
        return None, False
    cutimg = np.uint32(int(above[-1]) << K_BINS_SHIFT)
    cand = np.nonzero((k1 >= cutimg) & (k1 != 0))[0]
    per_cta = np.bincount((cand // 32) % CLUSTER, minlength=CLUSTER)
    over = bool((per_cta > REGION).any())
    below = k1[k1 < cutimg]
    ncmax = np.uint32(below.max()) if len(below) else np.uint32(0)

    def reach(im):
        v = np.float32(_image_score(np.uint32(im) | np.uint32(31)) * np.float32(qscale))
        return np.float32(v + np.float32(4e-6) * np.abs(v))

    def exact(row):
        return np.float32(ol.dot(rows[row], q))

    keys = {}            # slot -> exact canonical score
    second = {}          # group -> second key (0 once the group is expanded)
    for g in cand:
        row = int(g) * 32 + int(k1[g] & 31)
        keys[row] = exact(row)
        second[int(g)] = np.uint32(k2[g])
    ok = not over
    out = None
    for rnd in range(2):
        order = sorted(keys.items(), key=lambda kv: (-float(kv[1]), kv[0]))
        # NaN scores sort last in the kernels (image 0); the model's data has none
        out = order[:K]
        if len(order) < min(K, len(rows)):
            return out, False
        tk = out[min(K, len(rows)) - 1][1]
        expand = [g for g, s in second.items() if s != 0 and not (tk > reach(s) + delta)]
        if not expand:
            break
        if rnd == 1 or len(expand) > 8 or len(keys) + 31 * len(expand) > 1024:
            ok = False
            break
        for g in expand:
            second[g] = np.uint32(0)
            for r in range(g * 32, min(g * 32 + 32, len(rows))):
                keys.setdefault(r, exact(r))
    if ok and ncmax != 0 and not (out[min(K, len(rows)) - 1][1] > reach(ncmax) + delta):
        ok = False
    return out, ok


def _search(rows, queries, K):
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    rows8, rscale, rres2 = _quantise(rows)
    q8, qscale, qres2 = _quantise(queries)
    mn = np.float32(np.sqrt(np.float32((rows.astype(np.float64) ** 2).sum(axis=1).max() * 1.0001)))
    mres = np.float32(np.sqrt(np.float32(rres2.max() * 1.0001)))
    kp = max(max(K + 96, 3 * K + 64), 6 * K + 64)
    results = []
    for i, q in enumerate(queries):
        qn = np.float32(np.sqrt(np.float32((q.astype(np.float64) ** 2).sum() * 1.0001)))
        qr = np.float32(np.sqrt(np.float32(qres2[i] * 1.0001)))
        delta = np.float32(2e-6) * mn * qn + mres * (qn + qr) * np.float32(1.0001) + mn * qr
        k1, k2 = _grouptop(rows8, rscale, q8[i])
        results.append(_tail(rows, q, k1, k2, qscale[i], K, kp, np.float32(delta)))
    return results


def _check(rows, queries, K, want_certified=None):
    exact = ol.scan_topk(rows, queries, K)
    certified = 0
    for (out, ok), (sc, idx) in zip(_search(rows, queries, K), exact):
        if not ok:
            continue
        certified += 1
        assert [r for r, _ in out] == [int(i) for i in idx], "a certified answer differs from the exact top-K"
        assert [np.float32(s) for _, s in out] == [np.float32(s) for s in sc]
    if want_certified is not None:
        assert certified >= want_certified, (certified, want_certified)
    return certified


def _unit(v):
    v = np.asarray(v, dtype=np.float32)
    return (v / np.float32(np.sqrt((v.astype(np.float64) ** 2).sum()))).astype(np.float32)


@pytest.mark.parametrize("dims,n,K", [(128, 20000, 10), (128, 9000, 1), (256, 12000, 32)])
def test_random_rows_are_certified_and_exact(dims, n, K):
    rows = ol.synth_rows(5 + dims, 0, n, dims)
    queries = ol.synth_rows(6 + dims, 0, 6, dims)
    queries[0] = rows[n // 2]
    _check(rows, queries, K, want_certified=5)


def test_rows_hidden_behind_a_group_maximum_are_found_by_expansion():
    dims, n, K = 128, 16000, 10
    rng = np.random.default_rng(3)
    rows = ol.synth_rows(31, 0, n, dims)
    queries = ol.synth_rows(32, 0, 4, dims)
    for j, r in enumerate([3203, 3204, 3217, 3231]):           # four of query 0's best rows in group 100
        rows[r] = _unit(queries[0] + rng.normal(0, 0.02 * (j + 1), dims).astype(np.float32) / np.sqrt(dims))
    for j in range(32):                                         # a whole group of good rows for query 1
        rows[8000 + j] = _unit(queries[1] + rng.normal(0, 0.05 + 0.01 * j, dims).astype(np.float32) / np.sqrt(dims))
    got = _search(rows, queries, K)
    assert got[0][1] and got[1][1], "the expansion should certify these"
    assert [r for r, _ in got[0][0]][:4] == [3203, 3204, 3217, 3231]
    _check(rows, queries, K, want_certified=4)


def test_near_duplicates_unnormalised_rows_and_tombstones_never_certify_a_wrong_answer():
    dims, n, K = 128, 16000, 10
    rng = np.random.default_rng(9)
    rows = ol.synth_rows(41, 0, n, dims)
    rows[::3] *= np.float32(2.5)
    centre = rows[1].copy()
    rows[5000:6500] = centre[None, :] + rng.normal(0, 2e-6, (1500, dims)).astype(np.float32)
    rows[100:400] = 0                                           # deleted rows: score exactly 0
    queries = ol.synth_rows(42, 0, 6, dims)
    queries[0] = _unit(centre + rng.normal(0, 1e-3, dims).astype(np.float32))
    queries[1] = -queries[1]
    queries[2] = np.float32(1e-3) * queries[2]
    certified = _check(rows, queries, K)
    assert certified <= 5                                       # the duplicates' query cannot be certified
    neg = np.abs(ol.synth_rows(43, 0, 4000, dims)) * np.float32(-1.0)
    neg[10:20] = 0                                              # every live score negative: the zero rows win
    qpos = np.abs(ol.synth_rows(44, 0, 2, dims))
    _check(neg, qpos, K)
