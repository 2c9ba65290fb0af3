"""The certificate on NON-uniform data: clusters of near neighbours and exact duplicates, which is
what a face-feature index holds.  Rows: 100 members per centre at cosine ~0.975 to it (~0.95 to each other),
scattered over the set, 1 % exact copies of earlier rows (gf_set_add_synthetic_ex, bit-identical to the oracle's
generator); queries: near a centre, so ~100 rows score within 0.01 of the winner -- inside the int8 pass's error.

Bar: ids / order / float32 scores equal a full oracle scan bit for bit, and (the point of the test) the tensor
path certifies them itself: at most 1 % of the queries may fall back to the exact scan."""
import numpy as np
import pytest

from oracle import lib as ol

pytestmark = pytest.mark.gpu

DIMS, CAP = 512, 16384
SEED, CENTRES, SPREAD, DUP = 4242, 3000, 0.228, 10000


def _near_centre_queries(n, seed, noise=0.1):
    rng = np.random.default_rng(seed)
    out = np.empty((n, DIMS), dtype=np.float32)
    cs = rng.integers(0, CENTRES, n)
    for i, c in enumerate(cs):
        v = ol.synth_centre(SEED, int(c), DIMS) + noise * rng.uniform(-1, 1, DIMS).astype(np.float32)
        out[i] = v / np.linalg.norm(v)
    return out


def test_clustered_fill_is_bit_identical_to_the_oracle(gpu_ctx):
    from gofeature_b200 import api
    cache = api.NewCache(gpu_ctx, 3, CAP * DIMS * 4)
    cache.NewSet("fill", DIMS, 4, 8)
    st = cache.GetSet("fill")
    st.AddSynthetic(20000, SEED, 1000, True, "f", CENTRES, SPREAD, DUP)
    vals, ids = st.Export()
    want = ol.synth_rows_ex(SEED, 1000, 20000, DIMS, CENTRES, SPREAD, DUP)
    assert vals.tobytes() == want.tobytes()
    assert ids[0] == "f%011d" % 1000 and ids[-1] == "f%011d" % 20999
    cache.Close()


@pytest.mark.parametrize("batch,limit", [(8, 10), (32, 10), (10, 32), (64, 10)])
def test_clustered_queries_certify_and_match_the_oracle(gpu_ctx, batch, limit):
    from gofeature_b200 import api
    n = 300_000
    cache = api.NewCache(gpu_ctx, n // CAP + 2, CAP * DIMS * 4)
    cache.NewSet("clu", DIMS, 4, 64)
    st = cache.GetSet("clu")
    st.AddSynthetic(n, SEED, 0, True, "f", CENTRES, SPREAD, DUP)
    rows = ol.synth_rows_ex(SEED, 0, n, DIMS, CENTRES, SPREAD, DUP)
    nq = max(64, 2 * batch)
    q = _near_centre_queries(nq, 9)
    import os
    want = ol.scan_topk(rows, q, limit, ol.MODE_CANONICAL, os.cpu_count() or 1)
    # the workload really is clustered: the 10th best score is within 0.02 of the best
    assert np.median([w[0][0] - w[0][9] for w in want]) < 0.02
    gpu_ctx.ResetStats()
    got = []
    for b in range(0, nq, batch):
        got += st.SearchArray(-2.0, limit, q[b:b + batch])
    s = gpu_ctx.Stats()
    assert s["tensor_passes"] >= nq // batch
    for g, (sc, idx) in zip(got, want):
        assert [r.ID for r in g] == ["f%011d" % i for i in idx]
        assert [np.float32(r.Score) for r in g] == [np.float32(x) for x in sc]
    assert s["uncertified_queries"] <= max(1, nq // 100), s      # >= 99 % certified without the exact scan
    cache.Close()


def test_refused_batch_takes_the_tf32_tier_before_the_exact_scan(gpu_ctx):
    """More than one scan pass worth (> 8) of queries the int8 pass cannot certify: 4 800 rows spread evenly within 0.03
    of the winner (160 rows per 0.001), so the 124 groups the one-pass chain re-scores end 0.0008 below the winner --
    far inside the int8 bound (~0.008).  The second tier -- ONE kind::tf32 pass over the float32 rows for the compact
    batch, bound 2.2e-3, every collected row re-scored -- certifies them; answers stay bit-identical to the oracle and
    the exact scan is not launched."""
    from gofeature_b200 import api
    dims, cap = 128, 16384
    rng = np.random.default_rng(77)
    n_bg, n_near = 70_000, 4_800
    base = rng.uniform(-1, 1, dims).astype(np.float32)
    base /= np.linalg.norm(base)
    bg = rng.uniform(-1, 1, (n_bg, dims)).astype(np.float32)
    bg /= np.linalg.norm(bg, axis=1, keepdims=True)
    # near rows: cos(row, base) = 1 - t, t spread evenly over [0, 0.03]
    t = np.linspace(0.0, 0.03, n_near).astype(np.float64)
    noise = rng.uniform(-1, 1, (n_near, dims))
    noise -= np.outer(noise @ base.astype(np.float64), base.astype(np.float64))
    noise /= np.linalg.norm(noise, axis=1, keepdims=True)
    c = 1.0 - t
    near = (c[:, None] * base.astype(np.float64)[None, :] + np.sqrt(1.0 - c * c)[:, None] * noise).astype(np.float32)
    rows = np.concatenate([bg, near])
    rows = rows[rng.permutation(len(rows))]
    cache = api.NewCache(gpu_ctx, len(rows) // cap + 2, cap * dims * 4)
    cache.NewSet("near", dims, 4, 16)
    st = cache.GetSet("near")
    st.AddArray(rows, ["r%06d" % i for i in range(len(rows))])
    nq, limit = 12, 10
    q = base[None, :] + 1e-4 * rng.uniform(-1, 1, (nq, dims)).astype(np.float32)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    import os
    want = ol.scan_topk(rows, q, limit, ol.MODE_CANONICAL, os.cpu_count() or 1)
    gpu_ctx.ResetStats()
    got = st.SearchArray(-2.0, limit, q)
    s = gpu_ctx.Stats()
    for g, (sc, idx) in zip(got, want):
        assert [r.ID for r in g] == ["r%06d" % i for i in idx]
        assert [np.float32(r.Score) for r in g] == [np.float32(x) for x in sc]
    assert s["uncertified_queries"] == nq, s          # the int8 pass refused all of them ...
    assert s["tier2_queries"] == nq, s                # ... the tf32 tier certified all of them ...
    assert s["scan_launches"] == 0, s                 # ... and the exact scan never ran
    cache.Close()
