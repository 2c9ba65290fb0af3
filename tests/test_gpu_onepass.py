"""GPU parity of the ONE-PASS chain small batches take (prep -> GROUPTOP pass -> tail; gf_gemm.cu): the tcgen05 pass
emits the best two approximate keys of every 32-row group plus a histogram, the tail kernel cuts the histogram at the
K'-th best group, re-scores the candidates in the canonical float32 order, re-scores whole groups whose second key
could still matter, and certifies.  Every result must be BIT-identical to the oracle (IDs, order, scores), the same
as through the two-pass chain (threshold sample + COLLECT) and the fused exact scan.
"""
import numpy as np
import pytest

from oracle import lib as ol
from oracle import model

pytestmark = pytest.mark.gpu


def _pair(gpu_ctx, name, dims, batch, block_rows, block_num, shadow="int8"):
    from gofeature_b200 import api
    bs = block_rows * dims * 4
    cache = api.NewCache(gpu_ctx, block_num, bs, shadow=shadow)
    cache.NewSet(name, dims, 4, batch)
    oc = model.Cache(block_num, bs)
    oc.new_set(name, dims, 4, batch)
    return cache, cache.GetSet(name), oc.get_set(name)


def _assert_same(got, want, what=""):
    assert len(got) == len(want), what
    for q, (g, w) in enumerate(zip(got, want)):
        assert [r.ID for r in g] == [i for _, i in w], "%s query %d ids" % (what, q)
        gs = np.array([r.Score for r in g], dtype=np.float32)
        ws = np.array([s for s, _ in w], dtype=np.float32)
        assert gs.tobytes() == ws.tobytes(), "%s query %d scores" % (what, q)


def _unit(v):
    v = np.asarray(v, dtype=np.float32)
    return (v / np.float32(np.sqrt((v.astype(np.float64) ** 2).sum()))).astype(np.float32)


@pytest.mark.parametrize("fmt", ["int8", "bf16", False])
@pytest.mark.parametrize("dims,n,q,limit", [
    (512, 70000, 1, 10),
    (512, 70017, 10, 10),       # a ragged last tile (1 live row) and a partial last block
    (512, 66000, 16, 1),
    (512, 90077, 17, 32),       # 17 queries pad to 32 columns
    (512, 70000, 32, 32),
    (128, 140000, 5, 10),
    (1024, 66000, 2, 32),
    (128, 1200000, 10, 10),     # more than 32 768 groups per query: the tail's threads take several rounds of entries
])
def test_onepass_parity(gpu_ctx, dims, n, q, limit, fmt):
    if n > 1000000 and fmt != "int8":
        pytest.skip("the large geometry runs once, on the default shadow")
    block_rows = 16384
    cache, st, ost = _pair(gpu_ctx, "onepass", dims, q, block_rows, (n + block_rows - 1) // block_rows + 1, shadow=fmt)
    st.AddSynthetic(n, 4242 + dims)
    rows = ol.synth_rows(4242 + dims, 0, n, dims)
    ost.add(rows, ["f%011d" % i for i in range(n)])
    queries = ol.synth_rows(77 + q, 0, q, dims)
    queries[0] = rows[n - 1]           # the only live row of the ragged tile
    want = ost.search(-1.0, limit, queries)
    st.SetSearchPath(2)
    gpu_ctx.ResetStats()
    got = st.SearchArray(-1.0, limit, queries)
    stats = gpu_ctx.Stats()
    assert stats["tensor_passes"] == 1 and stats["gemm_launches"] == 1, stats   # ONE big pass, no sample launch
    assert stats["uncertified_queries"] == 0 and stats["scan_launches"] == 0, stats
    assert got[0][0].ID == "f%011d" % (n - 1)
    _assert_same(got, want, "one-pass")
    got = st.SearchArray(-1.0, limit, queries)   # replayed from the recorded graph
    _assert_same(got, want, "one-pass replay")
    st.SetSearchPath(4)                          # the two-pass chain on the same set: sample + COLLECT
    gpu_ctx.ResetStats()
    got = st.SearchArray(-1.0, limit, queries)
    assert gpu_ctx.Stats()["gemm_launches"] == 2
    _assert_same(got, want, "two-pass")
    cache.Close()


@pytest.mark.parametrize("fmt", ["int8", "bf16"])
def test_onepass_rows_hidden_behind_a_group_maximum(gpu_ctx, fmt):
    """Several of the best rows inside ONE 32-row group: the pass only reports the group's best two keys, so the
    tail must re-score the whole group -- without falling back to the exact scan."""
    dims, n, limit = 512, 70000, 10
    cache, st, ost = _pair(gpu_ctx, "hidden", dims, 16, 16384, 6, shadow=fmt)
    rng = np.random.default_rng(5)
    rows = ol.synth_rows(99, 0, n, dims)
    queries = ol.synth_rows(100, 0, 6, dims)
    # query 0: its 4 nearest rows all sit in group 1000 (rows 32000..32031); query 1: 2 in group 7 + 3 in group 2100
    for j, r in enumerate([32003, 32004, 32017, 32031]):
        rows[r] = _unit(queries[0] + rng.normal(0, 0.02 * (j + 1), dims).astype(np.float32) / np.sqrt(dims))
    for j, r in enumerate([7 * 32 + 0, 7 * 32 + 31, 2100 * 32 + 5, 2100 * 32 + 6, 2100 * 32 + 7]):
        rows[r] = _unit(queries[1] + rng.normal(0, 0.03 * (j + 1), dims).astype(np.float32) / np.sqrt(dims))
    # query 2: a whole group of good rows (more than limit of them in one group)
    for j in range(32):
        rows[50016 + j] = _unit(queries[2] + rng.normal(0, 0.05 + 0.01 * j, dims).astype(np.float32) / np.sqrt(dims))
    ids = ["h%06d" % i for i in range(n)]
    st.AddArray(rows, ids)
    ost.add(rows, ids)
    st.SetSearchPath(2)
    gpu_ctx.ResetStats()
    got = st.SearchArray(-1.0, limit, queries)
    stats = gpu_ctx.Stats()
    assert stats["gemm_launches"] == 1 and stats["uncertified_queries"] == 0 and stats["scan_launches"] == 0, stats
    want = ost.search(-1.0, limit, queries)
    assert [i for _, i in want[0]][:4] == ["h032003", "h032004", "h032017", "h032031"]
    _assert_same(got, want, "hidden rows")
    cache.Close()


def test_onepass_follows_mutations_and_tombstones(gpu_ctx):
    dims, n, limit = 512, 80000, 10
    cache, st, ost = _pair(gpu_ctx, "mut", dims, 16, 16384, 8)
    rows = ol.synth_rows(7, 0, n, dims)
    ids = ["m%06d" % i for i in range(n)]
    st.AddArray(rows, ids)
    ost.add(rows, ids)
    queries = ol.synth_rows(8, 0, 10, dims)
    queries[3] = rows[40000]
    queries[4] = rows[123]
    st.SetSearchPath(2)
    _assert_same(st.SearchArray(-1.0, limit, queries), ost.search(-1.0, limit, queries), "before")
    dead = [ids[40000], ids[123]] + ids[1000:1400]
    assert sorted(st.Delete(*dead)) == sorted(ost.delete(dead))
    _assert_same(st.SearchArray(-1.0, limit, queries), ost.search(-1.0, limit, queries), "after delete")
    more = ol.synth_rows(9, 0, 700, dims)
    more[5] = _unit(queries[3] * np.float32(0.999) + more[5] * np.float32(0.03))
    mids = ["n%06d" % i for i in range(700)]
    st.AddArray(more, mids)
    ost.add(more, mids)
    _assert_same(st.SearchArray(-1.0, limit, queries), ost.search(-1.0, limit, queries), "after refill")
    _assert_same(st.SearchArray(0.05, 3, queries), ost.search(0.05, 3, queries), "threshold")
    cache.Close()


def test_onepass_certificate_refuses_near_duplicates(gpu_ctx):
    """Hundreds of rows closer to each other than the int8 error: no cut separates K of them from the rest, the tail
    flags the query and the exact scan answers -- same bits as the oracle."""
    dims, n, limit = 512, 70000, 10
    cache, st, ost = _pair(gpu_ctx, "dups", dims, 16, 16384, 6)
    rng = np.random.default_rng(11)
    rows = ol.synth_rows(21, 0, n, dims)
    centre = rows[5].copy()
    rows[20000:23000] = centre[None, :] + rng.normal(0, 2e-6, (3000, dims)).astype(np.float32)
    ids = ["d%06d" % i for i in range(n)]
    st.AddArray(rows, ids)
    ost.add(rows, ids)
    queries = ol.synth_rows(22, 0, 4, dims)
    queries[1] = _unit(centre + rng.normal(0, 1e-3, dims).astype(np.float32))
    st.SetSearchPath(2)
    gpu_ctx.ResetStats()
    got = st.SearchArray(-1.0, limit, queries)
    stats = gpu_ctx.Stats()
    assert stats["uncertified_queries"] >= 1 and stats["scan_launches"] >= 1, stats
    _assert_same(got, ost.search(-1.0, limit, queries), "near duplicates")
    cache.Close()
