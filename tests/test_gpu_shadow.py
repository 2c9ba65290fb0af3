"""GPU parity of the bf16-shadow tensor path (kind::f16 candidate pass over the half-size shadow slab +
canonical float32 re-score + certificate): every result must be BIT-identical to the oracle, whatever the
batch size, and stay so while Add / Delete / refill / Update / Compact keep rewriting rows, because the
shadow must follow every device-side write of the float32 rows.
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


@pytest.mark.parametrize("fmt", ["int8", "bf16"])
@pytest.mark.parametrize("dims,n,q,limit", [
    (512, 70000, 1, 10),        # a single query: the shadow path takes every batch size
    (512, 70000, 3, 1),
    (512, 70000, 8, 100),
    (512, 90000, 40, 10),
    (512, 70000, 300, 100),     # two passes of <= 256 queries, K' selection
    (128, 140000, 5, 10),
    (256, 80000, 17, 50),
    (1024, 66000, 2, 10),
    (192, 100000, 12, 10),      # 64 | dims but not 128: bf16 shadow pass, or TF32 over the float32 rows (int8 cache)
])
def test_shadow_path_parity(gpu_ctx, dims, n, q, limit, fmt):
    block_rows = 16384
    cache, st, ost = _pair(gpu_ctx, "shadow", dims, q, block_rows, (n + block_rows - 1) // block_rows + 1, shadow=fmt)
    assert cache.ShadowBytesPerElement() == {"int8": 1, "bf16": 2}[fmt]
    st.AddSynthetic(n, 777 + dims)
    rows = ol.synth_rows(777 + dims, 0, n, dims)
    ost.add(rows, ["f%011d" % i for i in range(n)])
    queries = ol.synth_rows(31 + q, 0, q, dims)
    queries[0] = rows[n // 3]
    gpu_ctx.ResetStats()
    got = st.SearchArray(-1.0, limit, queries)
    stats = gpu_ctx.Stats()
    assert stats["tensor_passes"] == (q + 255) // 256, stats
    assert stats["uncertified_queries"] == 0, stats
    assert got[0][0].ID == "f%011d" % (n // 3)
    want = ost.search(-1.0, limit, queries)
    _assert_same(got, want, "shadow")
    # the same call again replays the recorded CUDA graph
    _assert_same(st.SearchArray(-1.0, limit, queries), want, "shadow (graph replay)")
    # every other path gives the same bytes
    for path in (1, 3):
        st.SetSearchPath(path)
        _assert_same(st.SearchArray(-1.0, limit, queries), want, "path %d" % path)
    st.SetSearchPath(0)
    kill = ["f%011d" % i for i in range(0, n, 991)]
    assert sorted(st.Delete(*kill)) == sorted(ost.delete(kill))        # tombstones: zero rows in both slabs
    _assert_same(st.SearchArray(0.05, limit, queries), ost.search(0.05, limit, queries), "shadow+tombstones")
    cache.Close()


def test_shadow_off_and_required(gpu_ctx):
    """GF_CACHE_NO_SHADOW keeps the float32-only behaviour (tensor path from q = 9, fused scan below)."""
    n, dims = 70000, 512
    cache, st, ost = _pair(gpu_ctx, "noshadow", dims, 16, 16384, 6, shadow=False)
    assert not cache.HasShadow()
    st.AddSynthetic(n, 5)
    ost.add(ol.synth_rows(5, 0, n, dims), ["f%011d" % i for i in range(n)])
    q = ol.synth_rows(6, 0, 12, dims)
    gpu_ctx.ResetStats()
    _assert_same(st.SearchArray(-1.0, 10, q[:2]), ost.search(-1.0, 10, q[:2]))
    assert gpu_ctx.Stats()["tensor_passes"] == 0
    _assert_same(st.SearchArray(-1.0, 10, q), ost.search(-1.0, 10, q))
    assert gpu_ctx.Stats()["tensor_passes"] == 1
    cache.Close()


@pytest.mark.parametrize("fmt", ["int8", "bf16"])
def test_shadow_follows_every_mutation(gpu_ctx, fmt):
    """Explicit Add (tail + tombstone refill), Delete, Update and Compact on a set large enough for the
    tensor path: after each step the bf16 shadow must describe the float32 rows, or winners go missing."""
    from gofeature_b200 import api
    dims, block_rows = 512, 8192
    cache, st, ost = _pair(gpu_ctx, "mut", dims, 8, block_rows, 14, shadow=fmt)
    rng = np.random.default_rng(5)
    n0 = 70000
    rows = ol.synth_rows(900, 0, n0, dims)
    ids = ["m%06d" % i for i in range(n0)]
    st.AddArray(rows, ids)
    ost.add(rows, ids)
    queries = ol.synth_rows(901, 0, 6, dims)

    def check(what, limit=10, thr=-1.0, qs=queries):
        gpu_ctx.ResetStats()
        got = st.SearchArray(thr, limit, qs)
        assert gpu_ctx.Stats()["tensor_passes"] >= 1, what
        _assert_same(got, ost.search(thr, limit, qs), what)

    check("after add")
    # delete the current winners of query 0: their shadow rows must stop scoring
    top = [r.ID for r in st.SearchArray(-1.0, 10, queries[:1])[0]]
    kill = top[:5] + [ids[i] for i in sorted(set(rng.integers(0, n0, 3000).tolist()))]
    assert sorted(st.Delete(*kill)) == sorted(ost.delete(kill))
    check("after delete")
    # refill tombstones (scattered slots, > 8 and <= 8 per block) and grow the tail
    near = (queries[1][None, :] * np.float32(0.9) + rng.normal(0, 0.01, (20, dims))).astype(np.float32)
    more = np.concatenate([near, ol.synth_rows(902, 0, 5000, dims)])
    mids = ["n%06d" % i for i in range(len(more))]
    st.AddArray(more, mids)
    ost.add(more, mids)
    check("after refill + tail")                                        # the 20 near rows must win query 1
    got = st.SearchArray(-1.0, 10, queries[1:2])[0]
    assert all(r.ID.startswith("n") for r in got)
    few = (queries[2][None, :] * np.float32(0.8) + rng.normal(0, 0.01, (3, dims))).astype(np.float32)
    st.Delete(mids[100], mids[200], mids[300])
    ost.delete([mids[100], mids[200], mids[300]])
    st.AddArray(few, ["p0", "p1", "p2"])                                # <= 8 refills: per-slot conversion
    ost.add(few, ["p0", "p1", "p2"])
    check("after small refill")
    assert {r.ID for r in st.SearchArray(-1.0, 3, queries[2:3])[0]} == {"p0", "p1", "p2"}
    # Update in place: the row becomes the best hit of query 3
    new = (queries[3] * np.float32(0.95)).astype(np.float32)
    target = ids[12345] if ids[12345] not in kill else ids[12346]
    assert st.Update(api.Feature(new.tobytes(), target)) == [target]
    for b in ost.blocks:
        if target in b.ids:
            b.rows[b.ids.index(target)] = new
            break
    check("after update")
    assert st.SearchArray(-1.0, 1, queries[3:4])[0][0].ID == target
    # Compact moves rows inside blocks
    assert st.Compact() == ost.compact()
    check("after compact", limit=50, thr=0.0)
    st.AddSynthetic(3000, 903, first_ordinal=10 ** 6)
    ost.add(ol.synth_rows(903, 10 ** 6, 3000, dims), ["f%011d" % (10 ** 6 + i) for i in range(3000)])
    check("after synthetic append", limit=100)
    cache.Close()


@pytest.mark.parametrize("fmt", ["int8", "bf16"])
def test_shadow_certificate_fallback_and_extreme_rows(gpu_ctx, fmt):
    """Near-duplicates closer than the bf16 error, un-normalised rows, and rows whose elements leave the bf16
    range (the residual bound becomes infinite): the certificate must refuse and the exact scan must answer."""
    dims, n, q, limit = 512, 70000, 12, 10
    cache, st, ost = _pair(gpu_ctx, "scert", dims, q, 16384, 6, shadow=fmt)
    rng = np.random.default_rng(23)
    rows = ol.synth_rows(818, 0, n, dims)
    rows[::3] *= np.float32(2.5)
    centre = rows[1].copy()
    rows[10000:16000] = centre[None, :] + rng.normal(0, 2e-6, (6000, dims)).astype(np.float32)
    ids = ["c%06d" % i for i in range(n)]
    st.AddArray(rows, ids)
    ost.add(rows, ids)
    queries = ol.synth_rows(819, 0, q, dims)
    for i in range(6):
        queries[i] = centre / np.linalg.norm(centre) + rng.normal(0, 1e-3, dims).astype(np.float32)
    gpu_ctx.ResetStats()
    got = st.SearchArray(-10.0, limit, queries)
    stats = gpu_ctx.Stats()
    assert stats["tensor_passes"] == 1
    assert stats["uncertified_queries"] > 0 or stats["scan_launches"] >= 1
    _assert_same(got, ost.search(-10.0, limit, queries), "bf16 certificate fallback")
    # one row far outside the bf16 range: 3.4e38 rounds to +Inf in bf16
    huge = np.zeros((1, dims), dtype=np.float32)
    huge[0, 0] = np.float32(3.4e38)
    huge[0, 1] = np.float32(1.0)
    st.AddArray(huge, ["huge"])
    ost.add(huge, ["huge"])
    small_q = (queries * np.float32(1e-3)).astype(np.float32)
    _assert_same(st.SearchArray(-10.0, limit, small_q), ost.search(-10.0, limit, small_q), "huge row")
    cache.Close()


def test_shadow_cache_with_dims_not_multiple_of_64(gpu_ctx):
    """dims = 96 cannot use 64-element bf16 K blocks: the set keeps the TF32 pass over the float32 rows."""
    dims, n = 96, 140000
    cache, st, ost = _pair(gpu_ctx, "d96", dims, 16, 16384, 10)
    st.AddSynthetic(n, 11)
    ost.add(ol.synth_rows(11, 0, n, dims), ["f%011d" % i for i in range(n)])
    q = ol.synth_rows(12, 0, 10, dims)
    gpu_ctx.ResetStats()
    _assert_same(st.SearchArray(-1.0, 10, q), ost.search(-1.0, 10, q))
    assert gpu_ctx.Stats()["tensor_passes"] == 1
    _assert_same(st.SearchArray(-1.0, 10, q[:1]), ost.search(-1.0, 10, q[:1]))     # q = 1: fused scan
    cache.Close()


def test_coalesced_callers_fast_handoff_and_tombstone_retry(gpu_ctx):
    """Concurrent gf_set_search callers folded into one tensor-path pass: the leader publishes raw keys and
    every caller resolves its own ids under the leader's read lock (fast hand-off); a caller whose winners
    contain a tombstone re-runs alone through the per-block path (block.go:236-243).  Every caller must get
    exactly what an uncoalesced Search (= the oracle) returns, with its own threshold."""
    import threading
    dims, n, nthreads, rounds = 512, 70000, 6, 12
    cache, st, ost = _pair(gpu_ctx, "co", dims, 8, 16384, 6)
    rows = ol.synth_rows(1234, 0, n, dims).copy()
    rows[:, 0] = np.abs(rows[:, 0]) + np.float32(0.05)                 # every row has a positive first component
    ids = ["k%06d" % i for i in range(n)]
    st.AddArray(rows, ids)
    ost.add(rows, ids)
    kill = [ids[i] for i in range(7, n, 1999)]
    assert sorted(st.Delete(*kill)) == sorted(ost.delete(kill))
    queries = ol.synth_rows(1235, 0, nthreads, dims).copy()
    queries[0, :] = 0
    queries[0, 0] = -1.0                                               # all live scores negative: tombstones (0.0) win
    thresholds = [-10.0, 0.0, 0.12, -1.0, 0.15, -0.3]
    limits = [10, 10, 10, 10, 10, 10]
    want = [ost.search(thresholds[t], limits[t], queries[t:t + 1]) for t in range(nthreads)]
    st.SetCoalescing(200)
    gpu_ctx.ResetStats()
    errors = []

    def worker(t):
        try:
            for _ in range(rounds):
                got = st.SearchArray(thresholds[t], limits[t], queries[t:t + 1])
                _assert_same(got, want[t], "thread %d" % t)
        except Exception as exc:  # noqa: BLE001
            errors.append((t, repr(exc)))

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert not errors, errors
    stats = gpu_ctx.Stats()
    assert stats["coalesced_batches"] > 0 and stats["tensor_passes"] > 0, stats
    assert stats["quirk_fallbacks"] >= rounds, stats                   # thread 0 took the per-block path every time
    assert want[0][0] and all(s < 0 for s, _ in want[0][0])            # and still got the live (negative) rows
    st.SetCoalescing(0)
    cache.Close()


def test_device_search_flags(gpu_ctx):
    """gf_set_search_device_ex: more uncertifiable queries than one device-side fix-up pass takes (8) leave
    flags behind; every unflagged query is exact, flagged ones are exact after a path=1 re-run."""
    import torch
    from gofeature_b200 import api
    dims, n, q, limit = 512, 70000, 24, 10
    cache, st, ost = _pair(gpu_ctx, "flags", dims, q, 16384, 6)
    rng = np.random.default_rng(29)
    rows = ol.synth_rows(828, 0, n, dims)
    centre = rows[1].copy()
    rows[10000:16000] = centre[None, :] + rng.normal(0, 2e-6, (6000, dims)).astype(np.float32)
    ids = ["c%06d" % i for i in range(n)]
    st.AddArray(rows, ids)
    ost.add(rows, ids)
    queries = ol.synth_rows(829, 0, q, dims)
    for i in range(16):                                                # 16 queries cannot be certified
        queries[i] = centre / np.linalg.norm(centre) + rng.normal(0, 1e-3, dims).astype(np.float32)
    dq = torch.from_numpy(queries).cuda()
    keys = torch.zeros((q, limit), dtype=torch.int64, device="cuda")
    flags = torch.zeros((q,), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    L = api.lib()
    api._check(L.gf_set_search_device_ex(st._h, limit, q, dq.data_ptr(), keys.data_ptr(), flags.data_ptr(), 0, stream))
    torch.cuda.synchronize()
    hf = flags.cpu().numpy()
    assert 9 <= int(hf.sum()) <= 16, hf                                # a caller that takes the flags repairs them itself
    want = ost.search(-10.0, limit, queries)
    hk = keys.cpu().numpy().view(np.uint64)
    for i in range(q):
        if hf[i]:
            continue
        assert [float(api.key_score(k)) for k in hk[i]] == [float(np.float32(s_)) for s_, _ in want[i]], i
    # without a flags buffer up to 8 uncertified queries per call are repaired on the device (fix-up scan)
    few = queries.copy()
    few[6:16] = ol.synth_rows(830, 0, 10, dims)
    dq2 = torch.from_numpy(few).cuda()
    st.SearchDevice(limit, q, dq2.data_ptr(), keys.data_ptr(), 0, stream)
    torch.cuda.synchronize()
    _assert_same(st.ResolveKeys(-10.0, limit, keys.cpu().numpy().view(np.uint64)), ost.search(-10.0, limit, few),
                 "device fix-up")
    api._check(L.gf_set_search_device_ex(st._h, limit, q, dq.data_ptr(), keys.data_ptr(), flags.data_ptr(), 1, stream))
    torch.cuda.synchronize()
    assert int(flags.cpu().numpy().sum()) == 0
    res = st.ResolveKeys(-10.0, limit, keys.cpu().numpy().view(np.uint64))
    _assert_same(res, want, "path 1 re-run")
    cache.Close()
