"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle.

Bar: feature IDs exact (ties broken by slot), scores BIT-exact against the oracle's
canonical float32 order, and within 1e-5 absolute (the tolerance north_star states) of the
float64 ground truth and of a plain sequential float32 loop.
"""
import json
import os

import numpy as np
import pytest

from oracle import MODE_CANONICAL, MODE_F64, MODE_SEQ32
from oracle import lib as ol
from oracle import model

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5  # north_star: scores within 1e-5 absolute


def _pair(gpu_ctx, name, dims, batch, block_rows, block_num, shadow=None):
    """The same empty set twice: on the GPU and in the oracle's model."""
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


def _check_modes(oset, got, threshold, limit, queries):
    """scores within 1e-5 of float64 truth and of the sequential float32 loop; ids equal unless a near-tie."""
    for mode in (MODE_F64, MODE_SEQ32):
        ref = oset.search(threshold, limit, queries, mode)
        for g, w in zip(got, ref):
            n = min(len(g), len(w))
            gs = np.array([float(r.Score) for r in g[:n]])
            ws = np.array([s for s, _ in w[:n]])
            assert np.all(np.abs(gs - ws) <= TOL)
            for k in range(n):
                if g[k].ID != w[k][1]:  # allowed only among candidates closer than the float32 noise
                    alt = [s for s, i in w if i == g[k].ID]
                    assert alt and abs(alt[0] - w[k][0]) < 4e-6, (g[k], w[k])


# ------------------------------------------------------------------ the reference's own test
def test_basic_search_golden_through_c_abi(gpu_ctx):
    """TestBasicSearch K1-K5 (search_test.go:77-145), same calls, same assertions."""
    from gofeature_b200 import api
    with open(os.path.join(GOLD, "basic_search.json")) as f:
        basic = json.load(f)
    cache = api.NewCache(gpu_ctx, basic["cache"]["blockNum"], basic["cache"]["blockSize"])
    s = basic["set"]
    cache.NewSet(s["name"], s["dims"], s["precision"], s["batch"])
    st = cache.GetSet(s["name"])
    v = {k: api.NoramlizeFloat32(x) for k, x in basic["raw"].items()}
    f = {k: api.Feature(api.TFeatureValue(v[k]), k) for k in v}
    st.Add(f["f1"], f["f2"])
    oc = model.Cache(basic["cache"]["blockNum"], basic["cache"]["blockSize"])
    oc.new_set(s["name"], s["dims"], s["precision"], s["batch"])
    ost = oc.get_set(s["name"])
    ost.add(np.stack([v["f1"], v["f2"]]), ["f1", "f2"])
    for k in basic["pinned"]:
        if "delete" in k:
            assert st.Delete(*k["delete"]) == k["deleted"]
            ost.delete(k["delete"])
        call, exp = k["call"], k["expect"]
        ret = st.Search(call["threshold"], call["limit"], *[f[n].Value for n in call["queries"]])
        assert len(ret) == exp["batch"], k["id"]
        assert [len(r) for r in ret] == exp["counts"], k["id"]
        assert [r[0].ID for r in ret] == exp["first_ids"], k["id"]
        for r in ret:
            if "first_score_min" in exp:
                assert r[0].Score >= np.float32(exp["first_score_min"])
            if "first_score_max" in exp:
                assert r[0].Score <= np.float32(exp["first_score_max"])
        _assert_same(ret, ost.search(call["threshold"], call["limit"], [v[n] for n in call["queries"]]), k["id"])
    cache.Close()


# ------------------------------------------------------------------ random normalised data
@pytest.mark.parametrize("dims,n,block_rows,q,limit", [
    (512, 10000, 10000, 1, 1),      # BASELINE config 1: BenchmarkSearch shape, test geometry
    (512, 40000, 16384, 1, 10),     # demo geometry, 3 blocks, ragged last block
    (512, 20000, 4096, 3, 10),
    (512, 20000, 4096, 8, 100),
    (512, 9000, 2048, 13, 10),      # more queries than one scan pass holds
    (512, 3000, 1024, 2, 1024),     # widest single-pass K
    (128, 30000, 8192, 4, 10),
    (256, 20000, 4096, 8, 16),
    (1024, 8000, 2048, 4, 10),
    (1024, 6000, 2048, 7, 5),       # 1024-d takes 4 queries per pass
    (5, 1000, 100, 2, 3),           # generic kernel (the reference's unit-test dims)
    (100, 5000, 1000, 5, 10),
    (130, 4000, 999, 1, 7),
    (516, 3000, 700, 2, 10),        # dims % 4 == 0 but not a bulk-copy geometry
])
def test_random_parity(gpu_ctx, dims, n, block_rows, q, limit):
    cache, st, ost = _pair(gpu_ctx, "rnd", dims, max(q, 1), block_rows, (n + block_rows - 1) // block_rows + 1)
    rows = ol.synth_rows(1000 + dims, 0, n, dims)
    queries = ol.synth_rows(2000 + dims, 0, q, dims)
    ids = ["id%07d" % i for i in range(n)]
    st.AddArray(rows, ids)
    ost.add(rows, ids)
    assert st.ScannedRows() == n and st.BlockCount() == len(ost.blocks)
    got = st.SearchArray(-1.0, limit, queries)
    _assert_same(got, ost.search(-1.0, limit, queries), "canonical")
    _check_modes(ost, got, -1.0, limit, queries)
    thr = float(np.float32(got[0][min(2, len(got[0]) - 1)].Score))     # a threshold that cuts the list
    _assert_same(st.SearchArray(thr, limit, queries), ost.search(thr, limit, queries), "threshold")
    cache.Close()


def test_limit_beyond_rows_and_band_passes(gpu_ctx):
    """limit > rows returns everything (util.go:88 `i < len(scores)`); limit > 1024 runs in bands."""
    cache, st, ost = _pair(gpu_ctx, "bands", 512, 2, 2000, 4)
    rows = ol.synth_rows(7, 0, 5000, 512)
    queries = ol.synth_rows(8, 0, 2, 512)
    ids = [str(i) for i in range(5000)]
    st.AddArray(rows, ids)
    ost.add(rows, ids)
    for limit in (1025, 3000, 5000, 9999):
        got = st.SearchArray(-2.0, limit, queries)
        assert [len(g) for g in got] == [min(limit, 5000)] * 2
        _assert_same(got, ost.search(-2.0, limit, queries), "limit %d" % limit)
    cache.Close()


def test_empty_and_ragged(gpu_ctx):
    from gofeature_b200 import api
    cache, st, ost = _pair(gpu_ctx, "empty", 512, 4, 100, 3)
    q = ol.synth_rows(1, 0, 2, 512)
    got = st.SearchArray(0.0, 5, q)                                   # no block yet: len(ret) == batch
    assert [len(g) for g in got] == [0, 0]
    assert st.Search(0.0, 5) == []                                    # zero queries
    st.AddArray(ol.synth_rows(2, 0, 1, 512), ["only"])                # a single row
    ost.add(ol.synth_rows(2, 0, 1, 512), ["only"])
    _assert_same(st.SearchArray(-1, 5, q), ost.search(-1, 5, q))
    assert [len(g) for g in st.SearchArray(-1, 0, q)] == [0, 0]       # limit 0
    rows = ol.synth_rows(3, 0, 150, 512)                              # spills into a second, ragged block
    st.AddArray(rows, ["r%d" % i for i in range(150)])
    ost.add(rows, ["r%d" % i for i in range(150)])
    _assert_same(st.SearchArray(-1, 20, q), ost.search(-1, 20, q))
    with pytest.raises(api.GoFeatureError) as e:
        st.SearchArray(0, 1, ol.synth_rows(1, 0, 5, 512))             # set.go:119-122
    assert e.value.name == "ErrOutOfBatch"
    with pytest.raises(api.GoFeatureError) as e:
        st.Search(0, 1, b"\0" * 100)
    assert e.value.name == "ErrMismatchDimension"
    with pytest.raises(api.GoFeatureError) as e:
        st.AddArray(ol.synth_rows(4, 0, 1000, 512), [str(i) for i in range(1000)])
    assert e.value.name == "ErrNotEnoughBlocks"                       # cache.go:111-113
    _assert_same(st.SearchArray(-1, 20, q), ost.search(-1, 20, q))   # failed Add changed nothing
    cache.Close()


# ------------------------------------------------------------------ ties and exact arithmetic
def test_exact_arithmetic_ties_and_duplicates(gpu_ctx):
    """Components k/64: every sum is exact, so duplicates tie exactly and must come out in slot order."""
    cache, st, ost = _pair(gpu_ctx, "ties", 512, 4, 3000, 5)
    rng = np.random.default_rng(9)
    base = (rng.integers(-4, 5, (40, 512)) / 64.0).astype(np.float32)
    rows = base[rng.integers(0, 40, 10000)]                           # 10 000 rows, 40 distinct values
    queries = (rng.integers(-4, 5, (4, 512)) / 64.0).astype(np.float32)
    ids = ["d%05d" % i for i in range(10000)]
    st.AddArray(rows, ids)
    ost.add(rows, ids)
    for limit in (1, 10, 300, 2000):
        got = st.SearchArray(-100.0, limit, queries)
        _assert_same(got, ost.search(-100.0, limit, queries), "limit %d" % limit)
        _assert_same(got, ost.search(-100.0, limit, queries, MODE_SEQ32), "seq32 limit %d" % limit)
        for g in got:                                                  # ties in ascending slot order
            for a, b in zip(g, g[1:]):
                assert a.Score > b.Score or (a.Score == b.Score and a.Slot < b.Slot)
    cache.Close()


def test_special_values(gpu_ctx):
    """NaN sorts last, +-0 tie on the slot, infinities order correctly."""
    cache, st, ost = _pair(gpu_ctx, "special", 4, 1, 64, 2)
    rows = np.array([[0, 0, 0, 0], [1, 0, 0, 0], [np.nan, 0, 0, 0], [-1, 0, 0, 0], [np.inf, 0, 0, 0],
                     [-0.0, 0, 0, 0], [-np.inf, 0, 0, 0], [1, 0, 0, 0]], dtype=np.float32)
    ids = list("abcdefgh")
    st.AddArray(rows, ids)
    ost.add(rows, ids)
    q = np.array([[1, 0, 0, 0]], dtype=np.float32)
    got = st.SearchArray(float("-inf"), 8, q)
    want = ost.search(float("-inf"), 8, q)
    assert [r.ID for r in got[0]] == [i for _, i in want[0]] == ["e", "b", "h", "a", "f", "d", "g"]
    cache.Close()


# ------------------------------------------------------------------ Add / Delete / tombstones
def test_add_delete_refill_sequence(gpu_ctx):
    """set.go:77-116 / 163-193 and block.go:80-165 bookkeeping under a random mutation sequence."""
    cache, st, ost = _pair(gpu_ctx, "mut", 512, 4, 700, 12)
    rng = np.random.default_rng(21)
    queries = ol.synth_rows(99, 0, 4, 512)
    next_id, alive = 0, []
    for step in range(14):
        n = int(rng.integers(1, 900))
        rows = ol.synth_rows(500 + step, 0, n, 512)
        ids = ["m%06d" % (next_id + i) for i in range(n)]
        next_id += n
        st.AddArray(rows, ids)
        ost.add(rows, ids)
        alive += ids
        kill = [alive[i] for i in sorted(set(rng.integers(0, len(alive), int(rng.integers(0, 200))).tolist()))]
        kill += ["missing-%d" % step]
        got_d = st.Delete(*kill)
        want_d = ost.delete(kill)
        assert sorted(got_d or []) == sorted(want_d)       # order inside a call is a Go map walk (block.go:150)
        alive = [a for a in alive if a not in set(want_d)]
        assert st.LiveRows() == len(alive)
        assert st.ScannedRows() == sum(b.next_index for b in ost.blocks)
        for limit in (1, 10, 50):
            _assert_same(st.SearchArray(-1.0, limit, queries), ost.search(-1.0, limit, queries), "step %d" % step)
    # deleted ids are gone, survivors are found with score ~1
    probe = ost.blocks[0].rows[[i for i, x in enumerate(ost.blocks[0].ids) if x][0]]
    r = st.SearchArray(0.99, 1, probe)
    assert len(r[0]) == 1 and r[0][0].Score > 0.9999
    cache.Close()


def test_tombstone_selection_quirk(gpu_ctx):
    """block.go:236-243: tombstones (score exactly 0.0) take places in the per-block top-limit and
    are dropped afterwards -- pinned by search_test.go:139-144 (K5)."""
    cache, st, ost = _pair(gpu_ctx, "quirk", 8, 2, 50, 6)
    rng = np.random.default_rng(4)
    rows = rng.uniform(-1, 1, (230, 8)).astype(np.float32)
    ids = ["t%03d" % i for i in range(230)]
    st.AddArray(rows, ids)
    ost.add(rows, ids)
    kill = [ids[i] for i in range(0, 230, 3)]
    assert st.Delete(*kill) == ost.delete(kill)
    q = rng.uniform(-1, 1, (2, 8)).astype(np.float32)
    before = gpu_ctx.Stats()["quirk_fallbacks"]
    for limit in (1, 5, 30, 60, 500):
        for thr in (-10.0, 0.0, 0.2):
            _assert_same(st.SearchArray(thr, limit, q), ost.search(thr, limit, q), "limit %d thr %g" % (limit, thr))
    assert gpu_ctx.Stats()["quirk_fallbacks"] > before      # the exact per-block path really ran
    cache.Close()


def test_tombstone_quirk_large_blocks(gpu_ctx):
    """Same, with 512-d bulk-copy geometry and more than 1024 winners per block."""
    cache, st, ost = _pair(gpu_ctx, "quirk2", 512, 1, 1500, 4)
    rows = ol.synth_rows(31, 0, 4000, 512)
    ids = ["u%04d" % i for i in range(4000)]
    st.AddArray(rows, ids)
    ost.add(rows, ids)
    kill = ids[::2]
    assert st.Delete(*kill) == ost.delete(kill)
    q = ol.synth_rows(32, 0, 1, 512)
    for limit in (10, 900, 1200, 4000):
        _assert_same(st.SearchArray(-1.0, limit, q), ost.search(-1.0, limit, q), "limit %d" % limit)
    cache.Close()


def test_duplicate_ids_and_delete_semantics(gpu_ctx):
    """block.go:140-144: the LAST slot holding an id inside the FIRST block that has it is removed."""
    cache, st, ost = _pair(gpu_ctx, "dup", 4, 1, 3, 4)
    rows = np.eye(4, dtype=np.float32)[[0, 1, 2, 3, 0, 1]]
    ids = ["x", "x", "y", "x", "z", "x"]
    st.AddArray(rows, ids)
    ost.add(rows, ids)
    q = np.eye(4, dtype=np.float32)
    for _ in range(4):
        assert (st.Delete("x", "x") or []) == ost.delete(["x", "x"])
        for i in range(4):
            _assert_same(st.SearchArray(0.5, 6, q[i]), ost.search(0.5, 6, q[i:i + 1]))
    from gofeature_b200 import api
    with pytest.raises(api.GoFeatureError) as e:
        st.Delete("")
    assert e.value.name == "ErrInvalidFeautres"
    cache.Close()


def test_update_and_read(gpu_ctx):
    """interface.go:47-55 (empty TODO stubs in the reference): in-place overwrite and read-back."""
    from gofeature_b200 import api
    cache, st, ost = _pair(gpu_ctx, "upd", 512, 1, 500, 3)
    rows = ol.synth_rows(41, 0, 900, 512)
    ids = ["k%03d" % i for i in range(900)]
    st.AddArray(rows, ids)
    got = st.Read("k007", "nope", "k899")
    assert [f.ID for f in got] == ["k007", "k899"]
    assert got[0].Value == rows[7].tobytes() and got[1].Value == rows[899].tobytes()
    new = ol.synth_rows(42, 0, 1, 512)
    assert st.Update(api.Feature(new.tobytes(), "k500"), api.Feature(new.tobytes(), "zzz")) == ["k500"]
    r = st.SearchArray(0.999, 2, new)
    assert [x.ID for x in r[0]] == ["k500"]
    assert st.Read("k500")[0].Value == new.tobytes()
    cache.Close()


# ------------------------------------------------------------------ Cache / Block / Buffer API
def test_cache_and_set_lifecycle(gpu_ctx):
    from gofeature_b200 import api
    cache = api.NewCache(gpu_ctx, 5, 100 * 64 * 4)
    assert cache.GetBlockSize() == 100 * 64 * 4
    cache.NewSet("a", 64, 4, 2)
    with pytest.raises(api.GoFeatureError) as e:
        cache.NewSet("a", 64, 4, 2)
    assert e.value.name == "ErrFeatureSetExist"                        # cache.go:47-50
    with pytest.raises(api.GoFeatureError) as e:
        cache.GetSet("b")
    assert e.value.name == "ErrFeatureSetNotFound"
    a = cache.GetSet("a")
    a.AddArray(ol.synth_rows(1, 0, 250, 64), [str(i) for i in range(250)])
    assert a.BlockCount() == 3
    assert [b.IsOwned() for b in cache.AllBlocks()] == [True, True, True, False, False]
    assert len(cache.GetEmptyBlock(2)) == 2
    with pytest.raises(api.GoFeatureError) as e:
        cache.GetEmptyBlock(3)
    assert e.value.name == "ErrNotEnoughBlocks"
    cache.DestroySet("a")                                              # set.go:202-214 releases the blocks
    assert not any(b.IsOwned() for b in cache.AllBlocks())
    with pytest.raises(api.GoFeatureError) as e:
        cache.DestroySet("a")
    assert e.value.name == "ErrFeatureSetNotFound"
    cache.NewSet("a", 64, 4, 2)                                        # blocks were zeroed and are reusable
    a = cache.GetSet("a")
    rows = ol.synth_rows(2, 0, 10, 64)
    a.AddArray(rows, [str(i) for i in range(10)])
    r = a.SearchArray(-1, 20, rows[3])
    assert len(r[0]) == 10 and r[0][0].ID == "3"
    cache.Close()


def test_block_level_api(gpu_ctx):
    """Block.Accquire / Insert / Delete / Search (block.go:59-249) with the dims-major query
    buffer FeatureValueTranspose1D produces (util.go:21-39, set.go:57-63)."""
    from gofeature_b200 import api
    dims, cap = 512, 300
    cache = api.NewCache(gpu_ctx, 2, cap * dims * 4)
    blk = cache.GetEmptyBlock(1)[0]
    assert not blk.IsOwned() and blk.Capacity() == 0
    blk.Accquire(None, "owner", dims, 4, 3)
    assert blk.IsOwned() and blk.Capacity() == cap and blk.Margin() == cap
    with pytest.raises(api.GoFeatureError) as e:
        blk.Accquire(None, "other", dims, 4, 3)
    assert e.value.name == "ErrBlockUsed"                              # block.go:60-62
    rows = ol.synth_rows(51, 0, 200, dims)
    feats = [api.Feature(rows[i].tobytes(), "b%03d" % i) for i in range(200)]
    inp = api.NewGPUBuffer(gpu_ctx, None, 3 * dims * 4)
    out = api.NewGPUBuffer(gpu_ctx, None, 3 * cap * 4)
    assert blk.Search(inp, out, 3, 5) is None                         # NextIndex == 0 (block.go:186-189)
    blk.Insert(*feats)
    assert blk.Margin() == cap - 200 and blk.NextIndex() == 200
    with pytest.raises(api.GoFeatureError) as e:
        blk.Insert(*([feats[0]] * 101))
    assert e.value.name == "ErrBlockIsFull"                            # block.go:81-83
    assert blk.Delete("b010", "nope", "b011") == ["b010", "b011"]
    q = ol.synth_rows(52, 0, 3, dims)
    inp.Write(api.FeatureValueTranspose1D(4, *[q[i].tobytes() for i in range(3)]))
    got = blk.Search(inp, out, 3, 7)
    live = np.ones(200, dtype=np.uint8)
    live[[10, 11]] = 0
    zr = rows.copy()
    zr[[10, 11]] = 0
    want = ol.set_search([zr], [200], [live], q, 7, float("-inf"), MODE_CANONICAL)
    for g, w in zip(got, want):
        assert [r.ID for r in g] == ["b%03d" % row for _, _, row in w]
        assert [float(r.Score) for r in g] == [sc for sc, _, _ in w]
    blk.Release()
    assert not blk.IsOwned() and blk.NextIndex() == 0
    cache.Close()


def test_buffer_api(gpu_ctx):
    """buffer.go:15-85."""
    from gofeature_b200 import api
    b = api.NewGPUBuffer(gpu_ctx, None, 64)
    assert b.Size() == 64 and b.Read() == b"\0" * 64                  # allocated zeroed (buffer.go:24)
    b.Write(bytes(range(32)))
    assert b.Read()[:32] == bytes(range(32))
    with pytest.raises(api.GoFeatureError) as e:
        b.Write(b"x" * 65)
    assert e.value.name == "ErrBufferWriteOutofRange"                  # buffer.go:35-37
    s = b.Slice(8, 16)
    assert s.Size() == 8 and s.Read() == bytes(range(8, 16))          # aliases the parent (buffer.go:81-84)
    s.Reset()
    assert b.Read()[8:16] == b"\0" * 8
    for bad in ((-1, 4), (0, 65), (10, 5), (65, 65)):
        with pytest.raises(api.GoFeatureError) as e:
            b.Slice(*bad)
        assert e.value.name == "ErrBufferSliceOutofRange"              # buffer.go:74-79
    c = api.NewGPUBuffer(gpu_ctx, None, 128)
    c.Copy(b)
    assert c.Read()[:64] == b.Read()
    with pytest.raises(api.GoFeatureError) as e:
        s.Copy(b)
    assert e.value.name == "ErrBufferCopyOutofRange"
    for x in (s, b, c):
        x.Free()


# ------------------------------------------------------------------ synthetic fill + device path
def test_synthetic_fill_is_bit_identical_to_the_oracle(gpu_ctx):
    cache, st, ost = _pair(gpu_ctx, "synth", 512, 2, 1000, 4)
    st.AddSynthetic(2500, 777, first_ordinal=100)
    rows = ol.synth_rows(777, 100, 2500, 512)
    ids = ["f%011d" % (100 + i) for i in range(2500)]
    ost.add(rows, ids)
    got = st.Read(ids[0], ids[1234], ids[2499])
    assert [g.Value for g in got] == [rows[0].tobytes(), rows[1234].tobytes(), rows[2499].tobytes()]
    q = ol.synth_rows(778, 0, 2, 512)
    _assert_same(st.SearchArray(-1, 10, q), ost.search(-1, 10, q))
    assert st.Delete(ids[5], "f00000000000", ids[2000]) == ost.delete([ids[5], "f00000000000", ids[2000]])
    extra = ol.synth_rows(779, 0, 3, 512)                              # refill -> ids become explicit
    st.AddArray(extra, ["e0", "e1", "e2"])
    ost.add(extra, ["e0", "e1", "e2"])
    _assert_same(st.SearchArray(-1, 10, q), ost.search(-1, 10, q))
    _assert_same(st.SearchArray(0.99, 3, extra[1]), ost.search(0.99, 3, extra[1:2]))
    cache.Close()


def test_device_resident_search_matches_host_search(gpu_ctx):
    """gf_set_search_device (queries and keys stay in HBM) gives the keys of gf_set_search."""
    import torch
    from gofeature_b200 import api
    cache, st, ost = _pair(gpu_ctx, "dev", 512, 8, 4096, 6)
    st.AddSynthetic(20000, 5)
    q = ol.synth_rows(6, 0, 8, 512)
    dq = torch.from_numpy(q).cuda()
    keys = torch.zeros((8, 10), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    st.SearchDevice(10, 8, dq.data_ptr(), keys.data_ptr(), 0, None)
    gpu_ctx.Synchronize()
    hk = keys.cpu().numpy().view(np.uint64)
    host = st.SearchArray(-1, 10, q)
    for i in range(8):
        assert [api.key_slot(k) for k in hk[i]] == [r.Slot for r in host[i]]
        assert [float(api.key_score(k)) for k in hk[i]] == [float(r.Score) for r in host[i]]
    res = st.ResolveKeys(-1, 10, hk)
    assert [[r.ID for r in row] for row in res] == [[r.ID for r in row] for row in host]
    cache.Close()


# ------------------------------------------------------------------ full-size properties (BASELINE config 2)
def test_full_size_needles_1m(gpu_ctx):
    """1 000 000 x 512 (2.048 GB): stored rows used as queries must come back first with score ~1,
    results are sorted, and a sampled row range agrees with the oracle bit for bit."""
    from gofeature_b200 import api
    n, dims = 1_000_000, 512
    cache = api.NewCache(gpu_ctx, 62, 1024 * 1024 * 32)                # demo geometry (demo/main.go:19-20)
    cache.NewSet("full", dims, 4, 8)
    st = cache.GetSet("full")
    st.AddSynthetic(n, 20260101)
    assert st.ScannedRows() == n and st.BlockCount() == 62
    needles = [0, 16383, 16384, 500_000, 999_999]
    qrows = np.stack([ol.synth_rows(20260101, r, 1, dims)[0] for r in needles])
    got = st.SearchArray(0.0, 10, qrows)
    for r, g in zip(needles, got):
        assert g[0].ID == "f%011d" % r and abs(float(g[0].Score) - 1.0) < 1e-6
        sc = [float(x.Score) for x in g]
        assert sc == sorted(sc, reverse=True) and len(g) == 10
    # the winners of a random query, re-scored by the oracle from regenerated rows
    q = ol.synth_rows(20260102, 0, 1, dims)
    g = st.SearchArray(-1.0, 10, q)[0]
    for x in g:
        row = ol.synth_rows(20260101, int(x.ID[1:]), 1, dims)[0]
        assert np.float32(ol.dot(row, q[0], MODE_CANONICAL)) == x.Score
    # a full oracle scan of the first 200 000 rows must not beat the 10th winner unless it is a winner
    part = ol.synth_rows(20260101, 0, 200_000, dims)
    sc, idx = ol.scan_topk(part, q, 10, MODE_CANONICAL, 8)[0]
    tenth = float(g[-1].Score)
    ids = {x.ID for x in g}
    for s_, i_ in zip(sc, idx):
        assert s_ <= tenth or ("f%011d" % i_) in ids
    cache.Close()


# ------------------------------------------------------------------ large batches: tcgen05 candidate pass
@pytest.mark.parametrize("dims,n,q,limit", [
    (512, 70000, 9, 10),        # smallest batch the tensor path takes
    (512, 70000, 33, 1),
    (512, 90000, 100, 100),     # K' = 150 re-scored per query
    (512, 70000, 300, 10),      # two passes of <= 256 queries
    (128, 140000, 20, 10),
    (256, 80000, 17, 50),
])
def test_tensor_path_parity(gpu_ctx, dims, n, q, limit):
    """q >= 9 on a large set runs TF32 tensor-core candidate selection + canonical re-score + certificate;
    the answer must still be bit-identical to the oracle (and therefore to the fused scan)."""
    block_rows = 16384
    cache, st, ost = _pair(gpu_ctx, "tensor", dims, q, block_rows, (n + block_rows - 1) // block_rows + 1,
                           shadow=False)                               # the bf16 shadow path: tests/test_gpu_shadow.py
    st.AddSynthetic(n, 4242 + dims)
    rows = ol.synth_rows(4242 + dims, 0, n, dims)
    ost.add(rows, ["f%011d" % i for i in range(n)])
    queries = ol.synth_rows(99 + q, 0, q, dims)
    queries[0] = rows[n // 2]                                          # a needle: score 1 at a known slot
    gpu_ctx.ResetStats()
    got = st.SearchArray(-1.0, limit, queries)
    stats = gpu_ctx.Stats()
    assert stats["tensor_passes"] == (q + 255) // 256, stats           # the tcgen05 path really ran
    assert got[0][0].ID == "f%011d" % (n // 2)
    _assert_same(got, ost.search(-1.0, limit, queries), "tensor")
    assert stats["uncertified_queries"] == 0
    kill = ["f%011d" % i for i in range(0, n, 997)]
    assert sorted(st.Delete(*kill)) == sorted(ost.delete(kill))        # tombstones compete as zero rows
    _assert_same(st.SearchArray(0.05, limit, queries), ost.search(0.05, limit, queries), "tensor+tombstones")
    cache.Close()


def test_tensor_path_certificate_fallback(gpu_ctx):
    """Thousands of rows whose scores differ by less than the TF32 error bound: the certificate cannot
    hold, the flagged queries are re-run by the exact scan (on the device, then by the host call), and
    the answer is still exact.  Un-normalised rows exercise the |row|.|query| scaling of the bound."""
    dims, n, q, limit = 512, 70000, 24, 10
    cache, st, ost = _pair(gpu_ctx, "cert", dims, q, 16384, 6, shadow=False)
    rng = np.random.default_rng(17)
    rows = ol.synth_rows(808, 0, n, dims)
    rows[::3] *= np.float32(2.5)                                       # norms up to 2.5
    centre = rows[1].copy()
    near = centre[None, :] + rng.normal(0, 2e-6, (6000, dims)).astype(np.float32)
    rows[10000:16000] = near                                           # 6000 near-duplicates
    ids = ["c%06d" % i for i in range(n)]
    st.AddArray(rows, ids)
    ost.add(rows, ids)
    queries = ol.synth_rows(809, 0, q, dims)
    for i in range(12):                                                # half the batch looks at the cluster
        queries[i] = centre / np.linalg.norm(centre) + rng.normal(0, 1e-3, dims).astype(np.float32)
    gpu_ctx.ResetStats()
    got = st.SearchArray(-10.0, limit, queries)
    stats = gpu_ctx.Stats()
    assert stats["tensor_passes"] == 1
    assert stats["uncertified_queries"] > 0 or stats["scan_launches"] >= 1
    _assert_same(got, ost.search(-10.0, limit, queries), "certificate fallback")
    cache.Close()


# ------------------------------------------------------------------ compaction (not in the reference)
def test_compact_reclaims_tombstones(gpu_ctx):
    """gf_set_compact packs live rows per block and shrinks NextIndex: the scan gets shorter, live rows keep
    their order, Delete / Add / Search keep agreeing with the oracle's model of the same operation."""
    cache, st, ost = _pair(gpu_ctx, "compact", 512, 4, 1000, 8)
    rows = ol.synth_rows(71, 0, 4500, 512)
    ids = ["z%05d" % i for i in range(4500)]
    st.AddArray(rows, ids)
    ost.add(rows, ids)
    rng = np.random.default_rng(2)
    kill = [ids[i] for i in sorted(set(rng.integers(0, 4500, 1800).tolist()))]
    kill += ids[2000:3000]                                              # one block loses every row
    assert sorted(st.Delete(*kill)) == sorted(ost.delete(kill))
    q = ol.synth_rows(72, 0, 4, 512)
    before = st.SearchArray(0.0, 20, q)
    assert st.ScannedRows() == 4500
    freed = st.Compact()
    assert freed == ost.compact() == 4500 - st.LiveRows()
    assert st.ScannedRows() == st.LiveRows() == sum(b.next_index for b in ost.blocks)
    after = st.SearchArray(0.0, 20, q)
    _assert_same(after, ost.search(0.0, 20, q), "after compact")
    assert [[r.ID for r in row] for row in after] == [[r.ID for r in row] for row in before]   # positive scores: same hits
    assert st.Compact() == 0                                            # idempotent
    more = ol.synth_rows(73, 0, 1200, 512)
    mids = ["y%05d" % i for i in range(1200)]
    st.AddArray(more, mids)                                             # fills the reclaimed tails first
    ost.add(more, mids)
    assert st.BlockCount() == len(ost.blocks)
    assert sorted(st.Delete(ids[1], mids[7], "nope") or []) == sorted(ost.delete([ids[1], mids[7], "nope"]))
    _assert_same(st.SearchArray(-1.0, 50, q), ost.search(-1.0, 50, q), "after refill")
    got = st.Read(mids[0], ids[4499] if ids[4499] not in kill else mids[1])
    assert got and got[0].Value == more[0].tobytes()
    cache.Close()


def test_export_and_reimport_round_trip(gpu_ctx):
    """gf_set_export: live rows + ids in slot order; a set rebuilt from them answers alike."""
    cache, st, ost = _pair(gpu_ctx, "export", 256, 3, 600, 12)
    rows = ol.synth_rows(81, 0, 2000, 256)
    ids = ["e%04d" % i for i in range(2000)]
    st.AddArray(rows, ids)
    st.AddSynthetic(500, 82, first_ordinal=7)                          # implicit ids export too
    kill = ids[5:900:7] + ["f%011d" % 9]
    st.Delete(*kill)
    vals, out_ids = st.Export()
    dead = set(kill)
    want_ids = [i for i in ids if i not in dead] + ["f%011d" % (7 + i) for i in range(500) if 7 + i != 9]
    assert out_ids == want_ids and vals.shape == (len(want_ids), 256) and st.LiveRows() == len(want_ids)
    assert vals[0].tobytes() == rows[0].tobytes()
    assert vals[len(want_ids) - 1].tobytes() == ol.synth_rows(82, 7 + 499, 1, 256)[0].tobytes()
    cache.NewSet("reimport", 256, 4, 3)
    st2 = cache.GetSet("reimport")
    st2.AddArray(vals, out_ids)
    q = ol.synth_rows(83, 0, 3, 256)
    a, b = st.SearchArray(0.0, 15, q), st2.SearchArray(0.0, 15, q)
    assert [[(r.ID, r.Score) for r in row] for row in a] == [[(r.ID, r.Score) for r in row] for row in b]
    cache.Close()


# ------------------------------------------------------------------ random operation sequences
@pytest.mark.parametrize("dims,block_rows,seed", [(8, 37, 1), (128, 300, 2), (512, 256, 3), (100, 64, 4)])
def test_random_operation_sequences(gpu_ctx, dims, block_rows, seed):
    """Interleaved Add / Delete / Update / Compact / Search with random sizes, thresholds and limits; after
    every step the CUDA path and the oracle's model must agree on everything observable."""
    from gofeature_b200 import api
    cache, st, ost = _pair(gpu_ctx, "seq%d" % seed, dims, 6, block_rows, 40)
    rng = np.random.default_rng(seed)
    alive, nid = [], 0
    for step in range(40):
        op = rng.choice(["add", "add", "delete", "update", "compact", "search", "search"])
        if op == "add" or not alive:
            n = int(rng.integers(1, 3 * block_rows))
            if len(alive) + n > 30 * block_rows:
                continue
            rows = ol.synth_rows(1000 * seed + step, 0, n, dims, normalize=bool(rng.integers(0, 2)))
            ids = ["i%06d" % (nid + i) for i in range(n)]
            nid += n
            st.AddArray(rows, ids)
            ost.add(rows, ids)
            alive += ids
        elif op == "delete":
            k = int(rng.integers(1, max(2, len(alive) // 3)))
            kill = [alive[i] for i in sorted(set(rng.integers(0, len(alive), k).tolist()))] + ["ghost%d" % step]
            assert sorted(st.Delete(*kill) or []) == sorted(ost.delete(kill))
            dead = set(kill)
            alive = [a for a in alive if a not in dead]
        elif op == "update":
            pick = [alive[int(i)] for i in rng.integers(0, len(alive), 3)] + ["ghost"]
            new = ol.synth_rows(77 * seed + step, 0, len(pick), dims)
            got = st.Update(*[api.Feature(new[i].tobytes(), pick[i]) for i in range(len(pick))])
            assert set(got or []) == set(pick[:-1])
            for i, ident in enumerate(pick[:-1]):      # the model has no Update (a TODO stub in the reference):
                for b in ost.blocks:                   # overwrite the row the id lives in
                    if ident in b.ids:
                        b.rows[len(b.ids) - 1 - b.ids[::-1].index(ident)] = new[i]
                        break
        elif op == "compact":
            assert st.Compact() == ost.compact()
        assert st.LiveRows() == len(alive) and st.ScannedRows() == sum(b.next_index for b in ost.blocks)
        q = ol.synth_rows(5000 * seed + step, 0, int(rng.integers(1, 7)), dims)
        limit = int(rng.choice([1, 3, 10, 40, 200, 1500]))
        thr = float(rng.choice([-10.0, -0.1, 0.0, 0.05]))
        _assert_same(st.SearchArray(thr, limit, q), ost.search(thr, limit, q), "step %d %s" % (step, op))
    vals, out_ids = st.Export()
    assert out_ids == [i for b in ost.blocks for i in b.ids[:b.next_index] if i != ""]
    cache.Close()


def test_implicit_id_blocks_take_arbitrary_ids(gpu_ctx):
    """Synthetic (implicit-id) blocks whose tombstones are refilled by Add keep their generated ids and record
    the arbitrary ones in an overlay: Search / Delete / Update / Read / Export / Compact must behave exactly
    like the oracle's model, which only knows explicit ids."""
    from gofeature_b200 import api
    dims, block_rows, n = 64, 400, 2000
    cache, st, ost = _pair(gpu_ctx, "overlay", dims, 4, block_rows, 12)
    st.AddSynthetic(n, 91)
    syn = ol.synth_rows(91, 0, n, dims)
    ost.add(syn, ["f%011d" % i for i in range(n)])
    q = ol.synth_rows(92, 0, 4, dims)

    def same(what, limit=20, thr=-1.0):
        _assert_same(st.SearchArray(thr, limit, q), ost.search(thr, limit, q), what)
        assert st.LiveRows() == sum(sum(1 for i in b.ids[:b.next_index] if i != "") for b in ost.blocks)

    kill = ["f%011d" % i for i in (3, 4, 5, 401, 799, 800, 1999, 1200)]
    assert sorted(st.Delete(*kill)) == sorted(ost.delete(kill))
    extra = ol.synth_rows(93, 0, 30, dims)
    eids = ["x%02d" % i for i in range(30)]
    st.AddArray(extra, eids)                       # 8 refills spread over implicit blocks + a tail append
    ost.add(extra, eids)
    same("refill into implicit blocks")
    # the generated name of a refilled slot is gone; the arbitrary id that replaced it is there
    assert (st.Delete("f%011d" % 3, "x00", "f%011d" % 10) or []) == ost.delete(["f%011d" % 3, "x00", "f%011d" % 10])
    same("delete overlay + implicit ids")
    got = st.Read("x01", "f%011d" % 11, "f%011d" % 4)
    assert [f.ID for f in got] == ["x01", "f%011d" % 11]
    new = ol.synth_rows(94, 0, 1, dims)[0]
    assert st.Update(api.Feature(new.tobytes(), "x02")) == ["x02"]
    for b in ost.blocks:
        if "x02" in b.ids:
            b.rows[b.ids.index("x02")] = new
    same("update of an overlay row")
    vals, out_ids = st.Export()
    assert out_ids == [i for b in ost.blocks for i in b.ids[:b.next_index] if i != ""]
    # enough arbitrary ids into ONE implicit block to make it explicit (overlay > capacity / 4)
    victims = ["f%011d" % i for i in range(400, 400 + 150)]
    assert sorted(st.Delete(*victims)) == sorted(ost.delete(victims))
    more = ol.synth_rows(95, 0, 150, dims)
    mids = ["y%03d" % i for i in range(150)]
    st.AddArray(more, mids)
    ost.add(more, mids)
    same("block turned explicit")
    assert sorted(st.Delete("y000", "f%011d" % 560, "x05")) == sorted(ost.delete(["y000", "f%011d" % 560, "x05"]))
    assert st.Compact() == ost.compact()
    same("after compact", limit=50)
    st.AddSynthetic(100, 96, first_ordinal=10 ** 6)
    ost.add(ol.synth_rows(96, 10 ** 6, 100, dims), ["f%011d" % (10 ** 6 + i) for i in range(100)])
    same("synthetic append after all that")
    vals, out_ids = st.Export()
    assert out_ids == [i for b in ost.blocks for i in b.ids[:b.next_index] if i != ""]
    cache.Close()
