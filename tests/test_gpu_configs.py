"""Full-oracle parity at the exact BASELINE.json configurations ("parity gaps").

Every case compares the library's answer with a COMPLETE oracle scan of the same synthetic rows (regenerated
on the host in bounded chunks -- the rows are bit-identical to the device fill, tests/test_gpu_parity.py):
ids, order and float32 scores must match bit for bit.

  config 2   1 000 000 x 512, 10 queries top-10: the default int8 one-pass chain (prep -> GROUPTOP -> tail),
             through gf_set_search (host buffers) AND through gf_set_search_device_ex (what bench.py times)
  config 3   1 000 000 x 512, 1024 queries top-100: four tcgen05 passes of 256 queries, select + re-score
  config 4   one shard of the 100 M-row set: > 3 M rows, 64 queries top-10 -> the two-pass chain
             (threshold sample + COLLECT), slots of a rank-3-of-8 shard

Reference semantics: set.go:118-159 (top-limit over all blocks, score descending), block.go:236-243.
"""
import os

import numpy as np
import pytest

from oracle import lib as ol

pytestmark = pytest.mark.gpu

DIMS = 512
CAP = 16384                      # rows per 32 MiB block at 512-d (demo/main.go:19-23)
SEED, QSEED = 20260101, 20260102


def oracle_topk_synth(seed, n, queries, limit, first_ordinal=0, chunk=250_000, ordinals=None):
    """Exact canonical top-`limit` of `queries` over synthetic rows [first_ordinal, first_ordinal + n) (or the
    explicit, ascending `ordinals` ranges [(start, stop)]), scanned in chunks: [(scores f32, ordinals i64)]."""
    cores = os.cpu_count() or 1
    ranges = ordinals if ordinals is not None else [(first_ordinal, first_ordinal + n)]
    Q = queries.shape[0]
    best_s = [np.zeros(0, np.float32) for _ in range(Q)]
    best_i = [np.zeros(0, np.int64) for _ in range(Q)]
    pend_rows, pend_ord, pend_n = [], [], 0

    def flush():
        nonlocal pend_rows, pend_ord, pend_n
        if not pend_n:
            return
        X = np.concatenate(pend_rows) if len(pend_rows) > 1 else pend_rows[0]
        O = np.concatenate(pend_ord) if len(pend_ord) > 1 else pend_ord[0]
        for qi, (sc, idx) in enumerate(ol.scan_topk(X, queries, limit, ol.MODE_CANONICAL, cores)):
            s = np.concatenate([best_s[qi], sc.astype(np.float32)])
            o = np.concatenate([best_i[qi], O[idx]])
            order = np.lexsort((o, -s.astype(np.float64)))[:limit]   # score descending, then ordinal ascending
            best_s[qi], best_i[qi] = s[order], o[order]
        pend_rows, pend_ord, pend_n = [], [], 0

    for a, b in ranges:
        pos = a
        while pos < b:
            m = min(b - pos, chunk - pend_n)
            pend_rows.append(ol.synth_rows(seed, pos, m, DIMS))
            pend_ord.append(np.arange(pos, pos + m, dtype=np.int64))
            pend_n += m
            pos += m
            if pend_n >= chunk:
                flush()
    flush()
    return list(zip(best_s, best_i))


def bench_queries(n, seed):
    """bench.py's query recipe (U(-1,1), L2-normalised)."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(-1.0, 1.0, (n, DIMS)).astype(np.float32)
    q /= np.sqrt((q.astype(np.float64) ** 2).sum(axis=1, keepdims=True)).astype(np.float32)
    return np.ascontiguousarray(q)


@pytest.fixture(scope="module")
def set_1m(gpu_ctx):
    from gofeature_b200 import api
    cache = api.NewCache(gpu_ctx, 62, CAP * DIMS * 4)
    assert cache.ShadowBytesPerElement() == 1          # the default: int8 shadow slab
    cache.NewSet("cfg", DIMS, 4, 1024)
    st = cache.GetSet("cfg")
    st.AddSynthetic(1_000_000, SEED)
    yield st
    cache.Close()


def _check_host(got, want, prefix="f"):
    assert len(got) == len(want)
    for g, (sc, ordn) in zip(got, want):
        assert [r.ID for r in g] == ["%s%011d" % (prefix, o) for o in ordn]
        assert [np.float32(r.Score) for r in g] == list(sc)


def test_config2_full_oracle_host_call(set_1m, gpu_ctx):
    """1 M x 512, 10 queries top-10 through gf_set_search: the default path must be the int8 one-pass chain and
    the answer the full-scan oracle's, bit for bit -- three batches of bench.py's own query pool."""
    q = bench_queries(30, QSEED)
    want = oracle_topk_synth(SEED, 1_000_000, q, 10)
    gpu_ctx.ResetStats()
    got = []
    for b in range(3):
        got += set_1m.SearchArray(-2.0, 10, q[b * 10:(b + 1) * 10])
    st = gpu_ctx.Stats()
    assert st["tensor_passes"] == 3 and st["gemm_launches"] == 3 and st["scan_launches"] == 0, st   # one-pass, no fallback
    _check_host(got, want)


def test_config2_full_oracle_device_call(set_1m, gpu_ctx):
    """The call bench.py's `value` times: gf_set_search_device_ex with the queries in HBM.  Flags must be clear
    and the keys (score image, slot) equal to the oracle's; and the exact scan (path 1) returns the same keys."""
    import torch
    from gofeature_b200 import api
    q = bench_queries(30, QSEED)[10:20]
    want = oracle_topk_synth(SEED, 1_000_000, q, 10)
    dq = torch.from_numpy(q).cuda()
    keys = torch.zeros((10, 10), dtype=torch.int64, device="cuda")
    flags = torch.ones((10,), dtype=torch.int32, device="cuda")
    for path in (0, 1):
        keys.zero_()
        set_1m.SearchDeviceEx(10, 10, dq.data_ptr(), keys.data_ptr(), flags.data_ptr(), path, None)
        torch.cuda.synchronize()
        assert not flags.cpu().numpy().any()
        hk = keys.cpu().numpy().view(np.uint64)
        for i, (sc, ordn) in enumerate(want):
            assert [api.key_slot(k) for k in hk[i]] == list(ordn)
            assert [np.float32(api.key_score(k)) for k in hk[i]] == list(sc)
    flagged, unrepaired = set_1m.DeviceCounters()
    assert unrepaired == 0


def test_config3_full_oracle_1024x100(set_1m, gpu_ctx):
    """1 M x 512, 1024 queries top-100 (BASELINE configs[2]): four 256-query tcgen05 passes + select + re-score +
    certificate; every one of the 102 400 hits must equal the full-scan oracle's."""
    q = bench_queries(1024, QSEED + 7)
    want = oracle_topk_synth(SEED, 1_000_000, q, 100)
    gpu_ctx.ResetStats()
    got = set_1m.SearchArray(-2.0, 100, q)
    st = gpu_ctx.Stats()
    assert st["tensor_passes"] == 4, st
    _check_host(got, want)
    assert st["uncertified_queries"] <= 8, st          # (uniform data: the certificate holds for ~every query)


def test_config4_shard_full_oracle_two_pass(gpu_ctx):
    """Config 4's per-GPU shape: rank 3 of an 8-way sharded set holding 3.28 M rows (> 3 M: the two-pass chain),
    64 queries top-10.  The shard owns global blocks 3, 11, 19, ...; slots are global."""
    import torch
    from gofeature_b200 import api
    rank, world, blocks = 3, 8, 200
    n_local = blocks * CAP + 5000                      # a ragged last block
    cache = api.NewCache(gpu_ctx, blocks + 2, CAP * DIMS * 4)
    cache.NewSet("shard", DIMS, 4, 64)
    st = cache.GetSet("shard")
    st.SetShard(rank, world)
    ranges = []
    for lb in range(blocks + 1):
        g = lb * world + rank
        m = CAP if lb < blocks else 5000
        st.AddSynthetic(m, SEED, g * CAP)              # global block g holds ordinals [g*CAP, g*CAP + m)
        ranges.append((g * CAP, g * CAP + m))
    assert st.ScannedRows() == n_local
    q = bench_queries(64, QSEED + 11)
    q[5] = ol.synth_rows(SEED, ranges[77][0] + 123, 1, DIMS)[0]      # a needle this shard owns
    want = oracle_topk_synth(SEED, 0, q, 10, ordinals=ranges)
    gpu_ctx.ResetStats()
    dq = torch.from_numpy(q).cuda()
    keys = torch.zeros((64, 10), dtype=torch.int64, device="cuda")
    flags = torch.ones((64,), dtype=torch.int32, device="cuda")
    st.SearchDeviceEx(10, 64, dq.data_ptr(), keys.data_ptr(), flags.data_ptr(), 0, None)
    torch.cuda.synchronize()
    s = gpu_ctx.Stats()
    assert s["tensor_passes"] == 1 and s["gemm_launches"] == 2, s       # TILEMAX sample + COLLECT: the two-pass chain
    assert not flags.cpu().numpy().any()
    hk = keys.cpu().numpy().view(np.uint64)
    for i, (sc, ordn) in enumerate(want):
        assert [api.key_slot(k) for k in hk[i]] == list(ordn)           # global slot == ordinal
        assert [np.float32(api.key_score(k)) for k in hk[i]] == list(sc)
    assert api.key_slot(hk[5][0]) == ranges[77][0] + 123 and abs(float(api.key_score(hk[5][0])) - 1.0) < 1e-6
    cache.Close()
