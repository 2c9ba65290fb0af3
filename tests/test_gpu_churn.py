"""BASELINE config 5 in miniature (demo/main.go:116-206's shape): searcher threads hammer Set.Search while one thread
Adds and Deletes thousands of rows; afterwards the set must equal the oracle's model of the same operation sequence
row for row -- the slot map (Export in slot order: ids and float32 bytes), the scanned / live row counts, and a batch
of searches bit for bit (ids, order, scores; tombstones compete per block, block.go:236-243).

What it pins besides parity: Add stages its rows and their int8 shadow OUTSIDE the set's exclusive lock and places them
device-to-device under it, Delete resolves ids outside and tombstones under it -- searches running meanwhile must
never see a half-written row (every result list is checked for order and for ids that existed at some point)."""
import threading

import numpy as np
import pytest

from oracle import lib as ol
from oracle import model

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("multi", [False, True])
def test_churn_then_full_parity(built, multi):
    import torch
    from gofeature_b200 import api
    dims, cap, n0 = 512, 16384, 150_000
    have = torch.cuda.device_count()
    ctx = api.NewContextMulti([i % have for i in range(2)]) if multi else api.NewContext(0)
    bs = cap * dims * 4
    cache = api.NewCache(ctx, 16, bs)
    cache.NewSet("churn", dims, 4, 16)
    st = cache.GetSet("churn")
    oc = model.Cache(16, bs)
    oc.new_set("churn", dims, 4, 16)
    ost = oc.get_set("churn")
    rows0 = ol.synth_rows(31, 0, n0, dims)
    ids0 = ["f%011d" % i for i in range(n0)]
    st.AddSynthetic(n0, 31)                      # implicit ids on the device side, explicit in the model
    ost.add(rows0, ids0)
    st.SetCoalescing(50)

    stop = threading.Event()
    errs, counts = [], [0] * 4
    ever = set(ids0)
    qpool = ol.synth_rows(77, 0, 64, dims)

    def searcher(t):
        i = 0
        try:
            while not stop.is_set():
                res = st.SearchArray(-2.0, 10, qpool[(i * 4 + t) % 64:(i * 4 + t) % 64 + 1])[0]
                sc = [float(r.Score) for r in res]
                assert sc == sorted(sc, reverse=True) and 1 <= len(res) <= 10
                assert all(r.ID in ever for r in res), [r.ID for r in res]
                counts[t] += 1
                i += 1
        except Exception as exc:  # noqa: BLE001
            errs.append(exc)

    th = [threading.Thread(target=searcher, args=(t,)) for t in range(4)]
    for t in th:
        t.start()
    rng = np.random.default_rng(9)
    live = list(ids0)
    nxt = 0
    try:
        for step in range(6):
            victims = [live[j] for j in rng.choice(len(live), 1500, replace=False)]
            got = sorted(st.Delete(*victims) or [])
            assert got == sorted(ost.delete(victims)) == sorted(victims)
            vs = set(victims)
            live = [i for i in live if i not in vs]
            m = 2500 if step % 2 == 0 else 700          # refills only / refills + growth into fresh blocks
            v = rng.uniform(-1, 1, (m, dims)).astype(np.float32)
            v /= np.linalg.norm(v, axis=1, keepdims=True)
            ids = ["a%09d" % (nxt + i) for i in range(m)]
            nxt += m
            ever.update(ids)
            st.AddArray(v, ids)
            ost.add(v, ids)
            live += ids
    finally:
        stop.set()
        for t in th:
            t.join()
    assert not errs, errs[:3]
    assert min(counts) > 0, counts                      # searches really ran next to the mutations
    # ---- the slot map and the answers after the churn
    assert st.ScannedRows() == sum(b.next_index for b in ost.blocks)
    assert st.LiveRows() == len(live) == sum(1 for b in ost.blocks for r in range(b.next_index) if b.ids[r] != "")
    vals, ids = st.Export()
    want_ids = [b.ids[r] for b in ost.blocks for r in range(b.next_index) if b.ids[r] != ""]
    assert ids == want_ids
    want_vals = np.stack([b.rows[r] for b in ost.blocks for r in range(b.next_index) if b.ids[r] != ""])
    assert vals.tobytes() == want_vals.tobytes()
    st.SetCoalescing(0)
    q = qpool[:16].copy()
    q[0] = want_vals[123]
    q[1] = want_vals[-1]
    for limit in (1, 10, 40):
        got = st.SearchArray(-2.0, limit, q)
        want = ost.search(-2.0, limit, q)
        for g, w in zip(got, want):
            assert [r.ID for r in g] == [i for _, i in w]
            assert [np.float32(r.Score) for r in g] == [np.float32(s) for s, _ in w]
    cache.Close()
    ctx.Close()
