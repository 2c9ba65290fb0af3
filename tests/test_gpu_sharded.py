"""GPU tests of the row-sharded path.

On one GPU the ranks are emulated inside one process: two shard sets (gf_set_set_shard rank 0 / 1
of 2) on the same device, their device-resident winners concatenated exactly as the all-gather
would lay them out, folded by gf_merge_keys_device.  With >= 2 GPUs the real collective path runs
under torch.distributed (NCCL)."""
import os
import socket

import numpy as np
import pytest

from oracle import lib as ol
from oracle import model

pytestmark = pytest.mark.gpu


def test_two_shards_on_one_gpu_equal_the_unsharded_set(gpu_ctx):
    import torch
    from gofeature_b200 import api, sharded
    dims, cap, n, limit, q = 512, 1024, 9000, 10, 5
    rows = ol.synth_rows(61, 0, n, dims)
    ids = ["s%05d" % i for i in range(n)]
    queries = ol.synth_rows(62, 0, q, dims)
    caches, sets = [], []
    for rank in range(2):
        c = api.NewCache(gpu_ctx, 6, cap * dims * 4)
        c.NewSet("shard", dims, 4, 8)
        s = c.GetSet("shard")
        s.SetShard(rank, 2)
        for r, a, b in sharded.split_rows(0, n, cap, 2):
            if r == rank:
                s.AddArray(rows[a:b], ids[a:b])
        caches.append(c)
        sets.append(s)
    assert sets[0].ScannedRows() + sets[1].ScannedRows() == n
    dq = torch.from_numpy(queries).cuda()
    gathered = torch.zeros((2, q, limit), dtype=torch.int64, device="cuda")
    merged = torch.zeros((q, limit), dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for rank in range(2):
        sets[rank].SearchDevice(limit, q, dq.data_ptr(), gathered[rank].data_ptr(), 0, stream)
    api._check(api.lib().gf_merge_keys_device(gpu_ctx._h, 2, q, limit, gathered.data_ptr(), merged.data_ptr(), stream))
    torch.cuda.synchronize()
    keys = merged.cpu().numpy().view(np.uint64)
    assert (keys == sharded.merge_keys_host(gathered.cpu().numpy().view(np.uint64), limit)).all()
    oc = model.Cache(12, cap * dims * 4)
    oc.new_set("one", dims, 4, 8)
    o = oc.get_set("one")
    o.add(rows, ids)
    want = o.search(-1.0, limit, queries)
    for i in range(q):
        got_ids = []
        for k in keys[i]:
            slot = api.key_slot(k)
            got_ids.append(ids[slot])                  # append-only: global slot == row ordinal
            owner = sharded.slot_owner(slot, cap, 2)
            res = sets[owner].ResolveKeys(-1.0, 1, np.array([k], dtype=np.uint64))
            assert res[0][0].ID == ids[slot]           # the owning shard resolves the id
        assert got_ids == [x for _, x in want[i]]
        assert [float(api.key_score(k)) for k in keys[i]] == [float(np.float32(s_)) for s_, _ in want[i]]
    for c in caches:
        c.Close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


_BIG_ROWS = 180000


def _negative_set():
    rows = ol.synth_rows(63, 0, 3000, 512).copy()
    rows[:, 0] = np.abs(rows[:, 0]) + np.float32(0.5)
    ids = ["n%05d" % i for i in range(3000)]
    q = np.zeros((2, 512), dtype=np.float32)
    q[0, 0] = -1.0
    q[1, 0] = -1.0
    q[1, 1] = 0.25
    dead = [ids[i] for i in (3, 1025, 1030, 2500, 2999)]
    return rows, ids, q, dead


def _nccl_worker(rank, world, port, q_out):
    import torch
    import torch.distributed as dist
    from gofeature_b200 import api
    from gofeature_b200.sharded import ShardedSet
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ctx = api.NewContext(rank)
        cache = api.NewCache(ctx, 8, 1024 * 512 * 4)
        s = ShardedSet(ctx, cache, "nccl", 512, 8, rank, world)
        s.AddSynthetic(9000, 61)
        dead = ["f%011d" % i for i in (5, 1030, 4444)]
        assert sorted(s.Delete(*dead)) == sorted(dead)
        res = s.Search(-1.0, 10, ol.synth_rows(62, 0, 3, 512))
        # second set: every live score is negative, so the zeroed tombstones win the merged top-10 and the
        # exact per-block path (block.go:236-243) has to run across the ranks
        rows2, ids2, q2, dead2 = _negative_set()
        s2 = ShardedSet(ctx, cache, "neg", 512, 8, rank, world)
        s2.Add(rows2, ids2)
        s2.Delete(*dead2)
        res2 = s2.Search(-10.0, 10, q2)
        # third set: shards big enough for the tcgen05 path -- the one-pass chain whose tail kernel ENDS with the
        # cross-rank exchange (gf_set_search_device_xchg), 10 and then 40 queries (the latter: two-pass chain +
        # the separate exchange kernel), and a deleted winner
        cache3 = api.NewCache(ctx, 7, 16384 * 512 * 4)
        s3 = ShardedSet(ctx, cache3, "big", 512, 64, rank, world)
        s3.AddSynthetic(_BIG_ROWS, 71)
        big = ol.synth_rows(71, 0, _BIG_ROWS, 512)
        q3 = ol.synth_rows(72, 0, 40, 512)
        q3[0] = big[123456]
        q3[1] = big[7]
        ctx.ResetStats()
        res3 = s3.Search(-1.0, 10, q3[:10])
        st3 = ctx.Stats()
        assert st3["tensor_passes"] >= 1 and st3["merge_launches"] == 0, st3   # no exchange launch: the tail did it
        res4 = s3.Search(-1.0, 10, q3)
        s3.Delete("f%011d" % 123456)
        res5 = s3.Search(-1.0, 10, q3[:10])
        q_out.put((rank, [[(float(r.Score), r.ID) for r in row] for row in res],
                   [[(float(r.Score), r.ID) for r in row] for row in res2],
                   [[[(float(r.Score), r.ID) for r in row] for row in rr] for rr in (res3, res4, res5)]))
        cache3.Close()
        cache.Close()
        ctx.Close()
    finally:
        dist.destroy_process_group()


def test_nccl_sharded_search_two_gpus():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q_out.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = {r: a for r, a, _, _ in outs}
    got2 = {r: b for r, _, b, _ in outs}
    got3 = {r: c for r, _, _, c in outs}
    assert got[0] == got[1] and got2[0] == got2[1] and got3[0] == got3[1]
    oc = model.Cache(12, 1024 * 512 * 4)
    oc.new_set("one", 512, 4, 8)
    o = oc.get_set("one")
    o.add(ol.synth_rows(61, 0, 9000, 512), ["f%011d" % i for i in range(9000)])
    o.delete(["f%011d" % i for i in (5, 1030, 4444)])
    want = o.search(-1.0, 10, ol.synth_rows(62, 0, 3, 512))
    for g, w in zip(got[0], want):
        assert [i for _, i in g] == [i for _, i in w]
        assert [s_ for s_, _ in g] == [float(np.float32(s_)) for s_, _ in w]
    rows2, ids2, q2, dead2 = _negative_set()
    o2c = model.Cache(12, 1024 * 512 * 4)
    o2c.new_set("neg", 512, 4, 8)
    o2 = o2c.get_set("neg")
    o2.add(rows2, ids2)
    o2.delete(dead2)
    want2 = o2.search(-10.0, 10, q2)
    for g, w in zip(got2[0], want2):
        assert [i for _, i in g] == [i for _, i in w]
        assert [s_ for s_, _ in g] == [float(np.float32(s_)) for s_, _ in w]
    # the tcgen05 shards (fused exchange in the one-pass tail / separate exchange kernel) against the oracle
    ob = model.Cache(12, 16384 * 512 * 4)
    ob.new_set("big", 512, 4, 64)
    o3 = ob.get_set("big")
    o3.add(ol.synth_rows(71, 0, _BIG_ROWS, 512), ["f%011d" % i for i in range(_BIG_ROWS)])
    q3 = ol.synth_rows(72, 0, 40, 512)
    big = ol.synth_rows(71, 0, _BIG_ROWS, 512)
    q3[0] = big[123456]
    q3[1] = big[7]
    wants = [o3.search(-1.0, 10, q3[:10]), o3.search(-1.0, 10, q3)]
    o3.delete(["f%011d" % 123456])
    wants.append(o3.search(-1.0, 10, q3[:10]))
    for gs, ws in zip(got3[0], wants):
        assert len(gs) == len(ws)
        for g, w in zip(gs, ws):
            assert [i for _, i in g] == [i for _, i in w]
            assert [s_ for s_, _ in g] == [float(np.float32(s_)) for s_, _ in w]

