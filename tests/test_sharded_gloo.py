"""CPU tests (gloo, world_size 2) of the multi-rank host logic in gofeature_b200/sharded.py:
block striping, the all-gather of K candidate keys per query, the merge order and the id /
tombstone exchange.  The per-rank scan is stood in by the oracle (the CUDA scan itself is covered
by the -m gpu tests); everything between the ranks is the product code."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gofeature_b200 import sharded
from oracle import MODE_CANONICAL
from oracle import lib as ol
from oracle import model

DIMS, CAP, N, LIMIT = 64, 50, 530, 7


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _key(score, slot):
    im = int(ol.lib().gf_oracle_score_image(np.float32(score)))
    return np.uint64((im << 32) | (0xFFFFFFFF - slot))


def _worker(rank, world, port, dead, q_out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows = ol.synth_rows(77, 0, N, DIMS)
        ids = ["g%05d" % i for i in range(N)]
        queries = ol.synth_rows(78, 0, 3, DIMS)
        # this rank's rows, in local order, with their global slots (block striping)
        mine_rows, mine_slots = [], []
        for r, a, b in sharded.split_rows(0, N, CAP, world):
            if r == rank:
                mine_rows.append(rows[a:b])
                mine_slots.extend(range(a, b))        # append-only: global slot == global row ordinal
        local = np.concatenate(mine_rows)
        slots = np.array(mine_slots)
        assert all(sharded.slot_owner(s, CAP, world) == rank for s in slots)
        local = local.copy()
        live = np.array([ids[s] not in dead for s in slots])
        local[~live] = 0.0                            # the owner zeroes deleted rows (block.go:152-158)
        # local scan -> keys with GLOBAL slots (what gf_set_search_device emits after gf_set_set_shard)
        keys = np.zeros((3, LIMIT), dtype=np.uint64)
        for i, (sc, idx) in enumerate(ol.scan_topk(local, queries, LIMIT, MODE_CANONICAL, 1)):
            cand = sorted((_key(s_, int(slots[j])) for s_, j in zip(sc, idx)), reverse=True)
            keys[i, :len(cand)] = cand
        gathered = sharded.gather_keys(torch.from_numpy(keys.view(np.int64)), world)
        merged = sharded.merge_keys_host(gathered.numpy().view(np.uint64), LIMIT)
        mine = [{int(s): ids[s] for s in slots[live]} for _ in range(3)]
        table = sharded.exchange_ids(merged, mine, rank, world)
        if sharded.tombstone_among_winners(merged, table):
            # exact path: every rank answers with its own per-block lists (the oracle model of its shard)
            oc = model.Cache(20, CAP * DIMS * 4)
            oc.new_set("shard", DIMS, 4, 4)
            os_ = oc.get_set("shard")
            for r, a, b in sharded.split_rows(0, N, CAP, world):
                if r == rank:
                    os_.add(rows[a:b], ids[a:b])
            os_.delete(list(dead))
            shard_slots = {ids[s]: int(s) for s in slots}
            local_hits = [[(int(_key(sc, shard_slots[ident])), ident) for sc, ident in per_q]
                          for per_q in os_.search(-1.0, LIMIT, queries)]
            gathered_hits = [None] * world
            dist.all_gather_object(gathered_hits, local_hits)
            res = sharded.merge_hit_lists(gathered_hits, LIMIT)
        else:
            res = sharded.results_from_table(merged, table, -1.0)
        q_out.put((rank, [[(float(r.Score), r.ID, r.Slot) for r in row] for row in res]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dead", [(), ("g00003", "g00101", "g00417")])
def test_two_rank_sharded_search_matches_single_set(dead):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dead, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q_out.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == got[1]                           # every rank ends with the same answer
    # single-set oracle: one cache, same rows; tombstones compete then drop (block.go:236-243)
    c = model.Cache(20, CAP * DIMS * 4)
    c.new_set("s", DIMS, 4, 4)
    s = c.get_set("s")
    s.add(ol.synth_rows(77, 0, N, DIMS), ["g%05d" % i for i in range(N)])
    if dead:
        s.delete(list(dead))
    want = s.search(-1.0, LIMIT, ol.synth_rows(78, 0, 3, DIMS))
    for g, w in zip(got[0], want):
        # the sharded answer IS the single-set answer, tombstones included (per-block selection rule)
        assert [i for _, i, _ in g] == [i for _, i in w]
        assert [np.float32(sc) for sc, _, _ in g] == [np.float32(sc) for sc, _ in w]


def test_split_rows_and_owner():
    assert sharded.split_rows(0, 120, 50, 2) == [(0, 0, 50), (1, 50, 100), (0, 100, 120)]
    assert sharded.split_rows(120, 40, 50, 2) == [(0, 0, 30), (1, 30, 40)]
    assert sharded.split_rows(0, 0, 50, 4) == []
    assert [sharded.owner_of(g, 4) for g in range(6)] == [(0, 0), (1, 0), (2, 0), (3, 0), (0, 1), (1, 1)]
    assert sharded.slot_owner(149, 50, 4) == 2


def test_merge_keys_host_order():
    a = np.array([[[9, 7, 0], [5, 0, 0]], [[8, 7 + 1, 6], [0, 0, 0]]], dtype=np.uint64)
    m = sharded.merge_keys_host(a, 3)
    assert m.tolist() == [[9, 8, 8], [5, 0, 0]]


def test_merge_hit_lists_is_maxn_over_the_union():
    from gofeature_b200 import api  # noqa: F401  (key helpers need the library only for key_make)
    k = lambda s, slot: (int(ol.lib().gf_oracle_score_image(np.float32(s))) << 32) | (0xFFFFFFFF - slot)
    gathered = [[[(k(0.5, 4), "a"), (k(0.1, 9), "b")]], [[(k(0.5, 2), "c"), (k(0.3, 7), "d")]]]
    out = sharded.merge_hit_lists(gathered, 3)
    assert [r.ID for r in out[0]] == ["c", "a", "d"]          # score desc, then slot asc
    assert [r.Slot for r in out[0]] == [2, 4, 7]
