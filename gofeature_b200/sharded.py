"""Row-sharded feature set over the ranks of one torch.distributed job (one process per GPU).

The reference is single-device (`devices[0]`, search_test.go:40); what it already does across
*blocks* -- independent per-block top-limit, then a merge (set.go:129-157) -- is done here
across *ranks*: global block g lives on rank g % world as that rank's block g // world
(SURVEY 8e), every rank scans only its own rows, and the only exchange is one all-gather of
`limit` candidate keys per query per rank (80 bytes per query per rank at limit 10), followed
by the same merge kernel that folds the per-CTA lists.

torch is plumbing here: device tensors for the keys, `torch.distributed` (NCCL over
NVLink/NVSwitch on GPUs, gloo for the CPU tests of this host logic) for the gather.

The collective steps are free functions on tensors of any device so that the host logic
(striping, key merge, id exchange) is testable with gloo at world_size 2 without a GPU;
`ShardedSet` binds them to the CUDA library.
"""
import os

import numpy as np

from . import api

ID_WIDTH = 64  # bytes of a feature id carried through the id exchange


def owner_of(global_block, world):
    """(rank, local block position) of a global block (block striping)."""
    return global_block % world, global_block // world


def split_rows(n_existing, n_new, cap, world):
    """Which rank receives which of `n_new` appended rows when the sharded set already holds
    `n_existing` rows: rows fill global blocks of `cap` rows in order (set.go:97-114), block g on
    rank g % world.  Returns a list of (rank, start, stop) slices of the new rows, in order."""
    out = []
    pos, end = n_existing, n_existing + n_new
    while pos < end:
        g = pos // cap
        stop = min(end, (g + 1) * cap)
        out.append((g % world, pos - n_existing, stop - n_existing))
        pos = stop
    return out


def slot_owner(slot, cap, world):
    """rank that owns a global slot (= global block * cap + row)."""
    return (slot // cap) % world


def merge_keys_host(gathered, limit):
    """Merge of per-rank key lists on the host: gathered [world, q, limit] uint64 (descending,
    zero padded) -> [q, limit].  Same total order as the device merge kernel (gf_merge.cu)."""
    g = np.asarray(gathered, dtype=np.uint64)
    world, q, k = g.shape
    out = np.zeros((q, limit), dtype=np.uint64)
    for i in range(q):
        allk = np.sort(g[:, i, :].reshape(-1))[::-1]
        allk = allk[allk != 0][:limit]
        out[i, :len(allk)] = allk
    return out


def gather_keys(local_keys, world, group=None, out=None):
    """all-gather of the per-rank winners: local_keys [q, limit] int64 tensor (any device) ->
    [world, q, limit].  The only data-path collective of a sharded search."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_keys.unsqueeze(0)
    q, limit = local_keys.shape
    if out is None:
        out = torch.empty((world, q, limit), dtype=local_keys.dtype, device=local_keys.device)
    # rank-major concatenation along dim 0 (the layout both NCCL and gloo accept)
    dist.all_gather_into_tensor(out.view(world * q, limit), local_keys.contiguous(), group=group)
    return out


def exchange_ids(keys, mine, rank, world, group=None, device=None):
    """Every rank knows ids / tombstones only for the slots it owns.  `mine[i]` maps slot ->
    id (str) for this rank's LIVE winners of query i.  One all-reduce(MAX) of a byte table makes
    the table complete on every rank.  Returns table uint8 [q, limit, 2 + ID_WIDTH]:
    [...,0] = 1 if the owner reported the slot live, [...,1] = id length, then the id bytes."""
    import torch
    import torch.distributed as dist
    q, limit = keys.shape
    table = np.zeros((q, limit, ID_WIDTH + 2), dtype=np.uint8)
    for i in range(q):
        for j in range(limit):
            k = int(keys[i, j])
            if k == 0:
                break
            slot = 0xFFFFFFFF - (k & 0xFFFFFFFF)
            ident = mine[i].get(slot)
            if ident is not None:
                b = ident.encode() if isinstance(ident, str) else bytes(ident)
                b = b[:ID_WIDTH]
                table[i, j, 0], table[i, j, 1] = 1, len(b)
                table[i, j, 2:2 + len(b)] = np.frombuffer(b, dtype=np.uint8)
    if world > 1:
        tt = torch.from_numpy(table)
        if device is not None:
            tt = tt.to(device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX, group=group)
        table = tt.cpu().numpy()
    return table


def results_from_table(keys, table, threshold):
    """[][]FeatureSearchResult from merged keys + the exchanged id table: tombstones (not reported
    live by their owner) and scores below the threshold (set.go:145) are dropped."""
    q, limit = keys.shape
    thr = np.float32(threshold)
    out = []
    for i in range(q):
        row = []
        for j in range(limit):
            k = int(keys[i, j])
            if k == 0:
                break
            if not table[i, j, 0]:
                continue
            sc = _key_score(k)
            if not (sc >= thr):
                continue
            n = int(table[i, j, 1])
            row.append(api.FeatureSearchResult(sc, bytes(table[i, j, 2:2 + n]).decode(errors="replace"),
                                               0xFFFFFFFF - (k & 0xFFFFFFFF)))
        out.append(row)
    return out


def tombstone_among_winners(keys, table):
    """True if a merged winner was not reported live by its owner (a zeroed, deleted row scored its way in)."""
    q, limit = keys.shape
    for i in range(q):
        for j in range(limit):
            if int(keys[i, j]) == 0:
                break
            if not table[i, j, 0]:
                return True
    return False


def merge_hit_lists(gathered, limit):
    """gathered[rank][query] = [(key, id)] already filtered per block / threshold -> MaxNFeatureResult
    (set.go:154-157) over the union, in canonical key order."""
    nq = len(gathered[0])
    out = []
    for i in range(nq):
        allh = sorted((h for per_rank in gathered for h in per_rank[i]), key=lambda t: -t[0])[:limit]
        out.append([api.FeatureSearchResult(_key_score(k), ident, 0xFFFFFFFF - (k & 0xFFFFFFFF)) for k, ident in allh])
    return out


def _key_score(k):
    """score of a key without calling into the library (same mapping as gf_key_score)."""
    im = (k >> 32) & 0xFFFFFFFF
    u = (im & 0x7FFFFFFF) if (im & 0x80000000) else (~im & 0xFFFFFFFF)
    return np.frombuffer(np.uint32(u).tobytes(), dtype=np.float32)[0]


class ShardedSet:
    """One logical set, row-sharded over `world` ranks.  Every rank constructs it collectively.

    Append-only through this wrapper (Add / AddSynthetic); Delete is forwarded to every rank and
    each removes the ids it owns.  Search is collective and returns the same result on every rank.
    """

    def __init__(self, ctx, cache, name, dims, batch, rank, world, group=None):
        self.ctx, self.cache, self.name = ctx, cache, name
        self.dims, self.batch, self.rank, self.world, self.group = dims, batch, rank, world, group
        cache.NewSet(name, dims, 4, batch)
        self.local = cache.GetSet(name)
        self.local.SetShard(rank, world)
        self.cap = cache.GetBlockSize() // (dims * 4)
        self.total_rows = 0
        self._bufs = {}
        self._xchg = None
        self.refused_queries = 0   # queries some rank's certificate refused (re-run through the exact scan)
        if world > 1 and os.environ.get("GF_NO_P2P") is None:
            self._connect_exchange()

    # ---- peer exchange (one kernel over NVLink peer memory instead of all-gather + merge)
    def _connect_exchange(self):
        """Collective.  Every rank maps every other rank's inbox through CUDA IPC; if any rank cannot
        (no peer access, IPC closed in the container ...) ALL ranks keep the NCCL all-gather path."""
        import ctypes
        import torch
        import torch.distributed as dist
        L = api.lib()
        dev = "cuda:%d" % self.ctx.device
        self._x_max_q = max(self.batch, 16)
        self._x_max_k = min(256, 2048 // self.world)
        h = ctypes.c_void_p()
        handle = (ctypes.c_uint8 * 64)()
        ok = L.gf_exchange_create(self.ctx._h, self.rank, self.world, self._x_max_q, self._x_max_k, ctypes.byref(h)) == 0
        ok = ok and L.gf_exchange_handle(h, handle, 64) == 0
        mine = torch.tensor(list(bytes(handle)) + [1 if ok else 0], dtype=torch.uint8, device=dev)
        everyone = torch.empty((self.world, 65), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        everyone = everyone.cpu().numpy()
        if ok and everyone[:, 64].all():
            blob = np.ascontiguousarray(everyone[:, :64]).tobytes()
            ok = L.gf_exchange_connect(h, blob, len(blob)) == 0
        else:
            ok = False
        agreed = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN, group=self.group)
        if int(agreed.item()) == 1:
            self._xchg = h
        elif h:
            L.gf_exchange_destroy(h)

    def exchange_trace(self, q):
        """[q, 4] device-clock ns of the last fused search's cross-rank step on THIS rank (entered, published, all
        ranks seen, folded); needs ctx.SetProfiling(True) while it ran.  None without the peer exchange."""
        if self._xchg is None:
            return None
        out = np.zeros((q, 4), dtype=np.uint64)
        api._check(api.lib().gf_exchange_trace(self._xchg, q, out.ctypes.data))
        return out

    def close_exchange(self):
        if self._xchg is not None:
            api.lib().gf_exchange_destroy(self._xchg)
            self._xchg = None

    # ---- mutation (collective: every rank passes the same arguments)
    def Add(self, values, ids):
        """ids longer than ID_WIDTH bytes are rejected (the id exchange of `Search` carries fixed-width ids; the
        in-library multi-device set, gf_context_create_multi, has no such limit).  A duplicate id is deleted by
        `Delete` on EVERY rank that holds it, whereas one set deletes only its first occurrence (set.go:163-193)."""
        for i in ids:
            if len(i.encode() if isinstance(i, str) else bytes(i)) > ID_WIDTH:
                raise api.GoFeatureError(api.ErrInvalidFeautres)
        values = np.ascontiguousarray(values, dtype=np.float32).reshape(len(ids), self.dims)
        for r, a, b in split_rows(self.total_rows, len(ids), self.cap, self.world):
            if r == self.rank:
                self.local.AddArray(values[a:b], ids[a:b])
        self.total_rows += len(ids)

    def AddSynthetic(self, n, seed, normalize=True, id_prefix="f", centres=0, spread=0.0, dup_ppm=0):
        for r, a, b in split_rows(self.total_rows, n, self.cap, self.world):
            if r == self.rank:
                self.local.AddSynthetic(b - a, seed, self.total_rows + a, normalize, id_prefix, centres, spread, dup_ppm)
        self.total_rows += n

    def Delete(self, *ids):
        import torch
        import torch.distributed as dist
        mine = set(self.local.Delete(*ids) or [])
        flags = torch.tensor([1 if i in mine else 0 for i in ids], dtype=torch.int32,
                             device="cuda:%d" % self.ctx.device)
        if self.world > 1:
            dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=self.group)
        return [i for i, f in zip(ids, flags.cpu().tolist()) if f] or None

    # ---- search
    def _tensors(self, q, limit):
        import torch
        key = (q, limit)
        if key not in self._bufs:
            dev = "cuda:%d" % self.ctx.device
            self._bufs[key] = dict(
                dq=torch.empty((q, self.dims), dtype=torch.float32, device=dev),
                local=torch.zeros((q, limit), dtype=torch.int64, device=dev),
                lflags=torch.zeros((q,), dtype=torch.int32, device=dev),
                mflags=torch.zeros((q,), dtype=torch.int32, device=dev),
                hflags=torch.zeros((q,), dtype=torch.int32).pin_memory(),
                gathered=torch.zeros((self.world, q, limit), dtype=torch.int64, device=dev),
                merged=torch.zeros((q, limit), dtype=torch.int64, device=dev),
                hq=torch.empty((q, self.dims), dtype=torch.float32).pin_memory(),
                hk=torch.empty((q, limit), dtype=torch.int64).pin_memory(),
            )
        return self._bufs[key]

    def search_device(self, limit, q, dq, path=0):
        """Collective device-resident search: dq [q, dims] float32 on this rank's GPU (same on all
        ranks) -> merged keys [q, limit] int64 tensor, identical on every rank.  Enqueues on the
        current torch stream; no host synchronisation.  `self.flags(q, limit)` is non-zero for a query
        that some rank could not certify on its tensor path (the caller re-runs with path=1)."""
        import torch
        import torch.distributed as dist
        t = self._tensors(q, limit)
        stream = torch.cuda.current_stream().cuda_stream
        L = api.lib()
        if self.world == 1:   # the local flags are the merged flags
            api._check(L.gf_set_search_device_ex(self.local._h, limit, q, dq.data_ptr(), t["local"].data_ptr(),
                                                 t["mflags"].data_ptr(), path, stream))
            return t["local"]
        if self._xchg is not None and q <= self._x_max_q and limit <= self._x_max_k:
            # search + exchange as one call: the one-pass chain's last kernel pushes my winners into every peer's
            # inbox, waits for theirs and merges (otherwise the exchange kernel is launched behind the search)
            api._check(L.gf_set_search_device_xchg(self.local._h, self._xchg, limit, q, dq.data_ptr(),
                                                   t["local"].data_ptr(), t["lflags"].data_ptr(),
                                                   t["merged"].data_ptr(), t["mflags"].data_ptr(), path, stream))
            return t["merged"]
        api._check(L.gf_set_search_device_ex(self.local._h, limit, q, dq.data_ptr(), t["local"].data_ptr(),
                                             t["lflags"].data_ptr(), path, stream))
        gathered = gather_keys(t["local"], self.world, self.group, out=t["gathered"])
        api._check(L.gf_merge_keys_device(self.ctx._h, self.world, q, limit, gathered.data_ptr(),
                                          t["merged"].data_ptr(), stream))
        t["mflags"].copy_(t["lflags"], non_blocking=True)
        dist.all_reduce(t["mflags"], op=dist.ReduceOp.MAX, group=self.group)
        return t["merged"]

    def flags(self, q, limit):
        return self._tensors(q, limit)["mflags"]

    def search_keys(self, limit, queries):
        """Host queries [q, dims] -> merged keys [q, limit] uint64 on the host (H2D, scan, exchange,
        merge, D2H all inside)."""
        import torch
        queries = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dims)
        q = queries.shape[0]
        if q > self.batch:
            raise api.GoFeatureError(api.ErrOutOfBatch)  # set.go:119-122
        t = self._tensors(q, limit)
        t["hq"].numpy()[:] = queries
        t["dq"].copy_(t["hq"], non_blocking=True)
        merged = self.search_device(limit, q, t["dq"])
        t["hk"].copy_(merged, non_blocking=True)
        t["hflags"].copy_(t["mflags"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        hf = t["hflags"].numpy()
        if hf.any():
            if (hf == -1).any():
                raise RuntimeError("gofeature_b200: a peer rank did not reach the exchange within 5 s")
            # some rank could not certify a query on its tensor path (every rank sees the same flags): ONLY those
            # queries go through the exact scan kernel again, as one compact batch on every rank
            bad = np.nonzero(hf)[0]
            self.refused_queries += len(bad)
            out = t["hk"].numpy().view(np.uint64).copy()
            sub = t["dq"][torch.from_numpy(bad).to(t["dq"].device)].contiguous()
            t2 = self._tensors(len(bad), limit)
            merged = self.search_device(limit, len(bad), sub, path=1)
            t2["hk"].copy_(merged, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            out[bad] = t2["hk"].numpy().view(np.uint64)
            return out
        return t["hk"].numpy().view(np.uint64).copy()

    def Search(self, threshold, limit, queries):
        """Collective Set.Search (set.go:118-159) over the sharded rows.  IDs are resolved by the
        rank that owns each winning slot and exchanged as fixed-width bytes.  Tombstones that reach
        the merged top-limit are dropped after selection, as block.go:236-243 does per block."""
        keys = self.search_keys(limit, queries)
        local = self.local.ResolveKeys(float("-inf"), limit, keys)     # live winners this rank owns
        mine = [{r.Slot: r.ID for r in row if r.ID != ""} for row in local]
        table = exchange_ids(keys, mine, self.rank, self.world, self.group, "cuda:%d" % self.ctx.device)
        if not tombstone_among_winners(keys, table):
            return results_from_table(keys, table, threshold)
        # A deleted row made it into the merged top-limit.  The reference selects per BLOCK before it drops
        # tombstones (block.go:236-243), so the exact answer is the merge of the per-block lists: every rank
        # runs its own Set.Search (which applies that rule to its blocks) and the hit lists are merged.
        return self._search_exact_lists(threshold, limit, queries)

    def _search_exact_lists(self, threshold, limit, queries):
        import torch.distributed as dist
        queries = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dims)
        local = self.local.SearchArray(threshold, limit, queries)
        mine = [[(api.key_make(r.Score, r.Slot), r.ID) for r in row] for row in local]
        if self.world > 1:
            gathered = [None] * self.world
            dist.all_gather_object(gathered, mine, group=self.group)
        else:
            gathered = [mine]
        return merge_hit_lists(gathered, limit)
