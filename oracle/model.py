"""Object-level restatement of goFeature's Cache / Set / Block bookkeeping.

TEST INFRASTRUCTURE ONLY (see oracle/gf_oracle.c).  Block storage is a numpy
array instead of device memory; scoring and top-N go through the C oracle.
Each method cites the reference lines it follows (/root/reference/...).

Where the reference is defective the evident intent is modelled, and the
divergence is listed in DESIGN.md ("Reference defects not reproduced"):
  * block.go:103,115  Insert slices the byte vector by element counts; here the
    refill of Empty slots and the tail append copy whole rows.
  * block.go:150      `for id, index := range targets` iterates a Go map (random
    order); here deleted ids are reported in request order.
  * an empty FeatureID ("") marks a free slot (block.go:159,239); Add / Delete of
    "" are rejected with ErrInvalidFeautres instead of corrupting the free list.
"""
import numpy as np

from . import lib as olib


class OracleError(Exception):
    """Carries the name of the error.go sentinel."""

    def __init__(self, name):
        super().__init__(name)
        self.name = name


class Block:
    """_Block (block.go:14-37)."""

    def __init__(self, index, block_size):
        self.index = index
        self.block_size = block_size
        self.owner = ""
        self.dims = 0
        self.precision = 0
        self.batch = 0
        self.empty = []       # Empty []int
        self.next_index = 0   # NextIndex
        self.ids = []         # IDs []FeatureID
        self.rows = None      # float32 [capacity, dims]

    def is_owned(self):       # block.go:50
        return self.owner != ""

    def capacity(self):       # block.go:52
        return self.block_size // (self.precision * self.dims)

    def margin(self):         # block.go:54-57
        return len(self.empty) + (self.capacity() - self.next_index)

    def acquire(self, owner, dims, precision, batch):   # block.go:59-78
        if self.owner != "":
            raise OracleError("ErrBlockUsed")
        self.dims, self.precision, self.owner, self.batch = dims, precision, owner, batch
        cap = self.block_size // (precision * dims)
        self.ids = [""] * cap
        self.rows = np.zeros((cap, dims), dtype=np.float32)

    def insert(self, values, ids):                      # block.go:80-131 (intent)
        n = len(ids)
        if n > self.margin():
            raise OracleError("ErrBlockIsFull")
        refill = min(len(self.empty), n)
        for i in range(refill):
            slot = self.empty[i]
            self.rows[slot] = values[i]
            self.ids[slot] = ids[i]
        if n > len(self.empty):
            tail = n - len(self.empty)
            self.rows[self.next_index:self.next_index + tail] = values[len(self.empty):]
            for i in range(tail):
                self.ids[self.next_index + i] = ids[len(self.empty) + i]
            self.next_index += tail
            self.empty = []
        else:
            self.empty = self.empty[n:]

    def delete(self, ids):                              # block.go:135-165
        targets = {}
        for i in ids:
            targets.setdefault(i, -1)
        for index, value in enumerate(self.ids):
            if value in targets:
                targets[value] = index                  # last match wins, block.go:140-144
        deleted = []
        for i in targets:                               # request order (dict keeps insertion order)
            index = targets[i]
            if index != -1:
                self.rows[index] = 0.0                  # buffer.Reset, block.go:152-158
                self.ids[index] = ""
                self.empty.append(index)
                deleted.append(i)
        return deleted

    def release(self):                                  # block.go:253-270
        self.dims = self.precision = 0
        self.owner = ""
        self.ids, self.empty, self.next_index, self.rows = [], [], 0, None


class Set:
    """FeatureSet (set.go:24-38)."""

    def __init__(self, cache, name, dims, precision, batch):
        self.cache, self.name, self.dims, self.precision, self.batch = cache, name, dims, precision, batch
        self.blocks = []

    def add(self, values, ids):                         # set.go:77-116
        values = np.ascontiguousarray(values, dtype=np.float32).reshape(len(ids), self.dims)
        if any(i == "" for i in ids):
            raise OracleError("ErrInvalidFeautres")
        empty = sum(b.margin() for b in self.blocks)
        n = len(ids)
        if n > empty:
            remain = n - empty
            block_length = self.cache.block_size // (self.dims * self.precision)
            block_num = (remain + block_length - 1) // block_length
            blocks = self.cache.get_empty_block(block_num)
            for b in blocks:
                b.acquire(self.name, self.dims, self.precision, self.batch)
            self.blocks.extend(blocks)
        offset, remain = 0, n
        for b in self.blocks:
            length = min(b.margin(), remain)
            if length > 0:
                b.insert(values[offset:offset + length], ids[offset:offset + length])
                offset += length
                remain -= length
            if remain <= 0:
                break

    def delete(self, ids):                              # set.go:163-193
        ids = list(ids)
        if any(i == "" for i in ids):
            raise OracleError("ErrInvalidFeautres")
        deleted = []
        for b in self.blocks:
            if not ids:
                break
            d = b.delete(ids)
            if not d:
                continue
            deleted.extend(d)
            ids = [i for i in ids if i not in d]
        return deleted

    def search(self, threshold, limit, queries, mode=olib.MODE_CANONICAL):   # set.go:118-159
        queries = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dims)
        if queries.shape[0] > self.batch:
            raise OracleError("ErrOutOfBatch")
        if queries.shape[0] == 0:
            return []
        rows = [b.rows for b in self.blocks]
        heights = [b.next_index for b in self.blocks]
        live = [np.array([1 if i != "" else 0 for i in b.ids], dtype=np.uint8) for b in self.blocks]
        raw = olib.set_search(rows, heights, live, queries, limit, threshold, mode)
        return [[(s, self.blocks[bi].ids[ri]) for (s, bi, ri) in per_q] for per_q in raw]

    def compact(self):
        """Extension (no reference counterpart): per block, live rows move to the front in
        order, NextIndex shrinks to the live count, the free list empties."""
        freed = 0
        for b in self.blocks:
            keep = [r for r in range(b.next_index) if b.ids[r] != ""]
            if len(keep) == b.next_index:
                continue
            new_rows = np.zeros_like(b.rows)
            new_rows[:len(keep)] = b.rows[keep]
            new_ids = [b.ids[r] for r in keep] + [""] * (len(b.ids) - len(keep))
            freed += b.next_index - len(keep)
            b.rows, b.ids, b.next_index, b.empty = new_rows, new_ids, len(keep), []
        return freed

    def destroy(self):                                  # set.go:202-214
        for b in self.blocks:
            b.release()
        self.blocks = []


class Cache:
    """_Cache (cache.go:12-19)."""

    def __init__(self, block_num, block_size):          # cache.go:21-43
        self.block_size = block_size
        self.all_blocks = [Block(i, block_size) for i in range(block_num)]
        self.sets = {}

    def new_set(self, name, dims, precision, batch):    # cache.go:45-72
        if name in self.sets:
            raise OracleError("ErrFeatureSetExist")
        self.sets[name] = Set(self, name, dims, precision, batch)

    def destroy_set(self, name):                        # cache.go:74-90
        if name not in self.sets:
            raise OracleError("ErrFeatureSetNotFound")
        self.sets[name].destroy()
        del self.sets[name]

    def get_set(self, name):                            # cache.go:92-100
        if name not in self.sets:
            raise OracleError("ErrFeatureSetNotFound")
        return self.sets[name]

    def get_block_size(self):                           # cache.go:102
        return self.block_size

    def get_empty_block(self, block_num):               # cache.go:104-115
        empty = [b for b in self.all_blocks if not b.is_owned()]
        if len(empty) < block_num:
            raise OracleError("ErrNotEnoughBlocks")
        return empty[:block_num]
