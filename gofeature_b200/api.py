"""Host-side mirror of goFeature's Go API over the C ABI.

Names, argument meaning and error behaviour follow the reference
(/root/reference/interface.go:12-143, proto.go:3-26, util.go, error.go) so that tests read
like the reference's own: NewCache / Cache.NewSet / Set.Add / Set.Search ...  Go's
`(value, err)` returns become a return value or a raised `GoFeatureError` whose `.name`
is the Go sentinel ("ErrOutOfBatch").  All work happens in libgofeature_b200.so; this file
only packs bytes.
"""
import ctypes
import random
import string

import numpy as np

from ._capi import c_i64, c_int, c_u32, c_u64, c_vp, gf_stats, lib

# ------------------------------------------------------------------ errors (error.go:9-49)


class GoFeatureError(Exception):
    def __init__(self, status):
        L = lib()
        self.status = status
        self.name = L.gf_status_name(status).decode()
        self.message = L.gf_strerror(status).decode()
        detail = L.gf_last_error_detail().decode(errors="replace")
        super().__init__("%s (%d): %s%s" % (self.name or "gf_status", status, self.message,
                                            (" [" + detail + "]") if detail else ""))


_SENTINELS = ["ErrBufferWriteOutofRange", "ErrBufferCopyOutofRange", "ErrBufferSliceOutofRange",
              "ErrInvalidBufferData", "ErrAllocateGPUBuffer", "ErrInvalidDeviceID", "ErrInvalidFeautres",
              "ErrFeatureSetNotFound", "ErrInvalidSetState", "ErrFeatureSetExist", "ErrTooMuchGPUMemory",
              "ErrAllocatGPUMemory", "ErrSliceGPUBuffer", "ErrNotEnoughBlocks", "ErrOutOfBatch",
              "ErrMismatchDimension", "ErrWriteInputBuffer", "ErrWriteOutputBuffer", "ErrBlockIsFull",
              "ErrBlockUsed", "ErrWriteCudaBuffer", "ErrSliceBuffer", "ErrClearCudaBuffer", "ErrAllDevice",
              "ErrCreateContext", "ErrBadTransposeValue"]
# status code of every sentinel: ErrOutOfBatch == 15, ...
for _i, _n in enumerate(_SENTINELS):
    globals()[_n] = _i + 1


def _check(status):
    if status != 0:
        raise GoFeatureError(status)


# ------------------------------------------------------------------ value types (proto.go:3-26)
class Feature:
    """proto.go:12-18.  Value: little-endian bytes (dimension * precision); ID: str or bytes."""
    __slots__ = ("Value", "ID")

    def __init__(self, Value=b"", ID=""):
        self.Value = Value
        self.ID = ID


class FeatureSearchResult:
    """proto.go:20-26."""
    __slots__ = ("Score", "ID", "Slot")

    def __init__(self, Score, ID, Slot=0):
        self.Score = Score
        self.ID = ID
        self.Slot = Slot

    def __repr__(self):
        return "{%r %r}" % (self.Score, self.ID)


# ------------------------------------------------------------------ util.go
def TFeatureValue(value):
    """util.go:41-74: reinterpret []byte / []int16 / []int32 / []float32 / []float64 as bytes."""
    if isinstance(value, (bytes, bytearray, memoryview)):
        return bytes(value)
    if isinstance(value, np.ndarray):
        if value.dtype in (np.int16, np.int32, np.float32, np.float64, np.uint8):
            return np.ascontiguousarray(value).tobytes()
        raise GoFeatureError(ErrInvalidBufferData)  # noqa: F821
    if isinstance(value, (list, tuple)):
        return np.asarray(value, dtype=np.float32).tobytes()
    raise GoFeatureError(ErrInvalidBufferData)  # noqa: F821


def FeatureToFeatureValue1D(*features):
    """util.go:14-19."""
    return b"".join(bytes(f.Value) for f in features)


def FeatureValueTranspose1D(precision, *features):
    """util.go:21-39."""
    if len(features) == 0:
        return b""
    len0 = len(features[0])
    flat = b"".join(bytes(f)[:len0].ljust(len0, b"\0") for f in features)
    out = ctypes.create_string_buffer(max(1, len(features) * len0))
    _check(lib().gf_util_transpose_1d(precision, flat, len(features), len0, out))
    return out.raw[: (len0 // precision) * precision * len(features)]


def NoramlizeFloat32(feature):
    """util.go:125-139 (the reference's spelling)."""
    v = np.ascontiguousarray(feature, dtype=np.float32)
    out = np.empty_like(v)
    _check(lib().gf_util_normalize_float32(v.ctypes.data, out.ctypes.data, v.shape[0]))
    return out


def GetRandomString(length):
    """util.go:114-123."""
    alphabet = string.digits + string.ascii_lowercase + string.ascii_uppercase + "_-"
    return "".join(random.choice(alphabet) for _ in range(length))


def _pack_ids(ids):
    bs = [i.encode() if isinstance(i, str) else bytes(i) for i in ids]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    return b"".join(bs), off


def _pack_values(values, n):
    """list of FeatureValue bytes / 2-D float32 array -> one contiguous byte buffer."""
    if isinstance(values, np.ndarray):
        a = np.ascontiguousarray(values)
        return a, a.nbytes
    blob = b"".join(bytes(v) for v in values)
    return blob, len(blob)


def _ptr(buf):
    if isinstance(buf, np.ndarray):
        return buf.ctypes.data
    return ctypes.cast(ctypes.c_char_p(buf), c_vp).value if buf else None


def _decode_id(b):
    try:
        return b.decode()
    except UnicodeDecodeError:
        return b


def _take_result(handle):
    """gf_result -> [][]FeatureSearchResult, freeing the handle."""
    L = lib()
    try:
        q = L.gf_result_batch(handle)
        total = L.gf_result_total(handle)
        counts = np.zeros(max(q, 1), dtype=np.int32)
        scores = np.zeros(max(total, 1), dtype=np.float32)
        slots = np.zeros(max(total, 1), dtype=np.uint64)
        idb = ctypes.create_string_buffer(max(1, L.gf_result_id_bytes(handle)))
        ido = np.zeros(total + 1, dtype=np.uint64)
        _check(L.gf_result_export(handle, counts.ctypes.data, scores.ctypes.data, slots.ctypes.data, idb,
                                  ido.ctypes.data))
        raw = idb.raw
        ret, k = [], 0
        for i in range(q):
            row = []
            for _ in range(int(counts[i])):
                row.append(FeatureSearchResult(np.float32(scores[k]), _decode_id(raw[int(ido[k]):int(ido[k + 1])]),
                                               int(slots[k])))
                k += 1
            ret.append(row)
        return ret
    finally:
        L.gf_result_free(handle)


# ------------------------------------------------------------------ L0
def AllDevices():
    """cuda.AllDevices (search_test.go:33): list of device ordinals."""
    n = c_int(0)
    _check(lib().gf_device_count(ctypes.byref(n)))
    return list(range(n.value))


def TotalMem(device):
    """Device.TotalMem (demo/main.go:61)."""
    b = c_u64(0)
    _check(lib().gf_device_total_mem(device, ctypes.byref(b)))
    return b.value


class Context:
    """Stands in for *cuda.Context (search_test.go:40)."""

    def __init__(self, device=0, devices=None):
        h = c_vp()
        if devices is not None:
            arr = (c_int * len(devices))(*devices)
            _check(lib().gf_context_create_multi(arr, len(devices), ctypes.byref(h)))
            device = devices[0]
        else:
            _check(lib().gf_context_create(device, ctypes.byref(h)))
        self._h = h
        self.device = device
        self.devices = list(devices) if devices is not None else [device]

    def DeviceCount(self):
        return int(lib().gf_context_device_count(self._h))

    def Synchronize(self):
        _check(lib().gf_context_synchronize(self._h))

    def Stats(self):
        s = gf_stats()
        _check(lib().gf_context_stats(self._h, ctypes.byref(s)))
        return {n: int(getattr(s, n)) for n, _ in gf_stats._fields_}

    def ResetStats(self):
        _check(lib().gf_context_reset_stats(self._h))

    def SetProfiling(self, on):
        _check(lib().gf_context_set_profiling(self._h, 1 if on else 0))

    def Close(self):
        if self._h:
            lib().gf_context_destroy(self._h)
            self._h = None


def NewContext(device=0, _flags=-1):
    return Context(device)


def NewContextMulti(devices):
    """One context over several devices of this process (gf_context_create_multi): caches created on it stripe
    their blocks over the devices, Set.Search scans all of them and returns one merged, id-resolved result."""
    return Context(devices=list(devices))


# ------------------------------------------------------------------ L1 Buffer (interface.go:120-143)
class Buffer:
    def __init__(self, handle):
        self._h = handle

    def GetBuffer(self):
        return lib().gf_buffer_device_ptr(self._h)

    def Write(self, value):
        value = bytes(value) if not isinstance(value, np.ndarray) else np.ascontiguousarray(value)
        n = value.nbytes if isinstance(value, np.ndarray) else len(value)
        _check(lib().gf_buffer_write(self._h, _ptr(value), n))

    def Read(self):
        n = self.Size()
        out = ctypes.create_string_buffer(max(1, n))
        _check(lib().gf_buffer_read(self._h, out, n))
        return out.raw[:n]

    def Copy(self, src):
        _check(lib().gf_buffer_copy(self._h, src._h))

    def Reset(self):
        _check(lib().gf_buffer_reset(self._h))

    def Slice(self, start, end):
        h = c_vp()
        _check(lib().gf_buffer_slice(self._h, start, end, ctypes.byref(h)))
        return Buffer(h)

    def Size(self):
        return int(lib().gf_buffer_size(self._h))

    def Free(self):
        if self._h:
            lib().gf_buffer_free(self._h)
            self._h = None


def NewGPUBuffer(ctx, _allocator, size):
    """buffer.go:15-32 (the allocator argument has no counterpart)."""
    h = c_vp()
    _check(lib().gf_buffer_create(ctx._h, size, ctypes.byref(h)))
    return Buffer(h)


def WrapDevicePointer(ctx, ptr, size):
    h = c_vp()
    _check(lib().gf_buffer_wrap(ctx._h, ptr, size, ctypes.byref(h)))
    return Buffer(h)


# ------------------------------------------------------------------ L2 Block (interface.go:69-117)
class Block:
    def __init__(self, handle):
        self._h = handle

    def Capacity(self):
        return lib().gf_block_capacity(self._h)

    def Margin(self):
        return lib().gf_block_margin(self._h)

    def IsOwned(self):
        return bool(lib().gf_block_is_owned(self._h))

    def NextIndex(self):
        return lib().gf_block_next_index(self._h)

    def Accquire(self, _handle, owner, dims, precision, batch, _worker=None):
        """block.go:59-78 (reference spelling; cuBLAS handle and worker are unused)."""
        _check(lib().gf_block_acquire(self._h, owner.encode(), dims, precision, batch))

    def Release(self):
        _check(lib().gf_block_release(self._h))

    def Insert(self, *features):
        blob = FeatureToFeatureValue1D(*features)
        ids, off = _pack_ids([f.ID for f in features])
        _check(lib().gf_block_insert(self._h, len(features), _ptr(blob), len(blob), _ptr(ids), off.ctypes.data))

    def Delete(self, *ids):
        idb, off = _pack_ids(ids)
        flags = np.zeros(max(len(ids), 1), dtype=np.uint8)
        n = c_i64(0)
        _check(lib().gf_block_delete(self._h, len(ids), _ptr(idb), off.ctypes.data, flags.ctypes.data,
                                     ctypes.byref(n)))
        return [i for i, f in zip(ids, flags) if f] or None

    def Update(self, *features):  # block.go:169-172: TODO stub in the reference
        return None

    def Read(self, *ids):         # block.go:176-179: TODO stub in the reference
        return None

    def Search(self, inputBuffer, outputBuffer, batch, limit):
        h = c_vp()
        _check(lib().gf_block_search(self._h, inputBuffer._h, outputBuffer._h if outputBuffer else None, batch,
                                     limit, ctypes.byref(h)))
        ret = _take_result(h)
        return ret if ret else None


# ------------------------------------------------------------------ L3 Set (interface.go:37-66)
class Set:
    def __init__(self, handle, cache):
        self._h = handle
        self._cache = cache

    # -- reference API
    def Add(self, *features):
        blob = FeatureToFeatureValue1D(*features)
        ids, off = _pack_ids([f.ID for f in features])
        _check(lib().gf_set_add(self._h, len(features), _ptr(blob), len(blob), _ptr(ids), off.ctypes.data))

    def Delete(self, *ids):
        idb, off = _pack_ids(ids)
        flags = np.zeros(max(len(ids), 1), dtype=np.uint8)
        n = c_i64(0)
        _check(lib().gf_set_delete(self._h, len(ids), _ptr(idb), off.ctypes.data, flags.ctypes.data,
                                   ctypes.byref(n)))
        return [i for i, f in zip(ids, flags) if f] or None

    def Update(self, *features):
        blob = FeatureToFeatureValue1D(*features)
        ids, off = _pack_ids([f.ID for f in features])
        flags = np.zeros(max(len(features), 1), dtype=np.uint8)
        n = c_i64(0)
        _check(lib().gf_set_update(self._h, len(features), _ptr(blob), len(blob), _ptr(ids), off.ctypes.data,
                                   flags.ctypes.data, ctypes.byref(n)))
        return [f.ID for f, ok in zip(features, flags) if ok] or None

    def Read(self, *ids):
        idb, off = _pack_ids(ids)
        rowb = self.Dims() * self.Precision()
        out = np.zeros(max(len(ids) * rowb, 1), dtype=np.uint8)
        flags = np.zeros(max(len(ids), 1), dtype=np.uint8)
        n = c_i64(0)
        _check(lib().gf_set_read(self._h, len(ids), _ptr(idb), off.ctypes.data, out.ctypes.data, out.nbytes,
                                 flags.ctypes.data, ctypes.byref(n)))
        feats = [Feature(out[i * rowb:(i + 1) * rowb].tobytes(), ids[i]) for i in range(len(ids)) if flags[i]]
        return feats or None

    def Destroy(self):
        _check(lib().gf_set_destroy(self._h))

    def Search(self, threshold, limit, *features):
        """set.go:118-159: features are FeatureValues (bytes); returns [][]FeatureSearchResult."""
        blob = b"".join(bytes(f) for f in features)
        h = c_vp()
        _check(lib().gf_set_search(self._h, float(threshold), int(limit), len(features), _ptr(blob), len(blob),
                                   ctypes.byref(h)))
        return _take_result(h)

    # -- array-friendly variants used by the bench and the bulk tests
    def AddArray(self, values, ids):
        a = np.ascontiguousarray(values, dtype=np.float32)
        idb, off = _pack_ids(ids)
        _check(lib().gf_set_add(self._h, len(ids), a.ctypes.data, a.nbytes, _ptr(idb), off.ctypes.data))

    def SearchArray(self, threshold, limit, queries):
        a = np.ascontiguousarray(queries, dtype=np.float32)
        a = a.reshape(-1, a.shape[-1]) if a.ndim > 1 else a.reshape(1, -1)
        h = c_vp()
        _check(lib().gf_set_search(self._h, float(threshold), int(limit), a.shape[0], a.ctypes.data, a.nbytes,
                                   ctypes.byref(h)))
        return _take_result(h)

    def AddSynthetic(self, n, seed, first_ordinal=0, normalize=True, id_prefix="f", centres=0, spread=0.0, dup_ppm=0):
        if centres or dup_ppm:
            _check(lib().gf_set_add_synthetic_ex(self._h, n, seed, first_ordinal, 1 if normalize else 0,
                                                 id_prefix.encode(), centres, spread, dup_ppm))
            return
        _check(lib().gf_set_add_synthetic(self._h, n, seed, first_ordinal, 1 if normalize else 0,
                                          id_prefix.encode()))

    def SearchDevice(self, limit, q, d_queries_ptr, d_out_keys_ptr, path=0, stream=None):
        _check(lib().gf_set_search_device(self._h, limit, q, d_queries_ptr, d_out_keys_ptr, path, stream))

    def SearchDeviceEx(self, limit, q, d_queries_ptr, d_out_keys_ptr, d_out_flags_ptr, path=0, stream=None):
        """gf_set_search_device_ex: d_out_flags[i] = 1 where the certificate refused query i (re-run with path=1)."""
        _check(lib().gf_set_search_device_ex(self._h, limit, q, d_queries_ptr, d_out_keys_ptr, d_out_flags_ptr, path,
                                             stream))

    def DeviceCounters(self):
        """(queries the certificate refused, queries the flag-less device call left unrepaired) so far."""
        a, b = c_u64(0), c_u64(0)
        _check(lib().gf_set_device_counters(self._h, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value), int(b.value)

    def ResolveKeys(self, threshold, limit, keys):
        k = np.ascontiguousarray(keys, dtype=np.uint64).reshape(-1, limit)
        h = c_vp()
        _check(lib().gf_set_resolve_keys(self._h, float(threshold), limit, k.shape[0], k.ctypes.data,
                                         ctypes.byref(h)))
        return _take_result(h)

    def Compact(self):
        """Pack live rows, shrink NextIndex (not in the reference).  Returns rows reclaimed."""
        n = c_i64(0)
        _check(lib().gf_set_compact(self._h, ctypes.byref(n)))
        return n.value

    def Export(self):
        """All live rows in slot order: (float32 array [rows, dims], list of ids).  SURVEY 8f N3."""
        n, nb = c_i64(0), c_u64(0)
        _check(lib().gf_set_export(self._h, None, 0, None, 0, None, ctypes.byref(n), ctypes.byref(nb)))
        vals = np.zeros((max(n.value, 1), self.Dims()), dtype=np.float32)
        idb = ctypes.create_string_buffer(max(1, nb.value))
        off = np.zeros(n.value + 1, dtype=np.uint64)
        _check(lib().gf_set_export(self._h, vals.ctypes.data, vals.nbytes, idb, nb.value, off.ctypes.data,
                                   ctypes.byref(n), ctypes.byref(nb)))
        raw = idb.raw
        return vals[:n.value], [_decode_id(raw[int(off[i]):int(off[i + 1])]) for i in range(n.value)]

    def SetShard(self, rank, world):
        _check(lib().gf_set_set_shard(self._h, rank, world))

    def SetCoalescing(self, max_wait_us):
        _check(lib().gf_set_set_coalescing(self._h, max_wait_us))

    def SetSearchPath(self, path):
        """0 choose, 1 fused scan, 2 tensor path, 3 tensor path on the float32 rows (kind::tf32)."""
        _check(lib().gf_set_set_search_path(self._h, path))

    def Dims(self):
        return lib().gf_set_dims(self._h)

    def Precision(self):
        return lib().gf_set_precision(self._h)

    def Batch(self):
        return lib().gf_set_batch(self._h)

    def BlockCount(self):
        return lib().gf_set_block_count(self._h)

    def ScannedRows(self):
        return int(lib().gf_set_scanned_rows(self._h))

    def LiveRows(self):
        return int(lib().gf_set_live_rows(self._h))


# ------------------------------------------------------------------ L4 Cache (interface.go:12-34)
class Cache:
    def __init__(self, handle, ctx):
        self._h = handle
        self._ctx = ctx

    def NewSet(self, name, dims, precision, batch):
        _check(lib().gf_cache_new_set(self._h, name.encode(), dims, precision, batch))

    def DestroySet(self, name):
        _check(lib().gf_cache_destroy_set(self._h, name.encode()))

    def GetSet(self, name):
        h = c_vp()
        _check(lib().gf_cache_get_set(self._h, name.encode(), ctypes.byref(h)))
        return Set(h, self)

    def GetBlockSize(self):
        return int(lib().gf_cache_get_block_size(self._h))

    def HasShadow(self):
        return bool(lib().gf_cache_has_shadow(self._h))

    def ShadowBytesPerElement(self):
        """0 = no shadow, 1 = int8 (+ one float scale per row), 2 = bf16."""
        return int(lib().gf_cache_has_shadow(self._h))

    def GetEmptyBlock(self, blocknum):
        arr = (c_vp * max(blocknum, 1))()
        _check(lib().gf_cache_get_empty_block(self._h, blocknum, arr))
        return [Block(c_vp(arr[i])) for i in range(blocknum)]

    def AllBlocks(self):
        out = []
        for i in range(lib().gf_cache_block_count(self._h)):
            h = c_vp()
            _check(lib().gf_cache_block_at(self._h, i, ctypes.byref(h)))
            out.append(Block(h))
        return out

    def Close(self):
        if self._h:
            lib().gf_cache_destroy(self._h)
            self._h = None


CACHE_NO_SHADOW = 1
CACHE_REQUIRE_SHADOW = 2
CACHE_SHADOW_BF16 = 4


def NewCache(ctx, blockNum, blockSize, shadow=None):
    """cache.go:21-43.  shadow: None = library default (an int8 shadow slab unless GF_SHADOW says otherwise),
    False = none, True / "int8" = int8 required, "bf16" = bf16 required."""
    h = c_vp()
    if shadow is None:
        _check(lib().gf_cache_create(ctx._h, blockNum, blockSize, ctypes.byref(h)))
    else:
        if shadow is False or shadow == "none":
            flags = CACHE_NO_SHADOW
        elif shadow == "bf16":
            flags = CACHE_REQUIRE_SHADOW | CACHE_SHADOW_BF16
        else:
            flags = CACHE_REQUIRE_SHADOW
        _check(lib().gf_cache_create_ex(ctx._h, blockNum, blockSize, flags, ctypes.byref(h)))
    return Cache(h, ctx)


# ------------------------------------------------------------------ keys
def key_score(key):
    return np.float32(lib().gf_key_score(int(key)))


def key_slot(key):
    return int(lib().gf_key_slot(int(key)))


def key_make(score, slot):
    return int(lib().gf_key_make(float(score), int(slot)))
