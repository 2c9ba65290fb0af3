/*
 * gofeature_b200.h -- C ABI of libgofeature_b200.so
 *
 * Drop-in boundary for goFeature's brute-force search path.  In the reference the
 * Go package reaches the GPU through two un-vendored cgo modules,
 * github.com/unixpickle/cuda (context thread, alloc/memcpy/memset) and
 * github.com/unixpickle/cuda/cublas (one Sgemm per block per search,
 * /root/reference/block.go:196-209), then copies the whole score matrix back and
 * sorts it on the CPU (block.go:214, util.go:76-93).  This library replaces those
 * bindings AND the host logic above them (block.go, buffer.go, cache.go, set.go,
 * util.go) with one coarse-grained C interface: plain pointers and sizes in, caller
 * owned buffers or opaque handles out, int status codes mapped 1:1 on error.go.
 * INTEGRATION.md shows the cgo stubs that bind it under the reference's own Go
 * interfaces (interface.go:12-143).
 *
 * Conventions
 *   - every function returns a gf_status (0 = GF_OK) unless documented otherwise;
 *     gf_last_error_detail() gives a thread-local human string (CUDA error text);
 *   - inputs are borrowed for the duration of the call only (cgo rule: no Go pointer
 *     is retained); outputs are written into caller buffers or returned as handles
 *     the caller frees;
 *   - all entry points are thread-safe; gf_set_search is re-entrant from many OS
 *     threads (the reference serialises everything on one context thread);
 *   - FeatureValue bytes are little-endian float32 (proto.go:3-4), feature IDs are
 *     arbitrary non-empty byte strings (proto.go:6-7) passed as one concatenated
 *     byte array plus n+1 offsets;
 *   - there is NO CPU fallback: without a CUDA device gf_context_create fails.
 */
#ifndef GOFEATURE_B200_H
#define GOFEATURE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GF_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ status */
/* 1..26 are the sentinels of /root/reference/error.go:9-49, in file order. */
typedef enum gf_status {
    GF_OK = 0,
    GF_ERR_BUFFER_WRITE_OUT_OF_RANGE = 1, /* ErrBufferWriteOutofRange error.go:11 */
    GF_ERR_BUFFER_COPY_OUT_OF_RANGE = 2,  /* ErrBufferCopyOutofRange  error.go:12 */
    GF_ERR_BUFFER_SLICE_OUT_OF_RANGE = 3, /* ErrBufferSliceOutofRange error.go:13 */
    GF_ERR_INVALID_BUFFER_DATA = 4,       /* ErrInvalidBufferData     error.go:14 */
    GF_ERR_ALLOCATE_GPU_BUFFER = 5,       /* ErrAllocateGPUBuffer     error.go:15 */
    GF_ERR_INVALID_DEVICE_ID = 6,         /* ErrInvalidDeviceID       error.go:18 */
    GF_ERR_INVALID_FEATURES = 7,          /* ErrInvalidFeautres       error.go:19 */
    GF_ERR_FEATURE_SET_NOT_FOUND = 8,     /* ErrFeatureSetNotFound    error.go:20 */
    GF_ERR_INVALID_SET_STATE = 9,         /* ErrInvalidSetState       error.go:21 */
    GF_ERR_FEATURE_SET_EXIST = 10,        /* ErrFeatureSetExist       error.go:24 */
    GF_ERR_TOO_MUCH_GPU_MEMORY = 11,      /* ErrTooMuchGPUMemory      error.go:25 */
    GF_ERR_ALLOCAT_GPU_MEMORY = 12,       /* ErrAllocatGPUMemory      error.go:26 */
    GF_ERR_SLICE_GPU_BUFFER = 13,         /* ErrSliceGPUBuffer        error.go:27 */
    GF_ERR_NOT_ENOUGH_BLOCKS = 14,        /* ErrNotEnoughBlocks       error.go:28 */
    GF_ERR_OUT_OF_BATCH = 15,             /* ErrOutOfBatch            error.go:31 */
    GF_ERR_MISMATCH_DIMENSION = 16,       /* ErrMismatchDimension     error.go:32 */
    GF_ERR_WRITE_INPUT_BUFFER = 17,       /* ErrWriteInputBuffer      error.go:33 */
    GF_ERR_WRITE_OUTPUT_BUFFER = 18,      /* ErrWriteOutputBuffer     error.go:34 */
    GF_ERR_BLOCK_IS_FULL = 19,            /* ErrBlockIsFull           error.go:37 */
    GF_ERR_BLOCK_USED = 20,               /* ErrBlockUsed             error.go:38 */
    GF_ERR_WRITE_CUDA_BUFFER = 21,        /* ErrWriteCudaBuffer       error.go:41 */
    GF_ERR_SLICE_BUFFER = 22,             /* ErrSliceBuffer           error.go:42 */
    GF_ERR_CLEAR_CUDA_BUFFER = 23,        /* ErrClearCudaBuffer       error.go:43 */
    GF_ERR_ALL_DEVICE = 24,               /* ErrAllDevice             error.go:44 */
    GF_ERR_CREATE_CONTEXT = 25,           /* ErrCreateContext         error.go:45 */
    GF_ERR_BAD_TRANSPOSE_VALUE = 26,      /* ErrBadTransposeValue     error.go:48 */
    /* not in error.go: conditions the reference turns into a panic or a raw CUDA error */
    GF_ERR_INVALID_ARGUMENT = 100,
    GF_ERR_CUDA = 101,           /* a CUDA runtime call failed; see gf_last_error_detail */
    GF_ERR_READ_CUDA_BUFFER = 102, /* "fail to read buffer vec3" block.go:216 */
    GF_ERR_UNSUPPORTED = 103     /* e.g. precision != 4 on the search path */
} gf_status;

/* The exact message of the matching error.go sentinel (or a description for >= 100). */
GF_API const char *gf_strerror(int status);
/* Name of the Go variable ("ErrOutOfBatch"); "" for codes >= 100 and GF_OK. */
GF_API const char *gf_status_name(int status);
GF_API const char *gf_last_error_detail(void);
GF_API const char *gf_version(void);

/* ------------------------------------------------------------------ handles */
typedef struct gf_context gf_context; /* replaces *cuda.Context   (cache.go:21, search_test.go:40) */
typedef struct gf_buffer gf_buffer;   /* GPUBuffer                (buffer.go:8-11)               */
typedef struct gf_cache gf_cache;     /* _Cache                   (cache.go:12-19)               */
typedef struct gf_block gf_block;     /* _Block                   (block.go:14-37)               */
typedef struct gf_set gf_set;         /* FeatureSet               (set.go:24-38)                 */
typedef struct gf_result gf_result;   /* [][]FeatureSearchResult  (proto.go:20-26)               */

/* ------------------------------------------------------- host utilities (util.go) */
/* NoramlizeFloat32 (util.go:125-139): sequential float32 sum of squares, square root through
 * float64 pow(x, 0.5), per-element float32 divide; an all-zero vector gives zeros. */
GF_API int gf_util_normalize_float32(const float *in, float *out, int64_t n);
/* FeatureValueTranspose1D (util.go:21-39): `row` values of `len0` bytes each, concatenated in
 * `features`; ret[(i*row + j)*precision ...] = element i of feature j.  len0 % precision != 0
 * -> ErrBadTransposeValue.  `ret` holds row*len0 bytes. */
GF_API int gf_util_transpose_1d(int precision, const void *features, int64_t row, int64_t len0, void *ret);

/* ------------------------------------------------------- L0: device + context */
/* cuda.AllDevices (search_test.go:33, demo/main.go:49): number of CUDA devices. */
GF_API int gf_device_count(int *out_count);
/* Device.TotalMem (demo/main.go:61). */
GF_API int gf_device_total_mem(int device, uint64_t *out_bytes);
/* cuda.NewContext(devices[i], -1) (search_test.go:40).  Fails with GF_ERR_CREATE_CONTEXT /
 * GF_ERR_INVALID_DEVICE_ID; never falls back to the CPU. */
GF_API int gf_context_create(int device, gf_context **out_ctx);
/* One context over several devices of THIS process (north_star: a set row-sharded over the GPUs of one box, behind
 * the same Cache / Set API).  Caches created on it stripe their blocks -- block i on devices[i % n], the order
 * Cache.GetEmptyBlock hands them out (cache.go:104-115) -- so a growing set spreads evenly; slots in results stay
 * global ((block position in Set.Blocks) * capacity + row).  gf_set_search launches every device's scan from the
 * calling thread; the per-device winners meet in devices[0] over peer memory (one kernel, no collective library)
 * and the merged, id-resolved result comes back exactly as from a one-device set.  Buffers and bare blocks created
 * through this context live on devices[0].  A device may be listed more than once (its shards share it).  The device-resident entry points (gf_set_search_device*) address one
 * device and return GF_ERR_UNSUPPORTED on such a set. */
GF_API int gf_context_create_multi(const int *devices, int n, gf_context **out_ctx);
/* Devices behind this context (1 for gf_context_create). */
GF_API int gf_context_device_count(const gf_context *ctx);
GF_API int gf_context_destroy(gf_context *ctx);
GF_API int gf_context_device(const gf_context *ctx);
/* Block the caller until all work submitted through this context has finished. */
GF_API int gf_context_synchronize(gf_context *ctx);

/* ------------------------------------------------- L1: Buffer (interface.go:120-143) */
/* NewGPUBuffer (buffer.go:15-32): allocate `size` bytes of device memory and zero it. */
GF_API int gf_buffer_create(gf_context *ctx, uint64_t size, gf_buffer **out_buf);
/* Wrap caller-owned device memory (e.g. a torch tensor) without taking ownership. */
GF_API int gf_buffer_wrap(gf_context *ctx, void *device_ptr, uint64_t size, gf_buffer **out_buf);
/* Buffer.Write (buffer.go:34-44): H2D of `len` bytes at offset 0; len > Size -> ErrBufferWriteOutofRange. */
GF_API int gf_buffer_write(gf_buffer *buf, const void *value, uint64_t len);
/* Buffer.Read (buffer.go:46-52): D2H of the whole buffer; `cap` must be >= Size. */
GF_API int gf_buffer_read(gf_buffer *buf, void *out, uint64_t cap);
/* Buffer.Copy (buffer.go:54-62): D2D of src into dst; src.Size > dst.Size -> ErrBufferCopyOutofRange
 * (the reference assigns that error and then overwrites it, buffer.go:55-60; here it is returned). */
GF_API int gf_buffer_copy(gf_buffer *dst, const gf_buffer *src);
/* Buffer.Reset (buffer.go:66-71): memset 0. */
GF_API int gf_buffer_reset(gf_buffer *buf);
/* Buffer.Slice (buffer.go:73-85): [start, end) view aliasing the parent. */
GF_API int gf_buffer_slice(gf_buffer *buf, int64_t start, int64_t end, gf_buffer **out_buf);
/* Buffer.Size (buffer.go:64). */
GF_API uint64_t gf_buffer_size(const gf_buffer *buf);
/* Buffer.GetBuffer (buffer.go:32): the raw device pointer. */
GF_API void *gf_buffer_device_ptr(const gf_buffer *buf);
/* Drop the handle (device memory of an owning root buffer is freed with its last view). */
GF_API int gf_buffer_free(gf_buffer *buf);

/* ------------------------------------------------- L4: Cache (interface.go:12-34) */
/* NewCache(ctx, blockNum, blockSize) (cache.go:21-43): one zeroed slab cut into blocks. */
GF_API int gf_cache_create(gf_context *ctx, int block_num, int64_t block_size, gf_cache **out_cache);
/* gf_cache_create with options.  By default the cache also keeps a SHADOW of the slab: the rows
 * quantised to int8 with one float scale per row (block_size / 4 + block_size / 128 more bytes per
 * block; GF_CACHE_SHADOW_BF16: bf16 instead, block_size / 2 more bytes), converted on the device
 * whenever Add / Update / Delete / Compact writes rows.  The tensor-core candidate pass then
 * streams 1 (or 2) bytes per element instead of 4; every candidate is still re-scored from the
 * float32 rows and certified against the MEASURED quantisation residuals, so results are identical
 * with and without it.  GF_CACHE_NO_SHADOW (environment for gf_cache_create: GF_SHADOW=0, =bf16)
 * turns it off; GF_CACHE_REQUIRE_SHADOW fails instead of continuing without when the device cannot
 * hold it.  gf_cache_has_shadow: 0 none, 1 int8, 2 bf16. */
#define GF_CACHE_NO_SHADOW 1u
#define GF_CACHE_REQUIRE_SHADOW 2u
#define GF_CACHE_SHADOW_BF16 4u
GF_API int gf_cache_create_ex(gf_context *ctx, int block_num, int64_t block_size, unsigned flags,
                              gf_cache **out_cache);
GF_API int gf_cache_has_shadow(const gf_cache *cache);
/* Releases every set, the slab and all handles obtained from this cache. */
GF_API int gf_cache_destroy(gf_cache *cache);
/* Cache.NewSet (cache.go:45-72). */
GF_API int gf_cache_new_set(gf_cache *cache, const char *name, int dims, int precision, int batch);
/* Cache.DestroySet (cache.go:74-90). */
GF_API int gf_cache_destroy_set(gf_cache *cache, const char *name);
/* Cache.GetSet (cache.go:92-100).  The handle is owned by the cache. */
GF_API int gf_cache_get_set(gf_cache *cache, const char *name, gf_set **out_set);
/* Cache.GetBlockSize (cache.go:102). */
GF_API int64_t gf_cache_get_block_size(const gf_cache *cache);
/* Cache.GetEmptyBlock (cache.go:104-115): the first `block_num` un-owned blocks, in AllBlocks order. */
GF_API int gf_cache_get_empty_block(gf_cache *cache, int block_num, gf_block **out_blocks);
/* AllBlocks[i] (cache.go:14); handle owned by the cache. */
GF_API int gf_cache_block_count(const gf_cache *cache);
GF_API int gf_cache_block_at(gf_cache *cache, int index, gf_block **out_block);

/* ------------------------------------------------- L2: Block (interface.go:69-117) */
GF_API int gf_block_capacity(const gf_block *block);  /* block.go:52 (0 when un-owned) */
GF_API int gf_block_margin(const gf_block *block);    /* block.go:54-57 */
GF_API int gf_block_is_owned(const gf_block *block);  /* block.go:50 */
GF_API int gf_block_next_index(const gf_block *block); /* NextIndex, block.go:30 */
/* Block.Accquire (block.go:59-78).  The cuBLAS handle and the worker goroutine of the
 * reference signature have no counterpart: searches are not queued per block. */
GF_API int gf_block_acquire(gf_block *block, const char *owner, int dims, int precision, int batch);
/* Block.Release (block.go:253-270). */
GF_API int gf_block_release(gf_block *block);
/* Block.Insert (block.go:80-131): `values` = FeatureToFeatureValue1D output (util.go:14-19),
 * n*dims*precision bytes; ids concatenated with n+1 offsets. */
GF_API int gf_block_insert(gf_block *block, int64_t n, const void *values, uint64_t values_len, const char *ids,
                           const uint64_t *id_offsets);
/* Block.Delete (block.go:135-165): out_deleted[i] = 1 if ids[i] was found and removed. */
GF_API int gf_block_delete(gf_block *block, int64_t n, const char *ids, const uint64_t *id_offsets,
                           uint8_t *out_deleted, int64_t *out_ndeleted);
/* Block.Search (block.go:184-249): `input` holds `batch` queries in the dims-major layout
 * FeatureValueTranspose1D produces (util.go:21-39, element l of query q at (l*batch+q)*4);
 * `output` is the reference's score scratch and is not touched (scores never leave the SM).
 * Returns per query the top-`limit` over rows [0, NextIndex) with tombstones dropped AFTER
 * selection (block.go:238-243).  NextIndex == 0 -> a result with batch 0 (block.go:186-189). */
GF_API int gf_block_search(gf_block *block, const gf_buffer *input, gf_buffer *output, int batch, int limit,
                           gf_result **out_result);

/* ------------------------------------------------- L3: Set (interface.go:37-66) */
/* Set.Add (set.go:77-116). */
GF_API int gf_set_add(gf_set *set, int64_t n, const void *values, uint64_t values_len, const char *ids,
                      const uint64_t *id_offsets);
/* Set.Delete (set.go:163-193). */
GF_API int gf_set_delete(gf_set *set, int64_t n, const char *ids, const uint64_t *id_offsets, uint8_t *out_deleted,
                         int64_t *out_ndeleted);
/* Set.Update (interface.go:47-50; an empty TODO in set.go:195-198): overwrite the value of
 * existing ids in place; out_updated[i] = 1 where found. */
GF_API int gf_set_update(gf_set *set, int64_t n, const void *values, uint64_t values_len, const char *ids,
                         const uint64_t *id_offsets, uint8_t *out_updated, int64_t *out_nupdated);
/* Set.Read (interface.go:52-55; TODO in set.go:216-219): out_values gets n*dims*precision bytes,
 * rows of ids that are not found are zero and out_found[i] = 0. */
GF_API int gf_set_read(gf_set *set, int64_t n, const char *ids, const uint64_t *id_offsets, void *out_values,
                       uint64_t out_values_cap, uint8_t *out_found, int64_t *out_nfound);
/* Set.Search (set.go:118-159): `queries` = q concatenated FeatureValues (q*dims*precision bytes).
 * q > batch -> ErrOutOfBatch (set.go:119-122).  The result always has q entries. */
GF_API int gf_set_search(gf_set *set, float threshold, int limit, int q, const void *queries, uint64_t queries_len,
                         gf_result **out_result);
/* Set.Destroy (set.go:202-214); normally reached through gf_cache_destroy_set. */
GF_API int gf_set_destroy(gf_set *set);

/* Not in the reference: NextIndex never shrinks there (block.go:159-160), so deleted rows are
 * scanned for ever.  Packs the live rows of every block to its front, in order, and shrinks NextIndex;
 * out_reclaimed_rows = rows a search no longer scans.  Slot order of live rows is preserved. */
GF_API int gf_set_compact(gf_set *set, int64_t *out_reclaimed_rows);

/* Bulk export (the reference has no persistence): every live row in slot order, with its id.
 * Call once with NULL buffers to learn out_rows / out_id_bytes, then with buffers of rows*dims*precision
 * bytes, out_id_bytes bytes and rows+1 offsets.  Re-import with gf_set_add. */
GF_API int gf_set_export(gf_set *set, void *out_values, uint64_t values_cap, char *out_ids, uint64_t ids_cap,
                         uint64_t *out_id_offsets, int64_t *out_rows, uint64_t *out_id_bytes);

GF_API int gf_set_dims(const gf_set *set);
GF_API int gf_set_precision(const gf_set *set);
GF_API int gf_set_batch(const gf_set *set);
GF_API int gf_set_block_count(const gf_set *set);
/* Sum of NextIndex over the set's blocks = rows one search scans (tombstones included). */
GF_API int64_t gf_set_scanned_rows(const gf_set *set);
GF_API int64_t gf_set_live_rows(const gf_set *set);

/* ------------------------------------------------- results ([][]FeatureSearchResult) */
GF_API int gf_result_batch(const gf_result *r);                 /* len(ret) */
GF_API int gf_result_count(const gf_result *r, int q);          /* len(ret[q]) */
GF_API float gf_result_score(const gf_result *r, int q, int i); /* ret[q][i].Score */
/* ret[q][i].ID; pointer valid until gf_result_free. */
GF_API const char *gf_result_id(const gf_result *r, int q, int i, uint32_t *out_len);
/* (block position in Set.Blocks) * capacity + row of the hit; the canonical tie-break key. */
GF_API uint64_t gf_result_slot(const gf_result *r, int q, int i);
/* Flat export for one cgo crossing: counts[q], then per hit score / slot / id bytes.
 * Sizes: gf_result_total() hits, gf_result_id_bytes() bytes of ids. */
GF_API int64_t gf_result_total(const gf_result *r);
GF_API int64_t gf_result_id_bytes(const gf_result *r);
GF_API int gf_result_export(const gf_result *r, int32_t *out_counts, float *out_scores, uint64_t *out_slots,
                            char *out_ids, uint64_t *out_id_offsets /* total+1 */);
GF_API int gf_result_free(gf_result *r);

/* ------------------------------------------------- device-resident entry points
 * Used by bench.py (`value`: inputs already in HBM), by the multi-GPU shard merge and by
 * callers that keep queries on the device.  `stream` is a cudaStream_t (NULL = the CUDA legacy
 * default stream, which is also torch's default).  Nothing here synchronises with the host;
 * calls on one set must be stream-ordered (they share one scratch arena). */

/* A candidate is one u64 key: (orderable image of the float32 score) << 32 | (0xFFFFFFFF - slot),
 * larger key = better hit (score descending, slot ascending); 0 = no candidate. */
GF_API uint64_t gf_key_make(float score, uint32_t slot);
GF_API float gf_key_score(uint64_t key);
GF_API uint32_t gf_key_slot(uint64_t key);

/* Row-shard identity of this set inside a `world`-way sharded set: block position p of this
 * shard is global block p*world + rank, and slots in keys/results are global.  Default (0,1). */
GF_API int gf_set_set_shard(gf_set *set, int rank, int world);

/* Top-`limit` over all scanned rows (tombstones compete with score 0.0 exactly like
 * block.go:236-243 lets them) for q row-major float32 queries resident on the device.
 * d_out_keys: q*limit keys, descending, zero padded.  path: 0 = choose, 1 = fused
 * scan kernel, 2 = tcgen05 contraction + exact rescoring (bf16 shadow when the cache has one),
 * 3 = the same on the float32 rows (kind::tf32). */
GF_API int gf_set_search_device(gf_set *set, int limit, int q, const float *d_queries, uint64_t *d_out_keys,
                                int path, void *stream);
/* gf_set_search_device with per-query flags.  With d_out_flags (q words) the library does NO device-side
 * repair: d_out_flags[i] = 1 where the tensor path's certificate refused query i; the keys of a flagged query
 * are not guaranteed exact and the caller re-runs it with path = 1 (gf_set_search does exactly that).
 * Without flags (NULL, = gf_set_search_device) up to 8 refused queries per call are re-run on the device by
 * the exact scan; any further ones keep their candidate-pass keys and are counted (gf_set_device_counters:
 * out_unrepaired) -- callers that cannot tolerate that pass flags. */
GF_API int gf_set_search_device_ex(gf_set *set, int limit, int q, const float *d_queries, uint64_t *d_out_keys,
                                   uint32_t *d_out_flags, int path, void *stream);

/* Device-side counters of this set since it was created (a blocking 16-byte read): queries the tensor
 * path's certificate refused (whoever repaired them), and queries the flag-less gf_set_search_device left
 * unrepaired (more than 8 refusals in one call).  bench.py reads them after its timed region. */
GF_API int gf_set_device_counters(gf_set *set, uint64_t *out_flagged, uint64_t *out_unrepaired);

/* ---- Peer exchange (multi-GPU, one process per GPU): the cross-rank step of a sharded search as ONE
 * kernel over NVLink peer memory instead of ncclAllGather + merge.  Every rank creates an exchange,
 * publishes the 64-byte CUDA IPC handle of its inbox (gf_exchange_handle), receives everybody's
 * handles (any transport: the tests use torch.distributed) and connects.  gf_exchange_merge_device is
 * collective: all ranks call it once per search, in the same order, each with its own [q][limit]
 * winners; every rank gets the merged top-`limit` per query.  d_out_flags[q] (optional) = OR of the
 * ranks' d_local_flags, 0xFFFFFFFF if a peer did not arrive within 5 s.  world * max_limit <= 2048. */
typedef struct gf_exchange gf_exchange;
GF_API int gf_exchange_create(gf_context *ctx, int rank, int world, int max_q, int max_limit, gf_exchange **out);
GF_API int gf_exchange_handle(gf_exchange *x, void *out_handle, uint64_t cap);           /* 64 bytes */
GF_API int gf_exchange_connect(gf_exchange *x, const void *handles, uint64_t len);        /* world * 64 bytes, rank order */
GF_API int gf_exchange_merge_device(gf_exchange *x, int q, int limit, const uint64_t *d_local_keys,
                                    const uint32_t *d_local_flags, uint64_t *d_out_keys, uint32_t *d_out_flags,
                                    void *stream);
/* A device-resident search of this rank's shard that ENDS with the exchange: gf_set_search_device_ex followed by
 * gf_exchange_merge_device, as one collective call (all ranks, same order).  When the search takes the one-pass chain
 * its last kernel pushes the winners to the peers, waits for theirs and folds them -- no separate launch; otherwise
 * the exchange kernel is launched behind it.  d_local_keys [q][limit] / d_local_flags [q] receive this rank's own
 * winners and flags, d_out_keys / d_out_flags the merged result (flags as in gf_exchange_merge_device). */
GF_API int gf_set_search_device_xchg(gf_set *set, gf_exchange *x, int limit, int q, const float *d_queries,
                                     uint64_t *d_local_keys, uint32_t *d_local_flags, uint64_t *d_out_keys,
                                     uint32_t *d_out_flags, int path, void *stream);
/* Where the cross-rank step of the LAST fused search spent its time (needs gf_context_set_profiling on while it
 * ran; blocking): out_ns[4 * i .. 4 * i + 3] = device clock (ns) when query i's CTA entered the exchange, had pushed
 * and published its winners, had seen every rank's, had folded them. */
GF_API int gf_exchange_trace(gf_exchange *x, int q, uint64_t *out_ns);
GF_API int gf_exchange_destroy(gf_exchange *x);

/* Merge `lists` candidate lists (each [q][limit], e.g. the all-gathered per-rank winners)
 * into the top-`limit` per query.  d_in: [lists][q][limit]; d_out: [q][limit]. */
GF_API int gf_merge_keys_device(gf_context *ctx, int lists, int q, int limit, const uint64_t *d_in, uint64_t *d_out,
                                void *stream);
/* Host-side post-processing of merged keys: drops tombstones and score < threshold, resolves
 * IDs for slots this shard owns (others get an empty id), builds a gf_result. */
GF_API int gf_set_resolve_keys(gf_set *set, float threshold, int limit, int q, const uint64_t *h_keys,
                               gf_result **out_result);

/* Append n synthetic rows generated ON the device (100 M rows never exist on the host):
 * value(seed,row,col) = U(-1,1) from a counter hash (search_test.go:65 shape), optionally
 * L2-normalised; ids are "<prefix>%011d" of first_ordinal+i, stored implicitly. */
GF_API int gf_set_add_synthetic(gf_set *set, int64_t n, uint64_t seed, int64_t first_ordinal, int normalize,
                                const char *id_prefix);

/* gf_set_add_synthetic with structure (what a face-feature index holds: many near neighbours and re-enrolled
 * duplicates): centres > 0 makes row(o) = centre(c(o)) + spread * U(-1,1)^dims, the cluster c(o) a hash of the ordinal
 * (members are scattered over the set); dup_ppm rows per million are exact copies of an earlier ordinal's row.
 * spread 0.228 gives members at cosine ~0.975 to their centre (~0.95 to each other) at 512-d.  Bit-identical to
 * oracle/gf_oracle.c gf_oracle_synth_rows_ex. */
GF_API int gf_set_add_synthetic_ex(gf_set *set, int64_t n, uint64_t seed, int64_t first_ordinal, int normalize,
                                   const char *id_prefix, uint64_t centres, float spread, uint32_t dup_ppm);

/* counters for the bench harness */
typedef struct gf_stats {
    uint64_t searches;          /* gf_set_search + gf_set_search_device calls */
    uint64_t scan_launches;     /* fused scan kernel launches */
    uint64_t gemm_launches;     /* tcgen05 contraction kernel launches */
    uint64_t merge_launches;    /* merge / rescore kernel launches */
    uint64_t other_launches;    /* fill, gather, transpose ... */
    uint64_t rows_scanned;      /* rows x passes */
    uint64_t quirk_fallbacks;   /* searches that needed the per-block tombstone path */
    uint64_t coalesced_batches; /* scan passes that served more than one caller */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    uint64_t scan_kernel_ns;    /* CUDA-event time of the timed scan launches (profiling on) */
    uint64_t scan_kernel_timed; /* how many scan launches were timed */
    uint64_t tensor_passes;     /* tcgen05 candidate passes (each serves up to 256 queries) */
    uint64_t uncertified_queries; /* tensor-path queries the host API re-ran through the exact scan */
    uint64_t gemm_kernel_ns;    /* CUDA-event time of the timed tcgen05 COLLECT launches (profiling on) */
    uint64_t gemm_kernel_timed;
    /* where the host-buffer search call (tensor path, one band) spends its wall time, summed over host_calls calls:
     * staging the queries, enqueueing the recorded chain, waiting for the device, building the id-resolved result */
    uint64_t host_stage_ns;
    uint64_t host_launch_ns;
    uint64_t host_wait_ns;
    uint64_t host_resolve_ns;
    uint64_t host_calls;
    /* where Set.Add spends its wall time, summed over add_rows rows: staging (host -> pinned -> device, shadow
     * quantisation; no set lock), placement (the set's exclusive lock: device-to-device copies, ids, segment table),
     * id locator (no set lock) */
    uint64_t add_stage_ns;
    uint64_t add_place_ns;
    uint64_t add_index_ns;
    uint64_t add_rows;
    /* of uncertified_queries: answered (and certified) by the second tier -- one kind::tf32 candidate pass over the
     * float32 rows for the whole compact batch -- instead of the exact scan (passes of 8 queries) */
    uint64_t tier2_queries;
} gf_stats;
GF_API int gf_context_stats(const gf_context *ctx, gf_stats *out);
GF_API int gf_context_reset_stats(gf_context *ctx);
/* Bracket every scan kernel launch with CUDA events on its own stream (bench.py's roofline
 * figure); the durations are folded into gf_stats by gf_context_stats.  Off by default. */
GF_API int gf_context_set_profiling(gf_context *ctx, int on);

/* Kernel path the host-side Search of this set takes: 0 = choose (default), 1 = fused scan,
 * 2 = tensor path, 3 = tensor path on the float32 rows (kind::tf32) even when a shadow exists,
 * 4 / 5 = 2 / 3 pinned to the two-pass chain (threshold sample + COLLECT) instead of the one-pass chain
 * small batches take by default.
 * Results do not depend on it; tests and benchmarks use it to pin one path. */
GF_API int gf_set_set_search_path(gf_set *set, int path);

/* Allow concurrent single callers of gf_set_search to be served by one scan pass
 * (the analogue of the reference's SearchQueue, set.go:36, cache.go:61). 0 = off. */
GF_API int gf_set_set_coalescing(gf_set *set, int max_wait_us);

#ifdef __cplusplus
}
#endif
#endif /* GOFEATURE_B200_H */
