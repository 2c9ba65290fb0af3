// gf_capi.cu -- extern "C" entry points declared in include/gofeature_b200.h.
//
// Thin: argument validation, string unpacking, exception containment.  All behaviour
// lives in gf_runtime.cu (host logic) and the kernels.
#include <cmath>
#include <cstring>
#include <new>

#include "gf_device.cuh"
#include "gf_runtime.h"

namespace gf {
const std::string &error_detail();
}

using namespace gf;

namespace {

struct Msg {
    int code;
    const char *name;
    const char *text;
};
// texts are the errors.New strings of error.go:9-49
const Msg kMsgs[] = {
    {GF_OK, "", "ok"},
    {GF_ERR_BUFFER_WRITE_OUT_OF_RANGE, "ErrBufferWriteOutofRange", "try to write too much data into buffer"},
    {GF_ERR_BUFFER_COPY_OUT_OF_RANGE, "ErrBufferCopyOutofRange", "try to copy too much data from source buffer"},
    {GF_ERR_BUFFER_SLICE_OUT_OF_RANGE, "ErrBufferSliceOutofRange", "slice buffer out of range"},
    {GF_ERR_INVALID_BUFFER_DATA, "ErrInvalidBufferData", "invalid buffer data"},
    {GF_ERR_ALLOCATE_GPU_BUFFER, "ErrAllocateGPUBuffer", "failed to allocate gpu buffer"},
    {GF_ERR_INVALID_DEVICE_ID, "ErrInvalidDeviceID", "invalid device id"},
    {GF_ERR_INVALID_FEATURES, "ErrInvalidFeautres", "invalid features in request"},
    {GF_ERR_FEATURE_SET_NOT_FOUND, "ErrFeatureSetNotFound", "feature set not found"},
    {GF_ERR_INVALID_SET_STATE, "ErrInvalidSetState", "invalid set state"},
    {GF_ERR_FEATURE_SET_EXIST, "ErrFeatureSetExist", "feature set is already exist"},
    {GF_ERR_TOO_MUCH_GPU_MEMORY, "ErrTooMuchGPUMemory", "use too much gpu memory"},
    {GF_ERR_ALLOCAT_GPU_MEMORY, "ErrAllocatGPUMemory", "fail to allocate gpu memory"},
    {GF_ERR_SLICE_GPU_BUFFER, "ErrSliceGPUBuffer", "fail to slice gpu buffer"},
    {GF_ERR_NOT_ENOUGH_BLOCKS, "ErrNotEnoughBlocks", "cache does not have enough blocks"},
    {GF_ERR_OUT_OF_BATCH, "ErrOutOfBatch", "requests out of batch limit"},
    {GF_ERR_MISMATCH_DIMENSION, "ErrMismatchDimension", "feature with mismatch dimension"},
    {GF_ERR_WRITE_INPUT_BUFFER, "ErrWriteInputBuffer", "failed to write input buffer"},
    {GF_ERR_WRITE_OUTPUT_BUFFER, "ErrWriteOutputBuffer", "failed to write output buffer"},
    {GF_ERR_BLOCK_IS_FULL, "ErrBlockIsFull", "block is full"},
    {GF_ERR_BLOCK_USED, "ErrBlockUsed", "block is used"},
    {GF_ERR_WRITE_CUDA_BUFFER, "ErrWriteCudaBuffer", "write to cuda buffer error"},
    {GF_ERR_SLICE_BUFFER, "ErrSliceBuffer", "fail to slice cuda buffer"},
    {GF_ERR_CLEAR_CUDA_BUFFER, "ErrClearCudaBuffer", "clear cuda buffer failed"},
    {GF_ERR_ALL_DEVICE, "ErrAllDevice", "fail to list all cuda device"},
    {GF_ERR_CREATE_CONTEXT, "ErrCreateContext", "fail to create cuda context"},
    {GF_ERR_BAD_TRANSPOSE_VALUE, "ErrBadTransposeValue", "invalid transpose value to transpose"},
    {GF_ERR_INVALID_ARGUMENT, "", "invalid argument"},
    {GF_ERR_CUDA, "", "cuda runtime error"},
    {GF_ERR_READ_CUDA_BUFFER, "", "fail to read buffer vec3"},
    {GF_ERR_UNSUPPORTED, "", "unsupported on the search path"},
};

const Msg *find_msg(int status)
{
    for (const Msg &m : kMsgs)
        if (m.code == status) return &m;
    return nullptr;
}

// concatenated ids + n+1 offsets -> strings; empty ids are invalid (an empty FeatureID marks a
// free slot, block.go:159,239)
int unpack_ids(int64_t n, const char *ids, const uint64_t *off, std::vector<std::string> &out)
{
    if (n < 0 || (n > 0 && !off)) return GF_ERR_INVALID_ARGUMENT;
    out.clear();
    out.reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i)
        if (off[i + 1] <= off[i]) return GF_ERR_INVALID_FEATURES;
    if (n > 0 && !ids) return GF_ERR_INVALID_ARGUMENT;
    for (int64_t i = 0; i < n; ++i) out.emplace_back(ids + off[i], (size_t)(off[i + 1] - off[i]));
    return GF_OK;
}

template <typename F>
int guarded(F &&f)
{
    try {
        return f();
    } catch (const std::bad_alloc &) {
        set_error_detail("out of host memory");
        return GF_ERR_INVALID_ARGUMENT;
    } catch (const std::exception &e) {
        set_error_detail(e.what());
        return GF_ERR_INVALID_ARGUMENT;
    } catch (...) {
        set_error_detail("unknown exception");
        return GF_ERR_INVALID_ARGUMENT;
    }
}

}  // namespace

extern "C" {

const char *gf_strerror(int status)
{
    const Msg *m = find_msg(status);
    return m ? m->text : "unknown status";
}
const char *gf_status_name(int status)
{
    const Msg *m = find_msg(status);
    return m ? m->name : "";
}
const char *gf_last_error_detail(void) { return error_detail().c_str(); }
const char *gf_version(void) { return "gofeature_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------ host utilities (util.go)
int gf_util_normalize_float32(const float *in, float *out, int64_t n)
{
    if (n < 0 || (n > 0 && (!in || !out))) return GF_ERR_INVALID_ARGUMENT;
    volatile float mode = 0.0f;  // volatile: keep the float32 rounding of every step (util.go:126-129)
    for (int64_t i = 0; i < n; ++i) {
        volatile float sq = in[i] * in[i];
        mode = mode + sq;
    }
    float m = (float)pow((double)mode, 0.5);  // util.go:130
    if (m == 0.0f) {
        for (int64_t i = 0; i < n; ++i) out[i] = 0.0f;  // util.go:131-133
        return GF_OK;
    }
    for (int64_t i = 0; i < n; ++i) out[i] = in[i] / m;  // util.go:135-137
    return GF_OK;
}

int gf_util_transpose_1d(int precision, const void *features, int64_t row, int64_t len0, void *ret)
{
    if (row == 0) return GF_OK;  // util.go:23-25
    if (precision <= 0 || row < 0 || len0 < 0 || !features || !ret) return GF_ERR_INVALID_ARGUMENT;
    if (len0 % precision != 0) return GF_ERR_BAD_TRANSPOSE_VALUE;  // util.go:26-28
    const int64_t col = len0 / precision;
    const uint8_t *src = (const uint8_t *)features;
    uint8_t *dst = (uint8_t *)ret;
    for (int64_t i = 0; i < col; ++i)
        for (int64_t j = 0; j < row; ++j)
            memcpy(dst + (i * row + j) * precision, src + j * len0 + i * precision, (size_t)precision);
    return GF_OK;
}

// ------------------------------------------------------------------ device / context
int gf_device_count(int *out_count)
{
    if (!out_count) return GF_ERR_INVALID_ARGUMENT;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *out_count = 0;
        return cuda_fail(e, "cudaGetDeviceCount", GF_ERR_ALL_DEVICE);
    }
    *out_count = n;
    return GF_OK;
}

int gf_device_total_mem(int device, uint64_t *out_bytes)
{
    if (!out_bytes) return GF_ERR_INVALID_ARGUMENT;
    cudaDeviceProp p;
    GF_CUDA(cudaGetDeviceProperties(&p, device), GF_ERR_INVALID_DEVICE_ID);
    *out_bytes = (uint64_t)p.totalGlobalMem;
    return GF_OK;
}

int gf_context_create_multi(const int *devices, int n, gf_context **out_ctx)
{
    if (!out_ctx || !devices || n < 1 || n > 64) return GF_ERR_INVALID_ARGUMENT;
    *out_ctx = nullptr;
    // (a device may be listed more than once: its shards then share that device -- how the 1-GPU test box
    // exercises the striping, the gather and the id resolution of a multi-device set)
    gf_context *parent = nullptr;
    int rc = gf_context_create(devices[0], &parent);
    if (rc) return rc;
    return guarded([&]() -> int {
        for (int i = 0; i < n; ++i) {
            gf_context *c = nullptr;
            int r = gf_context_create(devices[i], &c);
            if (r) {
                delete parent;
                return r;
            }
            parent->children.push_back(c);
        }
        // every device stores its winners into the root device's inbox: peer access (NVLink / NVSwitch on a B200 box)
        bool ok = true;
        for (int i = 0; i < n && ok; ++i)
            for (int j = 0; j < n; ++j) {
                if (devices[i] == devices[j]) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) != cudaSuccess || !can) {
                    ok = false;
                    break;
                }
                cudaSetDevice(devices[i]);
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
                cudaGetLastError();
            }
        parent->peer_ok = ok || n == 1;  // without it a search still works: the shards answer one by one (generic path)
        cudaSetDevice(devices[0]);
        *out_ctx = parent;
        return GF_OK;
    });
}

int gf_context_device_count(const gf_context *ctx) { return ctx ? (ctx->is_multi() ? (int)ctx->children.size() : 1) : 0; }

int gf_context_create(int device, gf_context **out_ctx)
{
    if (!out_ctx) return GF_ERR_INVALID_ARGUMENT;
    *out_ctx = nullptr;
    return guarded([&]() -> int {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount", GF_ERR_CREATE_CONTEXT);
        if (device < 0 || device >= n) return GF_ERR_INVALID_DEVICE_ID;
        GF_CUDA(cudaSetDevice(device), GF_ERR_CREATE_CONTEXT);
        cudaDeviceProp p;
        GF_CUDA(cudaGetDeviceProperties(&p, device), GF_ERR_CREATE_CONTEXT);
        if (p.major < 10) {
            set_error_detail("gofeature_b200 kernels are built for sm_100a only");
            return GF_ERR_CREATE_CONTEXT;
        }
        std::unique_ptr<gf_context> c(new gf_context());
        c->device = device;
        c->sms = p.multiProcessorCount;
        GF_CUDA(cudaStreamCreateWithFlags(&c->mut_stream, cudaStreamNonBlocking), GF_ERR_CREATE_CONTEXT);
        c->stage_bytes = (size_t)8 << 20;
        for (int i = 0; i < 2; ++i) {
            GF_CUDA(cudaMallocHost((void **)&c->stage[i], c->stage_bytes), GF_ERR_CREATE_CONTEXT);
            GF_CUDA(cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming), GF_ERR_CREATE_CONTEXT);
        }
        *out_ctx = c.release();
        return GF_OK;
    });
}

int gf_context_destroy(gf_context *ctx)
{
    if (!ctx) return GF_ERR_INVALID_ARGUMENT;
    for (gf_context *c : ctx->children) {
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
    }
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    delete ctx;
    return GF_OK;
}

int gf_context_device(const gf_context *ctx) { return ctx ? ctx->device : -1; }

int gf_context_synchronize(gf_context *ctx)
{
    if (!ctx) return GF_ERR_INVALID_ARGUMENT;
    for (gf_context *c : ctx->children) {
        int rc = c->bind();
        if (rc) return rc;
        GF_CUDA(cudaDeviceSynchronize(), GF_ERR_CUDA);
    }
    int rc = ctx->bind();
    if (rc) return rc;
    GF_CUDA(cudaDeviceSynchronize(), GF_ERR_CUDA);
    return GF_OK;
}

int gf_context_stats(const gf_context *ctx, gf_stats *out)
{
    if (!ctx || !out) return GF_ERR_INVALID_ARGUMENT;
    if (ctx->is_multi()) {  // the sum over the parent and its devices
        memset(out, 0, sizeof *out);
        std::vector<const gf_context *> all(ctx->children.begin(), ctx->children.end());
        uint64_t *acc = reinterpret_cast<uint64_t *>(out);
        for (const gf_context *c : all) {
            gf_stats one;
            int rc = gf_context_stats(c, &one);
            if (rc) return rc;
            const uint64_t *v = reinterpret_cast<const uint64_t *>(&one);
            for (size_t i = 0; i < sizeof(gf_stats) / sizeof(uint64_t); ++i) acc[i] += v[i];
        }
        const Stats &p = ctx->stats;
        out->searches = p.searches;  // a search of the set, not one per shard
        out->coalesced_batches += p.coalesced_batches;
        out->quirk_fallbacks += p.quirk_fallbacks;
        out->uncertified_queries += p.uncertified_queries;
        out->tier2_queries += p.tier2_queries;
        out->h2d_bytes += p.h2d_bytes;
        out->d2h_bytes += p.d2h_bytes;
        out->host_stage_ns = p.host_stage_ns;  // (the parent's call, not the shards')
        out->host_launch_ns = p.host_launch_ns;
        out->host_wait_ns = p.host_wait_ns;
        out->host_resolve_ns = p.host_resolve_ns;
        out->host_calls = p.host_calls;
        return GF_OK;
    }
    const_cast<gf_context *>(ctx)->prof_collect();
    const Stats &s = ctx->stats;
    out->scan_kernel_ns = s.scan_kernel_ns;
    out->scan_kernel_timed = s.scan_kernel_timed;
    out->tensor_passes = s.tensor_passes;
    out->uncertified_queries = s.uncertified_queries;
    out->gemm_kernel_ns = s.gemm_kernel_ns;
    out->gemm_kernel_timed = s.gemm_kernel_timed;
    out->host_stage_ns = s.host_stage_ns;
    out->host_launch_ns = s.host_launch_ns;
    out->host_wait_ns = s.host_wait_ns;
    out->host_resolve_ns = s.host_resolve_ns;
    out->host_calls = s.host_calls;
    out->add_stage_ns = s.add_stage_ns;
    out->add_place_ns = s.add_place_ns;
    out->add_index_ns = s.add_index_ns;
    out->add_rows = s.add_rows;
    out->tier2_queries = s.tier2_queries;
    out->searches = s.searches;
    out->scan_launches = s.scan_launches;
    out->gemm_launches = s.gemm_launches;
    out->merge_launches = s.merge_launches;
    out->other_launches = s.other_launches;
    out->rows_scanned = s.rows_scanned;
    out->quirk_fallbacks = s.quirk_fallbacks;
    out->coalesced_batches = s.coalesced_batches;
    out->h2d_bytes = s.h2d_bytes;
    out->d2h_bytes = s.d2h_bytes;
    return GF_OK;
}

int gf_context_reset_stats(gf_context *ctx)
{
    if (!ctx) return GF_ERR_INVALID_ARGUMENT;
    for (gf_context *c : ctx->children) gf_context_reset_stats(c);
    Stats &s = ctx->stats;
    s.searches = s.scan_launches = s.gemm_launches = s.merge_launches = s.other_launches = 0;
    s.rows_scanned = s.quirk_fallbacks = s.coalesced_batches = s.h2d_bytes = s.d2h_bytes = 0;
    ctx->prof_collect();
    s.scan_kernel_ns = s.scan_kernel_timed = 0;
    s.tensor_passes = s.uncertified_queries = 0;
    s.gemm_kernel_ns = s.gemm_kernel_timed = 0;
    s.host_stage_ns = s.host_launch_ns = s.host_wait_ns = s.host_resolve_ns = s.host_calls = 0;
    s.add_stage_ns = s.add_place_ns = s.add_index_ns = s.add_rows = 0;
    s.tier2_queries = 0;
    return GF_OK;
}

int gf_context_set_profiling(gf_context *ctx, int on)
{
    if (!ctx) return GF_ERR_INVALID_ARGUMENT;
    for (gf_context *c : ctx->children) c->profiling = on ? 1 : 0;
    ctx->profiling = on ? 1 : 0;
    return GF_OK;
}

// ------------------------------------------------------------------ Buffer (buffer.go)
int gf_buffer_create(gf_context *ctx, uint64_t size, gf_buffer **out_buf)
{
    if (!ctx || !out_buf) return GF_ERR_INVALID_ARGUMENT;
    *out_buf = nullptr;
    return guarded([&]() -> int {
        int rc = ctx->bind();
        if (rc) return rc;
        void *p = nullptr;
        GF_CUDA(cudaMalloc(&p, size ? size : 1), GF_ERR_ALLOCATE_GPU_BUFFER);  // buffer.go:20-23
        cudaError_t e = cudaMemset(p, 0, size);                                 // buffer.go:24
        if (e != cudaSuccess) {
            cudaFree(p);
            return cuda_fail(e, "cudaMemset", GF_ERR_CLEAR_CUDA_BUFFER);
        }
        auto root = std::make_shared<DeviceAlloc>(ctx, p, size, true);
        gf_buffer *b = new gf_buffer();
        b->ctx = ctx;
        b->root = root;
        b->ptr = (uint8_t *)p;
        b->size = size;
        *out_buf = b;
        return GF_OK;
    });
}

int gf_buffer_wrap(gf_context *ctx, void *device_ptr, uint64_t size, gf_buffer **out_buf)
{
    if (!ctx || !out_buf || (!device_ptr && size)) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int {
        auto root = std::make_shared<DeviceAlloc>(ctx, device_ptr, size, false);
        gf_buffer *b = new gf_buffer();
        b->ctx = ctx;
        b->root = root;
        b->ptr = (uint8_t *)device_ptr;
        b->size = size;
        *out_buf = b;
        return GF_OK;
    });
}

int gf_buffer_write(gf_buffer *buf, const void *value, uint64_t len)
{
    if (!buf || (!value && len)) return GF_ERR_INVALID_ARGUMENT;
    if (len > buf->size) return GF_ERR_BUFFER_WRITE_OUT_OF_RANGE;  // buffer.go:35-37
    int rc = buf->ctx->bind();
    if (rc) return rc;
    return buf->ctx->upload(buf->ptr, value, len);
}

int gf_buffer_read(gf_buffer *buf, void *out, uint64_t cap)
{
    if (!buf || (!out && buf->size)) return GF_ERR_INVALID_ARGUMENT;
    if (cap < buf->size) return GF_ERR_INVALID_ARGUMENT;
    int rc = buf->ctx->bind();
    if (rc) return rc;
    GF_CUDA(cudaMemcpy(out, buf->ptr, buf->size, cudaMemcpyDeviceToHost), GF_ERR_READ_CUDA_BUFFER);
    buf->ctx->stats.d2h_bytes += buf->size;
    return GF_OK;
}

int gf_buffer_copy(gf_buffer *dst, const gf_buffer *src)
{
    if (!dst || !src) return GF_ERR_INVALID_ARGUMENT;
    if (src->size > dst->size) return GF_ERR_BUFFER_COPY_OUT_OF_RANGE;  // buffer.go:55-57
    int rc = dst->ctx->bind();
    if (rc) return rc;
    GF_CUDA(cudaMemcpy(dst->ptr, src->ptr, src->size, cudaMemcpyDeviceToDevice), GF_ERR_WRITE_CUDA_BUFFER);
    return GF_OK;
}

int gf_buffer_reset(gf_buffer *buf)
{
    if (!buf) return GF_ERR_INVALID_ARGUMENT;
    int rc = buf->ctx->bind();
    if (rc) return rc;
    GF_CUDA(cudaMemset(buf->ptr, 0, buf->size), GF_ERR_CLEAR_CUDA_BUFFER);
    return GF_OK;
}

int gf_buffer_slice(gf_buffer *buf, int64_t start, int64_t end, gf_buffer **out_buf)
{
    if (!buf || !out_buf) return GF_ERR_INVALID_ARGUMENT;
    *out_buf = nullptr;
    // buffer.go:74-79
    if (start < 0 || (uint64_t)start > buf->size) return GF_ERR_BUFFER_SLICE_OUT_OF_RANGE;
    if (end < 0 || (uint64_t)end > buf->size || end < start) return GF_ERR_BUFFER_SLICE_OUT_OF_RANGE;
    return guarded([&]() -> int {
        gf_buffer *b = new gf_buffer();
        b->ctx = buf->ctx;
        b->root = buf->root;
        b->ptr = buf->ptr + start;
        b->size = (uint64_t)(end - start);
        *out_buf = b;
        return GF_OK;
    });
}

uint64_t gf_buffer_size(const gf_buffer *buf) { return buf ? buf->size : 0; }
void *gf_buffer_device_ptr(const gf_buffer *buf) { return buf ? buf->ptr : nullptr; }
int gf_buffer_free(gf_buffer *buf)
{
    delete buf;
    return GF_OK;
}

// ------------------------------------------------------------------ Cache (cache.go)
int gf_cache_create(gf_context *ctx, int block_num, int64_t block_size, gf_cache **out_cache)
{
    const char *env = getenv("GF_SHADOW");  // "0" = no shadow slab, "bf16" / "int8" = that format
    unsigned flags = 0;
    if (env && env[0] == '0') flags = GF_CACHE_NO_SHADOW;
    if (env && env[0] == 'b') flags = GF_CACHE_SHADOW_BF16;
    return gf_cache_create_ex(ctx, block_num, block_size, flags, out_cache);
}

int gf_cache_has_shadow(const gf_cache *cache)
{
    if (cache && cache->is_multi()) return gf_cache_has_shadow(cache->children[0]);
    return (cache && cache->shadow) ? cache->shadow_eb : 0;
}

int gf_cache_create_ex(gf_context *ctx, int block_num, int64_t block_size, unsigned flags, gf_cache **out_cache)
{
    if (!ctx || !out_cache || block_num < 0 || block_size <= 0) return GF_ERR_INVALID_ARGUMENT;
    *out_cache = nullptr;
    if (ctx->is_multi()) {
        // block i of the cache = block i / G of device i % G's slab (cache.go:33-41 cuts ONE slab; here one per device)
        return guarded([&]() -> int {
            const int G = (int)ctx->children.size();
            std::unique_ptr<gf_cache> c(new gf_cache());
            c->ctx = ctx;
            c->block_size = block_size;
            c->multi_blocks = block_num;
            bool shadow_everywhere = true;
            for (int d = 0; d < G; ++d) {
                gf_cache *cc = nullptr;
                const int nb = (block_num - d + G - 1) / G;
                int rc = gf_cache_create_ex(ctx->children[d], nb < 0 ? 0 : nb, block_size, flags, &cc);
                if (rc) {
                    for (gf_cache *x : c->children) gf_cache_destroy(x);
                    return rc;
                }
                c->children.push_back(cc);
                shadow_everywhere = shadow_everywhere && (cc->shadow != nullptr || nb <= 0);
            }
            (void)shadow_everywhere;
            ctx->bind();
            *out_cache = c.release();
            return GF_OK;
        });
    }
    return guarded([&]() -> int {
        int rc = ctx->bind();
        if (rc) return rc;
        std::unique_ptr<gf_cache> c(new gf_cache());
        c->ctx = ctx;
        c->block_size = block_size;
        uint64_t total = (uint64_t)block_num * (uint64_t)block_size;
        void *p = nullptr;
        // one slab (cache.go:31) + a tail pad so a 16-byte bulk copy never runs off the allocation
        GF_CUDA(cudaMalloc(&p, total + 256), GF_ERR_ALLOCATE_GPU_BUFFER);
        cudaError_t e = cudaMemset(p, 0, total + 256);
        if (e != cudaSuccess) {
            cudaFree(p);
            return cuda_fail(e, "cudaMemset", GF_ERR_CLEAR_CUDA_BUFFER);
        }
        c->slab = std::make_shared<DeviceAlloc>(ctx, p, total, true);
        // bf16 shadow of the slab for the tensor path (half the bytes per scanned row).  It needs block_size / 2
        // bytes per block; when the device cannot hold it the cache works without (GF_CACHE_REQUIRE_SHADOW: fail).
        const int seb = (flags & GF_CACHE_SHADOW_BF16) ? 2 : 1;  // default: int8 + one float scale per row
        const bool want_shadow = !(flags & GF_CACHE_NO_SHADOW) && block_size % 512 == 0 && total > 0;
        if (want_shadow) {
            void *sp = nullptr, *scp = nullptr;
            const uint64_t sbytes = total / 4 * seb, scbytes = seb == 1 ? total / 128 : 0;
            cudaError_t se = cudaMalloc(&sp, sbytes + 256);
            if (se == cudaSuccess) se = cudaMemset(sp, 0, sbytes + 256);
            if (se == cudaSuccess && scbytes) se = cudaMalloc(&scp, scbytes + 256);
            if (se == cudaSuccess && scbytes) se = cudaMemset(scp, 0, scbytes + 256);
            if (se != cudaSuccess) {
                if (sp) cudaFree(sp);
                if (scp) cudaFree(scp);
                cudaGetLastError();
                if (flags & GF_CACHE_REQUIRE_SHADOW) return cuda_fail(se, "cudaMalloc(shadow)", GF_ERR_ALLOCATE_GPU_BUFFER);
            } else {
                c->shadow = std::make_shared<DeviceAlloc>(ctx, sp, sbytes, true);
                if (scp) c->scales = std::make_shared<DeviceAlloc>(ctx, scp, scbytes, true);
                c->shadow_eb = seb;
            }
        } else if (flags & GF_CACHE_REQUIRE_SHADOW) {
            return GF_ERR_INVALID_ARGUMENT;
        }
        for (int i = 0; i < block_num; ++i) {  // cache.go:33-41
            std::unique_ptr<Block> b(new gf_block());
            b->cache = c.get();
            b->index = i;
            b->block_size = block_size;
            b->dev = (uint8_t *)p + (uint64_t)i * (uint64_t)block_size;
            if (c->shadow) {
                b->shadow = (uint8_t *)c->shadow->ptr + (uint64_t)i * (uint64_t)(block_size / 4 * c->shadow_eb);
                b->shadow_eb = c->shadow_eb;
            }
            c->all_blocks.push_back(std::move(b));
        }
        *out_cache = c.release();
        return GF_OK;
    });
}

int gf_cache_destroy(gf_cache *cache)
{
    if (!cache) return GF_ERR_INVALID_ARGUMENT;
    if (cache->is_multi()) {
        for (gf_context *c : cache->ctx->children) {  // no search of a parent set may still be running
            cudaSetDevice(c->device);
            cudaDeviceSynchronize();
        }
        for (gf_cache *cc : cache->children) gf_cache_destroy(cc);
        cache->children.clear();
        delete cache;
        return GF_OK;
    }
    cudaSetDevice(cache->ctx->device);
    cudaDeviceSynchronize();
    for (auto &s : cache->owned_sets) {
        if (s->d_segs) cudaFree(s->d_segs);
        s->d_segs = nullptr;
        if (s->d_small) cudaFree(s->d_small);
        s->d_small = nullptr;
        if (s->dev_ws) cache->ctx->release_ws(s->dev_ws);
        s->dev_ws = nullptr;
    }
    delete cache;
    return GF_OK;
}

int gf_cache_new_set(gf_cache *cache, const char *name, int dims, int precision, int batch)
{
    if (!cache || !name || dims <= 0 || precision <= 0 || batch <= 0) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int {
        std::lock_guard<std::mutex> g(cache->mu);
        if (cache->sets.count(name)) return GF_ERR_FEATURE_SET_EXIST;  // cache.go:47-50
        std::vector<gf_set *> shards;
        if (cache->is_multi()) {
            const int G = (int)cache->children.size();
            for (int d = 0; d < G; ++d) {
                gf_set *sh = nullptr;
                int rc = gf_cache_new_set(cache->children[d], name, dims, precision, batch);
                if (rc == GF_OK) rc = gf_cache_get_set(cache->children[d], name, &sh);
                if (rc == GF_OK) rc = gf_set_set_shard(sh, d, G);
                if (rc) {
                    for (int u = 0; u <= d; ++u) gf_cache_destroy_set(cache->children[u], name);
                    return rc;
                }
                shards.push_back(sh);
            }
        }
        std::unique_ptr<Set> s(new gf_set());
        s->shards = shards;
        s->cache = cache;
        s->ctx = cache->ctx;
        s->name = name;
        s->dims = dims;
        s->precision = precision;
        s->batch = batch;
        s->cap = (int)(cache->block_size / ((int64_t)dims * precision));  // cache.go:59
        cache->sets[name] = s.get();
        cache->owned_sets.push_back(std::move(s));
        return GF_OK;
    });
}

int gf_cache_destroy_set(gf_cache *cache, const char *name)
{
    if (!cache || !name) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int {
        Set *s;
        {
            std::lock_guard<std::mutex> g(cache->mu);
            auto it = cache->sets.find(name);
            if (it == cache->sets.end()) return GF_ERR_FEATURE_SET_NOT_FOUND;  // cache.go:76-80
            s = it->second;
        }
        int rc = s->destroy();
        if (rc) return rc;
        std::lock_guard<std::mutex> g(cache->mu);
        cache->sets.erase(name);
        return GF_OK;
    });
}

int gf_cache_get_set(gf_cache *cache, const char *name, gf_set **out_set)
{
    if (!cache || !name || !out_set) return GF_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> g(cache->mu);
    auto it = cache->sets.find(name);
    if (it == cache->sets.end()) {
        *out_set = nullptr;
        return GF_ERR_FEATURE_SET_NOT_FOUND;
    }
    *out_set = it->second;
    return GF_OK;
}

int64_t gf_cache_get_block_size(const gf_cache *cache) { return cache ? cache->block_size : 0; }

int gf_cache_get_empty_block(gf_cache *cache, int block_num, gf_block **out_blocks)
{
    if (!cache || block_num < 0 || (block_num > 0 && !out_blocks)) return GF_ERR_INVALID_ARGUMENT;
    if (cache->is_multi()) {  // AllBlocks order over the striped view
        std::lock_guard<std::mutex> g(cache->mu);
        int k = 0;
        for (int i = 0; i < cache->multi_blocks && k < block_num; ++i) {
            Block *b = cache->multi_block_at(i);
            if (!b->is_owned()) out_blocks[k++] = b;
        }
        return k < block_num ? GF_ERR_NOT_ENOUGH_BLOCKS : GF_OK;
    }
    std::lock_guard<std::mutex> g(cache->mu);
    int n = 0;
    for (auto &b : cache->all_blocks)
        if (!b->is_owned()) ++n;
    if (n < block_num) return GF_ERR_NOT_ENOUGH_BLOCKS;
    int k = 0;
    for (auto &b : cache->all_blocks) {
        if (k == block_num) break;
        if (!b->is_owned()) out_blocks[k++] = b.get();
    }
    return GF_OK;
}

int gf_cache_block_count(const gf_cache *cache)
{
    return cache ? (cache->is_multi() ? cache->multi_blocks : (int)cache->all_blocks.size()) : 0;
}

int gf_cache_block_at(gf_cache *cache, int index, gf_block **out_block)
{
    if (cache && cache->is_multi() && out_block && index >= 0 && index < cache->multi_blocks) {
        *out_block = cache->multi_block_at(index);
        return GF_OK;
    }
    if (!cache || !out_block || index < 0 || index >= (int)cache->all_blocks.size()) return GF_ERR_INVALID_ARGUMENT;
    *out_block = cache->all_blocks[index].get();
    return GF_OK;
}

// ------------------------------------------------------------------ Block (block.go)
int gf_block_capacity(const gf_block *block) { return block ? block->capacity() : 0; }
int gf_block_margin(const gf_block *block) { return (block && block->is_owned()) ? block->margin() : 0; }
int gf_block_is_owned(const gf_block *block) { return block ? (block->is_owned() ? 1 : 0) : 0; }
int gf_block_next_index(const gf_block *block) { return block ? block->next_index : 0; }

int gf_block_acquire(gf_block *block, const char *owner, int dims, int precision, int batch)
{
    if (!block || !owner) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int {
        std::lock_guard<std::mutex> g(block->cache->mu);
        return block->acquire(owner, dims, precision, batch);
    });
}

int gf_block_release(gf_block *block)
{
    if (!block) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int { return block->release(); });
}

int gf_block_insert(gf_block *block, int64_t n, const void *values, uint64_t values_len, const char *ids,
                    const uint64_t *id_offsets)
{
    if (!block || n < 0 || (n > 0 && !values)) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int {
        if (!block->is_owned()) return GF_ERR_INVALID_ARGUMENT;
        if (values_len != (uint64_t)n * block->dims * block->precision) return GF_ERR_MISMATCH_DIMENSION;
        std::vector<std::string> v;
        int rc = unpack_ids(n, ids, id_offsets, v);
        if (rc) return rc;
        if (block->set) return GF_ERR_BLOCK_USED;  // blocks of a live set are filled through Set.Add
        return block->insert(n, (const uint8_t *)values, v);
    });
}

int gf_block_delete(gf_block *block, int64_t n, const char *ids, const uint64_t *id_offsets, uint8_t *out_deleted,
                    int64_t *out_ndeleted)
{
    if (!block || n < 0) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int {
        std::vector<std::string> v;
        int rc = unpack_ids(n, ids, id_offsets, v);
        if (rc) return rc;
        if (block->set) return GF_ERR_BLOCK_USED;
        std::vector<uint8_t> found((size_t)n, 0);
        rc = block->remove(v, found);
        if (rc) return rc;
        int64_t c = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (out_deleted) out_deleted[i] = found[i];
            c += found[i];
        }
        if (out_ndeleted) *out_ndeleted = c;
        return GF_OK;
    });
}

int gf_block_search(gf_block *block, const gf_buffer *input, gf_buffer *output, int batch, int limit,
                    gf_result **out_result)
{
    (void)output;  // the reference's score matrix scratch (block.go:71); scores never leave the SM here
    if (!block || !input || !out_result) return GF_ERR_INVALID_ARGUMENT;
    *out_result = nullptr;
    return guarded([&]() -> int {
        Result *r = nullptr;
        int rc = block_search(block, input, batch, limit, &r);
        if (rc) return rc;
        *out_result = r;
        return GF_OK;
    });
}

// ------------------------------------------------------------------ Set (set.go)
int gf_set_add(gf_set *set, int64_t n, const void *values, uint64_t values_len, const char *ids,
               const uint64_t *id_offsets)
{
    if (!set || n < 0 || (n > 0 && !values)) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int {
        if (values_len != (uint64_t)n * set->dims * set->precision) return GF_ERR_MISMATCH_DIMENSION;
        std::vector<std::string> v;
        int rc = unpack_ids(n, ids, id_offsets, v);
        if (rc) return rc;
        return set->add(n, (const uint8_t *)values, v);
    });
}

int gf_set_delete(gf_set *set, int64_t n, const char *ids, const uint64_t *id_offsets, uint8_t *out_deleted,
                  int64_t *out_ndeleted)
{
    if (!set || n < 0) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int {
        std::vector<std::string> v;
        int rc = unpack_ids(n, ids, id_offsets, v);
        if (rc) return rc;
        std::vector<uint8_t> found;
        rc = set->remove(v, found);
        if (rc) return rc;
        int64_t c = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (out_deleted) out_deleted[i] = found[i];
            c += found[i];
        }
        if (out_ndeleted) *out_ndeleted = c;
        return GF_OK;
    });
}

int gf_set_update(gf_set *set, int64_t n, const void *values, uint64_t values_len, const char *ids,
                  const uint64_t *id_offsets, uint8_t *out_updated, int64_t *out_nupdated)
{
    if (!set || n < 0 || (n > 0 && !values)) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int {
        if (values_len != (uint64_t)n * set->dims * set->precision) return GF_ERR_MISMATCH_DIMENSION;
        std::vector<std::string> v;
        int rc = unpack_ids(n, ids, id_offsets, v);
        if (rc) return rc;
        std::vector<uint8_t> found;
        rc = set->update(n, (const uint8_t *)values, v, found);
        if (rc) return rc;
        int64_t c = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (out_updated) out_updated[i] = found[i];
            c += found[i];
        }
        if (out_nupdated) *out_nupdated = c;
        return GF_OK;
    });
}

int gf_set_read(gf_set *set, int64_t n, const char *ids, const uint64_t *id_offsets, void *out_values,
                uint64_t out_values_cap, uint8_t *out_found, int64_t *out_nfound)
{
    if (!set || n < 0 || (n > 0 && !out_values)) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int {
        if (out_values_cap < (uint64_t)n * set->dims * set->precision) return GF_ERR_INVALID_ARGUMENT;
        std::vector<std::string> v;
        int rc = unpack_ids(n, ids, id_offsets, v);
        if (rc) return rc;
        std::vector<uint8_t> found;
        rc = set->read(v, (uint8_t *)out_values, found);
        if (rc) return rc;
        int64_t c = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (out_found) out_found[i] = found[i];
            c += found[i];
        }
        if (out_nfound) *out_nfound = c;
        return GF_OK;
    });
}

int gf_set_search(gf_set *set, float threshold, int limit, int q, const void *queries, uint64_t queries_len,
                  gf_result **out_result)
{
    if (!set || !out_result || q < 0 || (q > 0 && !queries)) return GF_ERR_INVALID_ARGUMENT;
    *out_result = nullptr;
    return guarded([&]() -> int {
        if (q > set->batch) return GF_ERR_OUT_OF_BATCH;  // set.go:119-122 comes first
        if (queries_len != (uint64_t)q * set->dims * set->precision) return GF_ERR_MISMATCH_DIMENSION;
        Result *r = nullptr;
        int rc = set->search(threshold, limit, q, (const uint8_t *)queries, &r);
        if (rc) return rc;
        *out_result = r;
        return GF_OK;
    });
}

int gf_set_destroy(gf_set *set)
{
    if (!set) return GF_ERR_INVALID_ARGUMENT;
    return gf_cache_destroy_set(set->cache, set->name.c_str());
}

int gf_set_compact(gf_set *set, int64_t *out_reclaimed_rows)
{
    if (!set) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int { return set->compact(out_reclaimed_rows); });
}

int gf_set_export(gf_set *set, void *out_values, uint64_t values_cap, char *out_ids, uint64_t ids_cap,
                  uint64_t *out_id_offsets, int64_t *out_rows, uint64_t *out_id_bytes)
{
    if (!set) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int {
        return set->export_rows((uint8_t *)out_values, values_cap, out_ids, ids_cap, out_id_offsets, out_rows,
                                out_id_bytes);
    });
}

int gf_set_dims(const gf_set *set) { return set ? set->dims : 0; }
int gf_set_precision(const gf_set *set) { return set ? set->precision : 0; }
int gf_set_batch(const gf_set *set) { return set ? set->batch : 0; }
int gf_set_block_count(const gf_set *set)
{
    if (!set) return 0;
    int n = (int)set->blocks.size();
    for (const gf_set *sh : set->shards) n += (int)sh->blocks.size();
    return n;
}
int64_t gf_set_scanned_rows(const gf_set *set) { return set ? set->scanned_rows : 0; }
int64_t gf_set_live_rows(const gf_set *set)
{
    if (!set) return 0;
    int64_t n = 0;
    for (const gf_set *sh : set->shards) n += gf_set_live_rows(sh);
    for (const Block *b : set->blocks) n += b->live_count;
    return n;
}

// ------------------------------------------------------------------ results
int gf_result_batch(const gf_result *r) { return r ? r->batch : 0; }
int gf_result_count(const gf_result *r, int q)
{
    if (!r || q < 0 || q >= r->batch) return 0;
    return (int)(r->offsets[q + 1] - r->offsets[q]);
}
float gf_result_score(const gf_result *r, int q, int i)
{
    if (!r || q < 0 || q >= r->batch || i < 0 || i >= gf_result_count(r, q)) return 0.0f;
    return r->scores[(size_t)r->offsets[q] + i];
}
const char *gf_result_id(const gf_result *r, int q, int i, uint32_t *out_len)
{
    if (out_len) *out_len = 0;
    if (!r || q < 0 || q >= r->batch || i < 0 || i >= gf_result_count(r, q)) return nullptr;
    size_t k = (size_t)r->offsets[q] + i;
    if (out_len) *out_len = (uint32_t)(r->id_off[k + 1] - r->id_off[k]);
    return r->id_bytes.data() + r->id_off[k];
}
uint64_t gf_result_slot(const gf_result *r, int q, int i)
{
    if (!r || q < 0 || q >= r->batch || i < 0 || i >= gf_result_count(r, q)) return 0;
    return r->slots[(size_t)r->offsets[q] + i];
}
int64_t gf_result_total(const gf_result *r) { return r ? (int64_t)r->scores.size() : 0; }
int64_t gf_result_id_bytes(const gf_result *r) { return r ? (int64_t)r->id_bytes.size() : 0; }
int gf_result_export(const gf_result *r, int32_t *out_counts, float *out_scores, uint64_t *out_slots, char *out_ids,
                     uint64_t *out_id_offsets)
{
    if (!r) return GF_ERR_INVALID_ARGUMENT;
    if (out_counts)
        for (int q = 0; q < r->batch; ++q) out_counts[q] = (int32_t)(r->offsets[q + 1] - r->offsets[q]);
    size_t t = r->scores.size();
    if (out_scores && t) memcpy(out_scores, r->scores.data(), t * sizeof(float));
    if (out_slots && t) memcpy(out_slots, r->slots.data(), t * sizeof(uint64_t));
    if (out_ids && !r->id_bytes.empty()) memcpy(out_ids, r->id_bytes.data(), r->id_bytes.size());
    if (out_id_offsets) memcpy(out_id_offsets, r->id_off.data(), (t + 1) * sizeof(uint64_t));
    return GF_OK;
}
int gf_result_free(gf_result *r)
{
    delete r;
    return GF_OK;
}

// ------------------------------------------------------------------ device-resident entry points
uint64_t gf_key_make(float score, uint32_t slot) { return make_key(score, slot); }
float gf_key_score(uint64_t key) { return key_score(key); }
uint32_t gf_key_slot(uint64_t key) { return key_slot(key); }

int gf_set_set_shard(gf_set *set, int rank, int world)
{
    if (!set || world < 1 || rank < 0 || rank >= world) return GF_ERR_INVALID_ARGUMENT;
    if (set->is_multi()) return GF_ERR_UNSUPPORTED;  // its shards already carry their identities
    return guarded([&]() -> int {
        set->lock.lock();
        set->shard_rank = rank;
        set->shard_world = world;
        int rc = set->rebuild_segments();
        set->lock.unlock();
        return rc;
    });
}

int gf_set_search_device(gf_set *set, int limit, int q, const float *d_queries, uint64_t *d_out_keys, int path,
                         void *stream)
{
    if (!set || !d_queries || !d_out_keys) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int {
        return set->search_device(limit, q, d_queries, (unsigned long long *)d_out_keys, path, (cudaStream_t)stream);
    });
}

int gf_set_search_device_ex(gf_set *set, int limit, int q, const float *d_queries, uint64_t *d_out_keys,
                            uint32_t *d_out_flags, int path, void *stream)
{
    if (!set || !d_queries || !d_out_keys) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int {
        return set->search_device(limit, q, d_queries, (unsigned long long *)d_out_keys, path, (cudaStream_t)stream,
                                  (unsigned int *)d_out_flags);
    });
}

int gf_set_device_counters(gf_set *set, uint64_t *out_flagged, uint64_t *out_unrepaired)
{
    if (!set) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int {
        unsigned int h[4] = {0, 0, 0, 0};
        if (set->is_multi()) {
            uint64_t a = 0, b = 0;
            for (gf_set *sh : set->shards) {
                uint64_t x = 0, y = 0;
                int rc = gf_set_device_counters(sh, &x, &y);
                if (rc) return rc;
                a += x;
                b += y;
            }
            if (out_flagged) *out_flagged = a;
            if (out_unrepaired) *out_unrepaired = b;
            return GF_OK;
        }
        set->lock.lock_shared();
        cudaError_t e = cudaSuccess;
        if (set->d_small != nullptr && set->ctx->bind() == GF_OK) {
            e = cudaDeviceSynchronize();
            if (e == cudaSuccess) e = cudaMemcpy(h, set->d_small, sizeof h, cudaMemcpyDeviceToHost);
        }
        set->lock.unlock_shared();
        if (e != cudaSuccess) return cuda_fail(e, "gf_set_device_counters", GF_ERR_READ_CUDA_BUFFER);
        if (out_flagged) *out_flagged = h[3];
        if (out_unrepaired) *out_unrepaired = h[1];
        return GF_OK;
    });
}

// The low word of a sequence number tags every word of the step (gf_exchange.cuh) and must never be 0 (a fresh
// inbox is zeroed); skipping TWO numbers keeps the slot parity alternating.
static unsigned long long next_exchange_seq(unsigned long long s)
{
    unsigned long long n = s + 1;
    if ((n & 0xFFFFFFFFull) == 0ull) n += 2;
    return n;
}

// ---- peer exchange of per-rank winners over NVLink peer memory (one process per GPU)
// (gf_set_search_device_xchg, below gf_exchange_merge_device, runs a search that ends with the exchange)
struct gf_exchange {
    gf_context *ctx = nullptr;
    int rank = 0, world = 1, max_q = 0, max_k = 0;
    void *inbox = nullptr;
    std::vector<void *> peers;  // mapped inboxes; peers[rank] == inbox
    void **d_peers = nullptr;
    unsigned long long seq = 0;
    bool connected = false;
    unsigned long long *d_trace = nullptr;  // [max_q][4] timestamps of the last step (gf_exchange_trace)
};

int gf_exchange_create(gf_context *ctx, int rank, int world, int max_q, int max_limit, gf_exchange **out)
{
    if (!ctx || !out || world < 1 || world > 64 || rank < 0 || rank >= world || max_q < 1 || max_limit < 1)
        return GF_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if ((int64_t)world * max_limit > kExchangeMaxKeys) return GF_ERR_UNSUPPORTED;
    return guarded([&]() -> int {
        int rc = ctx->bind();
        if (rc) return rc;
        std::unique_ptr<gf_exchange> x(new gf_exchange());
        x->ctx = ctx;
        x->rank = rank;
        x->world = world;
        x->max_q = max_q;
        x->max_k = max_limit;
        const size_t bytes = exchange_inbox_bytes(world, max_q, max_limit);
        GF_CUDA(cudaMalloc(&x->inbox, bytes), GF_ERR_ALLOCATE_GPU_BUFFER);  // cudaMalloc: exportable through CUDA IPC
        GF_CUDA(cudaMemset(x->inbox, 0, bytes), GF_ERR_CLEAR_CUDA_BUFFER);
        GF_CUDA(cudaMalloc((void **)&x->d_peers, sizeof(void *) * world), GF_ERR_ALLOCATE_GPU_BUFFER);
        GF_CUDA(cudaMalloc((void **)&x->d_trace, sizeof(unsigned long long) * 4 * max_q), GF_ERR_ALLOCATE_GPU_BUFFER);
        GF_CUDA(cudaMemset(x->d_trace, 0, sizeof(unsigned long long) * 4 * max_q), GF_ERR_CLEAR_CUDA_BUFFER);
        GF_CUDA(cudaDeviceSynchronize(), GF_ERR_CUDA);
        x->peers.assign((size_t)world, nullptr);
        x->peers[rank] = x->inbox;
        *out = x.release();
        return GF_OK;
    });
}

int gf_exchange_handle(gf_exchange *x, void *out_handle, uint64_t cap)
{
    if (!x || !out_handle || cap < sizeof(cudaIpcMemHandle_t)) return GF_ERR_INVALID_ARGUMENT;
    int rc = x->ctx->bind();
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    GF_CUDA(cudaIpcGetMemHandle(&h, x->inbox), GF_ERR_CUDA);
    memcpy(out_handle, &h, sizeof h);
    return GF_OK;
}

int gf_exchange_connect(gf_exchange *x, const void *handles, uint64_t len)
{
    if (!x || !handles || len < (uint64_t)x->world * sizeof(cudaIpcMemHandle_t)) return GF_ERR_INVALID_ARGUMENT;
    if (x->connected) return GF_ERR_INVALID_ARGUMENT;
    int rc = x->ctx->bind();
    if (rc) return rc;
    for (int p = 0; p < x->world; ++p) {
        if (p == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t *)handles + (size_t)p * sizeof h, sizeof h);
        GF_CUDA(cudaIpcOpenMemHandle(&x->peers[p], h, cudaIpcMemLazyEnablePeerAccess), GF_ERR_CUDA);
    }
    GF_CUDA(cudaMemcpy(x->d_peers, x->peers.data(), sizeof(void *) * x->world, cudaMemcpyHostToDevice), GF_ERR_CUDA);
    x->connected = true;
    return GF_OK;
}

int gf_exchange_merge_device(gf_exchange *x, int q, int limit, const uint64_t *d_local_keys,
                             const uint32_t *d_local_flags, uint64_t *d_out_keys, uint32_t *d_out_flags, void *stream)
{
    if (!x || !x->connected || !d_local_keys || !d_out_keys) return GF_ERR_INVALID_ARGUMENT;
    if (q < 1 || limit < 1 || q > x->max_q || limit > x->max_k) return GF_ERR_UNSUPPORTED;
    int rc = x->ctx->bind();
    if (rc) return rc;
    // every rank calls this the same number of times in the same order: the sequence number is implicit
    const unsigned long long seq = x->seq = next_exchange_seq(x->seq);
    GF_CUDA(exchange_merge_launch(x->d_peers, x->rank, x->world, q, limit, x->max_q, x->max_k, seq,
                                  (const unsigned long long *)d_local_keys, d_local_flags,
                                  (unsigned long long *)d_out_keys, d_out_flags, (cudaStream_t)stream),
            GF_ERR_CUDA);
    x->ctx->stats.merge_launches++;
    return GF_OK;
}

int gf_set_search_device_xchg(gf_set *set, gf_exchange *x, int limit, int q, const float *d_queries,
                              uint64_t *d_local_keys, uint32_t *d_local_flags, uint64_t *d_out_keys,
                              uint32_t *d_out_flags, int path, void *stream)
{
    if (!set || !x || !x->connected || !d_queries || !d_local_keys || !d_local_flags || !d_out_keys || !d_out_flags)
        return GF_ERR_INVALID_ARGUMENT;
    if (q < 1 || limit < 1 || q > x->max_q || limit > x->max_k) return GF_ERR_UNSUPPORTED;
    return guarded([&]() -> int {
        // every rank calls this the same number of times in the same order: the sequence number is implicit
        const unsigned long long prev_seq = x->seq;
        const unsigned long long seq = x->seq = next_exchange_seq(prev_seq);
        ExchangeArgs a{x->d_peers, x->rank, x->world, q, limit, x->max_q, x->max_k, seq,
                       (const unsigned long long *)d_local_keys, d_local_flags, (unsigned long long *)d_out_keys,
                       d_out_flags, 5000000000ull};
        a.trace = x->ctx->profiling.load() ? x->d_trace : nullptr;
        bool fused = false;
        int rc = set->search_device(limit, q, d_queries, (unsigned long long *)d_local_keys, path, (cudaStream_t)stream,
                                    d_local_flags, &a, &fused);
        if (rc) {
            x->seq = prev_seq;  // nothing was pushed: the peers' next step still carries this number
            return rc;
        }
        if (!fused) {
            GF_CUDA(exchange_merge_launch(x->d_peers, x->rank, x->world, q, limit, x->max_q, x->max_k, seq,
                                          (const unsigned long long *)d_local_keys, d_local_flags,
                                          (unsigned long long *)d_out_keys, d_out_flags, (cudaStream_t)stream),
                    GF_ERR_CUDA);
            x->ctx->stats.merge_launches++;
        }
        return GF_OK;
    });
}

int gf_exchange_trace(gf_exchange *x, int q, uint64_t *out_ns)
{
    if (!x || !out_ns || q < 1 || q > x->max_q) return GF_ERR_INVALID_ARGUMENT;
    int rc = x->ctx->bind();
    if (rc) return rc;
    GF_CUDA(cudaDeviceSynchronize(), GF_ERR_CUDA);
    GF_CUDA(cudaMemcpy(out_ns, x->d_trace, sizeof(uint64_t) * 4 * q, cudaMemcpyDeviceToHost), GF_ERR_READ_CUDA_BUFFER);
    return GF_OK;
}

int gf_exchange_destroy(gf_exchange *x)
{
    if (!x) return GF_OK;
    x->ctx->bind();
    cudaDeviceSynchronize();
    for (int p = 0; p < x->world; ++p)
        if (p != x->rank && x->peers[p]) cudaIpcCloseMemHandle(x->peers[p]);
    if (x->d_peers) cudaFree(x->d_peers);
    if (x->d_trace) cudaFree(x->d_trace);
    if (x->inbox) cudaFree(x->inbox);
    delete x;
    return GF_OK;
}

int gf_merge_keys_device(gf_context *ctx, int lists, int q, int limit, const uint64_t *d_in, uint64_t *d_out,
                         void *stream)
{
    if (!ctx || !d_in || !d_out || lists < 1 || q < 1 || limit < 1) return GF_ERR_INVALID_ARGUMENT;
    if (limit > kMaxK) return GF_ERR_UNSUPPORTED;
    int rc = ctx->bind();
    if (rc) return rc;
    GF_CUDA(merge_launch((const unsigned long long *)d_in, 1, lists, q, limit, (unsigned long long *)d_out,
                         (cudaStream_t)stream),
            GF_ERR_CUDA);
    ctx->stats.merge_launches++;
    return GF_OK;
}

int gf_set_resolve_keys(gf_set *set, float threshold, int limit, int q, const uint64_t *h_keys, gf_result **out_result)
{
    if (!set || !h_keys || !out_result || q < 0 || limit < 0) return GF_ERR_INVALID_ARGUMENT;
    *out_result = nullptr;
    return guarded([&]() -> int {
        Result *r = nullptr;
        int rc = set->resolve_keys(threshold, limit, q, (const unsigned long long *)h_keys, &r);
        if (rc) return rc;
        *out_result = r;
        return GF_OK;
    });
}

int gf_set_add_synthetic(gf_set *set, int64_t n, uint64_t seed, int64_t first_ordinal, int normalize,
                         const char *id_prefix)
{
    if (!set || n < 0 || first_ordinal < 0) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int {
        return set->add_synthetic(n, seed, first_ordinal, normalize, id_prefix ? id_prefix : "f");
    });
}

int gf_set_add_synthetic_ex(gf_set *set, int64_t n, uint64_t seed, int64_t first_ordinal, int normalize,
                            const char *id_prefix, uint64_t centres, float spread, uint32_t dup_ppm)
{
    if (!set || n < 0 || first_ordinal < 0 || dup_ppm > 1000000u || !(spread >= 0.f)) return GF_ERR_INVALID_ARGUMENT;
    return guarded([&]() -> int {
        SynthSpec sp(seed);
        sp.centres = centres;
        sp.spread = spread;
        sp.dup_ppm = dup_ppm;
        return set->add_synthetic(n, sp, first_ordinal, normalize, id_prefix ? id_prefix : "f");
    });
}

int gf_set_set_search_path(gf_set *set, int path)
{
    if (!set || path < 0 || path > 5) return GF_ERR_INVALID_ARGUMENT;
    for (gf_set *sh : set->shards) sh->search_path = path;
    set->search_path = path;
    return GF_OK;
}

int gf_set_set_coalescing(gf_set *set, int max_wait_us)
{
    if (!set || max_wait_us < 0) return GF_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> g(set->co_mu);
    set->coalesce_wait_us = max_wait_us;
    return GF_OK;
}

}  // extern "C"
