// gf_runtime.cu -- host object model: Context / Buffer / Block / Set / Cache.
//
// Restates the reference's Go host code (block.go, buffer.go, cache.go, set.go, util.go)
// in C++ behind the C ABI; citations are /root/reference/<file>:<line>.  The GPU work is
// in gf_scan.cu / gf_merge.cu.  There is no CPU scoring path in this file or anywhere in
// the library: every score comes out of a CUDA kernel.
#include "gf_runtime.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>

#include "gf_device.cuh"

namespace gf {

// ------------------------------------------------------------------ errors
static thread_local std::string g_detail;
void set_error_detail(const std::string &s) { g_detail = s; }
const std::string &error_detail() { return g_detail; }

int cuda_fail(cudaError_t e, const char *what, int status)
{
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    g_detail = buf;
    cudaGetLastError();  // clear the sticky-free error state
    return status;
}

// ------------------------------------------------------------------ RWLock
void RWLock::lock_shared()
{
    std::unique_lock<std::mutex> l(m_);
    cv_.wait(l, [&] { return !writer_ && writers_waiting_ == 0; });
    ++readers_;
}
void RWLock::unlock_shared()
{
    std::unique_lock<std::mutex> l(m_);
    if (--readers_ == 0) cv_.notify_all();
}
void RWLock::lock()
{
    std::unique_lock<std::mutex> l(m_);
    ++writers_waiting_;
    cv_.wait(l, [&] { return !writer_ && readers_ == 0; });
    --writers_waiting_;
    writer_ = true;
}
void RWLock::unlock()
{
    std::unique_lock<std::mutex> l(m_);
    writer_ = false;
    cv_.notify_all();
}

struct SharedGuard {
    RWLock &l;
    explicit SharedGuard(RWLock &l_) : l(l_) { l.lock_shared(); }
    ~SharedGuard() { l.unlock_shared(); }
};
struct ExclusiveGuard {
    RWLock &l;
    explicit ExclusiveGuard(RWLock &l_) : l(l_) { l.lock(); }
    ~ExclusiveGuard() { l.unlock(); }
};

// ------------------------------------------------------------------ Workspace / Context
int Workspace::reserve(size_t d_bytes, size_t h_bytes)
{
    if (d_bytes > d_cap || h_bytes > h_cap) drop_graphs();  // cached graphs point into the old arenas
    if (d_bytes > d_cap) {
        if (d_buf) {
            GF_CUDA(cudaStreamSynchronize(stream), GF_ERR_CUDA);
            cudaFree(d_buf);
            d_buf = nullptr;
            d_cap = 0;
        }
        size_t want = std::max(d_bytes, (size_t)1 << 20);
        GF_CUDA(cudaMalloc((void **)&d_buf, want), GF_ERR_ALLOCATE_GPU_BUFFER);
        d_cap = want;
    }
    if (h_bytes > h_cap) {
        if (h_buf) {
            GF_CUDA(cudaStreamSynchronize(stream), GF_ERR_CUDA);
            cudaFreeHost(h_buf);
            h_buf = nullptr;
            h_cap = 0;
        }
        size_t want = std::max(h_bytes, (size_t)1 << 16);
        GF_CUDA(cudaMallocHost((void **)&h_buf, want), GF_ERR_ALLOCATE_GPU_BUFFER);
        h_cap = want;
    }
    return GF_OK;
}

void Workspace::drop_graphs()
{
    for (auto &g : graphs) cudaGraphExecDestroy(g.exec);
    graphs.clear();
}

int Workspace::reserve_tensor(size_t bytes)
{
    if (bytes <= d_tensor_cap) return GF_OK;
    drop_graphs();  // cached graphs point into the old arena
    if (d_tensor) {
        GF_CUDA(cudaStreamSynchronize(stream), GF_ERR_CUDA);
        cudaFree(d_tensor);
        d_tensor = nullptr;
        d_tensor_cap = 0;
    }
    GF_CUDA(cudaMalloc((void **)&d_tensor, bytes), GF_ERR_ALLOCATE_GPU_BUFFER);
    GF_CUDA(cudaMemset(d_tensor, 0, 256), GF_ERR_CUDA);  // the finalize ticket lives at offset 0, the overlap counters at 64
    d_tensor_cap = bytes;
    op_step = 0;
    done_total[0] = done_total[1] = 0u;
    prep_total[0] = prep_total[1] = 0u;
    return GF_OK;
}

Workspace::~Workspace()
{
    delete multi;
    drop_graphs();
    if (d_tensor) cudaFree(d_tensor);
    if (d_buf) cudaFree(d_buf);
    if (h_buf) cudaFreeHost(h_buf);
    if (stream) cudaStreamDestroy(stream);
}

}  // namespace gf

using namespace gf;

int Context::bind() const
{
    GF_CUDA(cudaSetDevice(device), GF_ERR_INVALID_DEVICE_ID);
    return GF_OK;
}

Workspace *Context::acquire_ws()
{
    {
        std::lock_guard<std::mutex> g(ws_mu);
        if (!ws_free.empty()) {
            Workspace *w = ws_free.back();
            ws_free.pop_back();
            return w;
        }
    }
    Workspace *w = new Workspace();
    if (cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete w;
        return nullptr;
    }
    std::lock_guard<std::mutex> g(ws_mu);
    ws_all.push_back(w);
    return w;
}

void Context::release_ws(Workspace *w)
{
    std::lock_guard<std::mutex> g(ws_mu);
    ws_free.push_back(w);
}

// ---- parallel host copy into the pinned ring
struct CopyPool {
    static constexpr int kHelpers = 3;
    std::mutex mu;
    std::condition_variable cv_start, cv_done;
    std::thread th[kHelpers];
    uint8_t *dst = nullptr;
    const uint8_t *src = nullptr;
    size_t part = 0, bytes = 0;
    uint64_t gen = 0;
    int left = 0;
    bool stop = false;
    CopyPool()
    {
        for (int i = 0; i < kHelpers; ++i)
            th[i] = std::thread([this, i]() {
                uint64_t seen = 0;
                std::unique_lock<std::mutex> lk(mu);
                for (;;) {
                    cv_start.wait(lk, [&] { return stop || gen != seen; });
                    if (stop) return;
                    seen = gen;
                    const size_t off = (size_t)(i + 1) * part;
                    const size_t n = off < bytes ? std::min(part, bytes - off) : 0;
                    uint8_t *d = dst + off;
                    const uint8_t *s = src + off;
                    lk.unlock();
                    if (n) memcpy(d, s, n);
                    lk.lock();
                    if (--left == 0) cv_done.notify_one();
                }
            });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> g(mu);
            stop = true;
        }
        cv_start.notify_all();
        for (auto &t : th) t.join();
    }
    void run(void *d, const void *s, size_t n)
    {
        std::unique_lock<std::mutex> lk(mu);
        dst = (uint8_t *)d;
        src = (const uint8_t *)s;
        bytes = n;
        part = (n / (kHelpers + 1) + 63) / 64 * 64;
        left = kHelpers;
        ++gen;
        cv_start.notify_all();
        lk.unlock();
        memcpy(d, s, std::min(part, n));
        lk.lock();
        cv_done.wait(lk, [&] { return left == 0; });
    }
};

void Context::parallel_copy(void *dst, const void *src, size_t bytes)
{
    if (bytes < ((size_t)1 << 20)) {  // small copies: the hand-off costs more than it saves
        memcpy(dst, src, bytes);
        return;
    }
    if (!copy_pool) copy_pool = new CopyPool();  // (callers hold stage_mu)
    copy_pool->run(dst, src, bytes);
}

int Context::upload(void *dst, const void *src, size_t bytes)
{
    if (bytes == 0) return GF_OK;
    std::lock_guard<std::mutex> g(stage_mu);
    size_t off = 0;
    int k = 0;
    while (off < bytes) {
        size_t n = std::min(stage_bytes, bytes - off);
        int i = k & 1;
        GF_CUDA(cudaEventSynchronize(stage_ev[i]), GF_ERR_WRITE_CUDA_BUFFER);
        parallel_copy(stage[i], (const uint8_t *)src + off, n);
        GF_CUDA(cudaMemcpyAsync((uint8_t *)dst + off, stage[i], n, cudaMemcpyHostToDevice, mut_stream),
                GF_ERR_WRITE_CUDA_BUFFER);
        GF_CUDA(cudaEventRecord(stage_ev[i], mut_stream), GF_ERR_WRITE_CUDA_BUFFER);
        off += n;
        ++k;
    }
    GF_CUDA(cudaStreamSynchronize(mut_stream), GF_ERR_WRITE_CUDA_BUFFER);
    stats.h2d_bytes += bytes;
    return GF_OK;
}

void Context::prof_begin(cudaStream_t st, std::pair<cudaEvent_t, cudaEvent_t> *ev)
{
    ev->first = ev->second = nullptr;
    if (!profiling.load()) return;
    {
        std::lock_guard<std::mutex> g(prof_mu);
        if (!prof_free.empty()) {
            *ev = prof_free.back();
            prof_free.pop_back();
        }
    }
    if (!ev->first) {
        if (cudaEventCreate(&ev->first) != cudaSuccess || cudaEventCreate(&ev->second) != cudaSuccess) {
            ev->first = ev->second = nullptr;
            return;
        }
    }
    cudaEventRecord(ev->first, st);
}

void Context::prof_end(cudaStream_t st, const std::pair<cudaEvent_t, cudaEvent_t> &ev, bool gemm)
{
    if (!ev.first) return;
    cudaEventRecord(ev.second, st);
    std::lock_guard<std::mutex> g(prof_mu);
    (gemm ? prof_pending_gemm : prof_pending).push_back(ev);
}

void Context::prof_collect()
{
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> todo;
    {
        std::lock_guard<std::mutex> g(prof_mu);
        todo.swap(prof_pending);
    }
    for (auto &ev : todo) {
        float ms = 0.f;
        if (cudaEventSynchronize(ev.second) == cudaSuccess && cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) {
            stats.scan_kernel_ns += (uint64_t)((double)ms * 1e6);
            stats.scan_kernel_timed++;
        }
    }
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> todo2;
    {
        std::lock_guard<std::mutex> g(prof_mu);
        todo2.swap(prof_pending_gemm);
    }
    for (auto &ev : todo2) {
        float ms = 0.f;
        if (cudaEventSynchronize(ev.second) == cudaSuccess && cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) {
            stats.gemm_kernel_ns += (uint64_t)((double)ms * 1e6);
            stats.gemm_kernel_timed++;
        }
    }
    std::lock_guard<std::mutex> g(prof_mu);
    prof_free.insert(prof_free.end(), todo.begin(), todo.end());
    prof_free.insert(prof_free.end(), todo2.begin(), todo2.end());
}

int Context::upload_scatter(uint8_t *base, const int *slots, size_t n, size_t rowb, const uint8_t *src)
{
    if (n == 0) return GF_OK;
    std::lock_guard<std::mutex> g(stage_mu);
    // more than a few rows of whole 32-bit words: the rows AND their slot list go into the pinned ring and one
    // kernel per chunk scatters them (it reads the ring over PCIe); otherwise one small copy per row
    const bool by_kernel = n > 4 && rowb % 4 == 0 && ((uintptr_t)base % 4) == 0;
    const size_t per = std::max<size_t>(1, (stage_bytes - (by_kernel ? 16 : 0)) / (rowb + (by_kernel ? 4 : 0)));
    int k = 0;
    for (size_t r0 = 0; r0 < n; r0 += per, ++k) {
        const size_t cnt = std::min(per, n - r0);
        const int i = k & 1;
        GF_CUDA(cudaEventSynchronize(stage_ev[i]), GF_ERR_WRITE_CUDA_BUFFER);
        memcpy(stage[i], src + r0 * rowb, cnt * rowb);
        if (by_kernel) {
            int *st_slots = reinterpret_cast<int *>(stage[i] + (cnt * rowb + 15) / 16 * 16);
            memcpy(st_slots, slots + r0, cnt * sizeof(int));
            GF_CUDA(scatter_rows_launch(reinterpret_cast<const float *>(stage[i]), st_slots, (int)cnt, (int)(rowb / 4),
                                        reinterpret_cast<float *>(base), mut_stream),
                    GF_ERR_WRITE_CUDA_BUFFER);
            stats.other_launches++;
        } else {
            for (size_t j = 0; j < cnt; ++j)
                GF_CUDA(cudaMemcpyAsync(base + (size_t)slots[r0 + j] * rowb, stage[i] + j * rowb, rowb,
                                        cudaMemcpyHostToDevice, mut_stream),
                        GF_ERR_WRITE_CUDA_BUFFER);
        }
        GF_CUDA(cudaEventRecord(stage_ev[i], mut_stream), GF_ERR_WRITE_CUDA_BUFFER);
    }
    GF_CUDA(cudaStreamSynchronize(mut_stream), GF_ERR_WRITE_CUDA_BUFFER);
    stats.h2d_bytes += n * rowb;
    return GF_OK;
}

// zero whole float32 rows (and their bf16 shadow rows) given by element offsets into the slab(s): the list
// travels through the pinned ring, one kernel per chunk
int Context::zero_rows(float *d_slab, void *d_shadow, int shadow_eb, const std::vector<unsigned long long> &elem_off,
                       int dims)
{
    if (elem_off.empty()) return GF_OK;
    std::lock_guard<std::mutex> g(stage_mu);
    const size_t per = stage_bytes / sizeof(unsigned long long);
    int k = 0;
    for (size_t r0 = 0; r0 < elem_off.size(); r0 += per, ++k) {
        const size_t cnt = std::min(per, elem_off.size() - r0);
        const int i = k & 1;
        GF_CUDA(cudaEventSynchronize(stage_ev[i]), GF_ERR_CLEAR_CUDA_BUFFER);
        memcpy(stage[i], elem_off.data() + r0, cnt * sizeof(unsigned long long));
        GF_CUDA(zero_rows_launch(d_slab, d_shadow, shadow_eb, reinterpret_cast<const unsigned long long *>(stage[i]), (int)cnt, dims,
                                 mut_stream),
                GF_ERR_CLEAR_CUDA_BUFFER);
        stats.other_launches++;
        GF_CUDA(cudaEventRecord(stage_ev[i], mut_stream), GF_ERR_CLEAR_CUDA_BUFFER);
    }
    GF_CUDA(cudaStreamSynchronize(mut_stream), GF_ERR_CLEAR_CUDA_BUFFER);
    return GF_OK;
}

Context::~gf_context()
{
    cudaSetDevice(device);
    for (auto &ev : prof_pending) prof_free.push_back(ev);
    for (auto &ev : prof_pending_gemm) prof_free.push_back(ev);
    for (auto &ev : prof_free) {
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    for (Workspace *w : ws_all) delete w;
    for (int i = 0; i < 2; ++i) {
        if (stage[i]) cudaFreeHost(stage[i]);
        if (stage_ev[i]) cudaEventDestroy(stage_ev[i]);
    }
    delete copy_pool;
    if (d_add) cudaFree(d_add);
    if (mut_stream) cudaStreamDestroy(mut_stream);
    for (gf_context *c : children) delete c;  // after ws_all: a multi-device workspace hands its parts back to them
}

gf::DeviceAlloc::~DeviceAlloc()
{
    if (owned && ptr) {
        cudaSetDevice(ctx->device);
        cudaFree(ptr);
    }
}

// ------------------------------------------------------------------ Result
void Result::begin(int q)
{
    batch = q;
    offsets.assign(1, 0);
    id_off.assign(1, 0);
}
void Result::push(float score, uint64_t slot, const std::string &id)
{
    scores.push_back(score);
    slots.push_back(slot);
    id_bytes.append(id);
    id_off.push_back(id_bytes.size());
}
void Result::end_query() { offsets.push_back((int64_t)scores.size()); }

// ------------------------------------------------------------------ Block
std::string Block::id_at(int row) const
{
    if (row < 0 || row >= (int)live.size() || !live[row]) return std::string();
    if (implicit_ids) {
        if (!over.empty()) {
            auto it = over.find(row);
            if (it != over.end()) return it->second;
        }
        long long v = (long long)(id_first + row);
        if (v >= 0 && v < 100000000000ll) {  // "%011lld" by hand: snprintf was most of the id resolution of a search
            char buf[11];
            for (int i = 10; i >= 0; --i) {
                buf[i] = (char)('0' + v % 10);
                v /= 10;
            }
            std::string out;
            out.reserve(id_prefix.size() + 11);
            out.append(id_prefix);
            out.append(buf, 11);
            return out;
        }
        char buf[32];
        snprintf(buf, sizeof buf, "%011lld", v);
        return id_prefix + buf;
    }
    return ids[row];
}

void Block::index_add(const std::string &id, int row)
{
    if (use_index) id_index[id].push_back(row);
}

void Block::rebuild_index()
{
    id_index.clear();
    use_index = true;
    if (implicit_ids) return;
    for (int r = 0; r < next_index; ++r)
        if (live[r]) id_index[ids[r]].push_back(r);
}

void Block::index_del(const std::string &id, int row)
{
    if (!use_index) return;
    auto it = id_index.find(id);
    if (it == id_index.end()) return;
    auto &v = it->second;
    v.erase(std::remove(v.begin(), v.end(), row), v.end());
    if (v.empty()) id_index.erase(it);
}

int Block::find_last(const std::string &id) const
{
    if (implicit_ids) {
        int best = -1;
        for (const auto &kv : over)  // overridden rows carry arbitrary ids (last match wins, block.go:140-144)
            if (kv.first < next_index && live[kv.first] && kv.second == id) best = std::max(best, kv.first);
        if (id.size() != id_prefix.size() + 11 || id.compare(0, id_prefix.size(), id_prefix) != 0) return best;
        long long ord = 0;
        for (size_t i = id_prefix.size(); i < id.size(); ++i) {
            if (id[i] < '0' || id[i] > '9') return best;
            ord = ord * 10 + (id[i] - '0');
        }
        long long row = ord - id_first;
        if (row < 0 || row >= next_index || !live[row] || over.count((int)row)) return best;
        return std::max(best, (int)row);
    }
    if (!use_index) {  // no per-block index: the reference's own linear scan, last match wins (block.go:140-144)
        for (int r = next_index - 1; r >= 0; --r)
            if (live[r] && ids[r] == id) return r;
        return -1;
    }
    auto it = id_index.find(id);
    if (it == id_index.end() || it->second.empty()) return -1;
    return *std::max_element(it->second.begin(), it->second.end());
}

void Block::materialize_ids()
{
    if (!implicit_ids) return;
    std::vector<std::string> v((size_t)capacity());
    for (int r = 0; r < next_index; ++r)
        if (live[r]) v[r] = id_at(r);
    implicit_ids = false;
    over.clear();
    ids.swap(v);
    id_index.clear();
    for (int r = 0; r < next_index; ++r)
        if (live[r]) index_add(ids[r], r);
}

// block.go:59-78
int Block::acquire(const std::string &owner_, int dims_, int precision_, int batch_)
{
    if (!owner.empty()) return GF_ERR_BLOCK_USED;
    if (owner_.empty() || dims_ <= 0 || precision_ <= 0 || batch_ <= 0) return GF_ERR_INVALID_ARGUMENT;
    if ((int64_t)dims_ * precision_ > block_size) return GF_ERR_INVALID_ARGUMENT;
    dims = dims_;
    precision = precision_;
    batch = batch_;
    owner = owner_;
    ids.assign((size_t)capacity(), std::string());
    live.assign((size_t)capacity(), 0);
    live_count = 0;
    implicit_ids = false;
    over.clear();
    use_index = true;
    id_index.clear();
    empty.clear();
    next_index = 0;
    return GF_OK;
}

// block.go:253-270
int Block::release()
{
    std::lock_guard<std::mutex> g(mu);
    int rc = cache->ctx->bind();
    if (rc) return rc;
    GF_CUDA(cudaMemsetAsync(dev, 0, (size_t)block_size, cache->ctx->mut_stream), GF_ERR_CLEAR_CUDA_BUFFER);
    if (shadow)
        GF_CUDA(cudaMemsetAsync(shadow, 0, (size_t)block_size * shadow_eb / 4, cache->ctx->mut_stream),
                GF_ERR_CLEAR_CUDA_BUFFER);
    GF_CUDA(cudaStreamSynchronize(cache->ctx->mut_stream), GF_ERR_CLEAR_CUDA_BUFFER);
    dims = precision = batch = 0;
    owner.clear();
    std::vector<std::string>().swap(ids);
    std::vector<uint8_t>().swap(live);
    id_index.clear();
    empty.clear();
    next_index = 0;
    live_count = 0;
    implicit_ids = false;
    over.clear();
    set = nullptr;
    return GF_OK;
}

// block.go:80-131, with the evident intent of lines 103/115 (whole rows, byte offsets)
int Block::insert(int64_t n, const uint8_t *values, const std::vector<std::string> &ids_)
{
    if (owner.empty()) return GF_ERR_INVALID_ARGUMENT;
    if (n > margin()) return GF_ERR_BLOCK_IS_FULL;
    if (n == 0) return GF_OK;
    std::lock_guard<std::mutex> g(mu);
    Context *ctx = cache->ctx;
    int rc = ctx->bind();
    if (rc) return rc;
    // an implicit-id block keeps its generated ids and records the arbitrary ones row by row, unless that
    // overlay would cover a quarter of the block
    if (implicit_ids && over.size() + (size_t)n > (size_t)capacity() / 4) materialize_ids();
    const bool overlay = implicit_ids;
    auto set_id = [&](int slot, const std::string &id) {
        if (overlay) {
            over[slot] = id;
        } else {
            ids[slot] = id;
            index_add(id, slot);
        }
    };
    const size_t rowb = (size_t)dims * precision;
    const size_t nempty = empty.size();
    const size_t refill = std::min(nempty, (size_t)n);
    if (refill > 0) {
        rc = ctx->upload_scatter(dev, empty.data(), refill, rowb, values);
        if (rc) return GF_ERR_WRITE_CUDA_BUFFER;
    }
    for (size_t i = 0; i < refill; ++i) {
        int slot = empty[i];
        set_id(slot, ids_[i]);
        live[slot] = 1;
        ++live_count;
    }
    if ((size_t)n > nempty) {
        size_t tail = (size_t)n - nempty;
        rc = ctx->upload(dev + (size_t)next_index * rowb, values + nempty * rowb, tail * rowb);
        if (rc) return rc;
        for (size_t i = 0; i < tail; ++i) {
            int slot = next_index + (int)i;
            set_id(slot, ids_[nempty + i]);
            live[slot] = 1;
        }
        live_count += (int)tail;
        next_index += (int)tail;
        empty.clear();
    } else {
        empty.erase(empty.begin(), empty.begin() + n);
    }
    return GF_OK;
}

int Block::insert_staged(int64_t n, const float *d_rows, const uint8_t *d_shadow_rows, const float *d_scales,
                         std::vector<std::string> &ids_, int64_t first, std::vector<int> *placed_rows)
{
    if (owner.empty() || precision != 4) return GF_ERR_INVALID_ARGUMENT;
    if (n > margin()) return GF_ERR_BLOCK_IS_FULL;
    if (n == 0) return GF_OK;
    std::lock_guard<std::mutex> g(mu);
    Context *ctx = cache->ctx;
    int rc = ctx->bind();
    if (rc) return rc;
    if (implicit_ids && over.size() + (size_t)n > (size_t)capacity() / 4) materialize_ids();
    const bool overlay = implicit_ids;
    auto set_id = [&](int slot, std::string &&id) {
        if (overlay) {
            over[slot] = std::move(id);
        } else {
            ids[slot] = std::move(id);
            index_add(ids[slot], slot);
        }
    };
    cudaStream_t st = ctx->mut_stream;
    const size_t rowb = (size_t)dims * 4;
    const bool sh = shadow != nullptr && d_shadow_rows != nullptr;
    const size_t shb = (size_t)dims * shadow_eb;  // bytes of a shadow row
    float *my_scales = nullptr;
    if (sh && shadow_eb == 1 && d_scales != nullptr)
        my_scales = (float *)cache->scales->ptr + (size_t)index * ((size_t)block_size / 512);
    const size_t nempty = empty.size();
    const size_t refill = std::min(nempty, (size_t)n);
    if (refill > 0) {
        // tombstoned slots, in free-list order (block.go:93-108): one scatter kernel per array, the slot list travels
        // through the pinned ring
        std::lock_guard<std::mutex> sg(ctx->stage_mu);
        if (refill * sizeof(int) > ctx->stage_bytes) return GF_ERR_INVALID_ARGUMENT;
        GF_CUDA(cudaEventSynchronize(ctx->stage_ev[0]), GF_ERR_WRITE_CUDA_BUFFER);
        int *slots = reinterpret_cast<int *>(ctx->stage[0]);
        memcpy(slots, empty.data(), refill * sizeof(int));
        GF_CUDA(scatter_rows_launch(d_rows, slots, (int)refill, dims, (float *)dev, st), GF_ERR_WRITE_CUDA_BUFFER);
        if (sh)
            GF_CUDA(scatter_rows_launch((const float *)d_shadow_rows, slots, (int)refill, (int)(shb / 4), (float *)shadow, st),
                    GF_ERR_WRITE_CUDA_BUFFER);
        if (my_scales) GF_CUDA(scatter_rows_launch(d_scales, slots, (int)refill, 1, my_scales, st), GF_ERR_WRITE_CUDA_BUFFER);
        GF_CUDA(cudaEventRecord(ctx->stage_ev[0], st), GF_ERR_WRITE_CUDA_BUFFER);
        ctx->stats.other_launches += 1 + (sh ? 1 : 0) + (my_scales ? 1 : 0);
    }
    for (size_t i = 0; i < refill; ++i) {
        const int slot = empty[i];
        set_id(slot, std::move(ids_[(size_t)first + i]));
        live[slot] = 1;
        ++live_count;
        if (placed_rows) placed_rows->push_back(slot);
    }
    if ((size_t)n > nempty) {
        const size_t tail = (size_t)n - nempty;
        GF_CUDA(cudaMemcpyAsync(dev + (size_t)next_index * rowb, d_rows + nempty * dims, tail * rowb,
                                cudaMemcpyDeviceToDevice, st),
                GF_ERR_WRITE_CUDA_BUFFER);
        if (sh)
            GF_CUDA(cudaMemcpyAsync(shadow + (size_t)next_index * shb, d_shadow_rows + nempty * shb, tail * shb,
                                    cudaMemcpyDeviceToDevice, st),
                    GF_ERR_WRITE_CUDA_BUFFER);
        if (my_scales)
            GF_CUDA(cudaMemcpyAsync(my_scales + next_index, d_scales + nempty, tail * 4, cudaMemcpyDeviceToDevice, st),
                    GF_ERR_WRITE_CUDA_BUFFER);
        for (size_t i = 0; i < tail; ++i) {
            const int slot = next_index + (int)i;
            set_id(slot, std::move(ids_[(size_t)first + nempty + i]));
            live[slot] = 1;
            if (placed_rows) placed_rows->push_back(slot);
        }
        live_count += (int)tail;
        next_index += (int)tail;
        empty.clear();
    } else {
        empty.erase(empty.begin(), empty.begin() + n);
    }
    return GF_OK;
}

int Block::insert_synthetic(int64_t n, const gf::SynthSpec &seed, int64_t first_ordinal, int normalize, const std::string &prefix)
{
    if (owner.empty() || precision != 4) return GF_ERR_INVALID_ARGUMENT;
    if (n > capacity() - next_index) return GF_ERR_BLOCK_IS_FULL;
    if (n == 0) return GF_OK;
    std::lock_guard<std::mutex> g(mu);
    Context *ctx = cache->ctx;
    int rc = ctx->bind();
    if (rc) return rc;
    bool can_implicit = (next_index == 0 && empty.empty()) ||
                        (implicit_ids && id_prefix == prefix && id_first + next_index == first_ordinal);
    if (!can_implicit) materialize_ids();
    GF_CUDA(synth_fill_launch((float *)dev + (size_t)next_index * dims, n, dims, seed, first_ordinal, normalize,
                              ctx->mut_stream),
            GF_ERR_WRITE_CUDA_BUFFER);
    ctx->stats.other_launches++;
    if (can_implicit) {
        if (next_index == 0) {
            implicit_ids = true;
            id_prefix = prefix;
            id_first = first_ordinal;
            std::vector<std::string>().swap(ids);
        }
        for (int64_t i = 0; i < n; ++i) live[next_index + i] = 1;
    } else {
        char buf[32];
        for (int64_t i = 0; i < n; ++i) {
            snprintf(buf, sizeof buf, "%011lld", (long long)(first_ordinal + i));
            int slot = next_index + (int)i;
            ids[slot] = prefix + buf;
            live[slot] = 1;
            index_add(ids[slot], slot);
        }
    }
    live_count += (int)n;
    next_index += (int)n;
    return GF_OK;
}

// block.go:135-165.  `found[i]` is set for every request id removed here; ids already
// flagged by an earlier block are skipped (set.go:176-190 forwards only the remainder).
int Block::remove(const std::vector<std::string> &ids_, std::vector<uint8_t> &found)
{
    if (owner.empty()) return GF_OK;
    std::lock_guard<std::mutex> g(mu);
    Context *ctx = cache->ctx;
    int rc = ctx->bind();
    if (rc) return rc;
    const size_t rowb = (size_t)dims * precision;
    bool any = false;
    // targets is a map: a repeated id is looked up once (block.go:136-139)
    std::unordered_map<std::string, size_t> first;
    first.reserve(ids_.size() * 2);
    for (size_t i = 0; i < ids_.size(); ++i) first.emplace(ids_[i], i);
    for (size_t i = 0; i < ids_.size(); ++i) {
        if (found[i]) continue;
        if (first[ids_[i]] != i) continue;
        int row = find_last(ids_[i]);
        if (row < 0) continue;
        GF_CUDA(cudaMemsetAsync(dev + (size_t)row * rowb, 0, rowb, ctx->mut_stream), GF_ERR_CLEAR_CUDA_BUFFER);
        if (shadow && precision == 4)
            GF_CUDA(cudaMemsetAsync(shadow + (size_t)row * dims * shadow_eb, 0, (size_t)dims * shadow_eb, ctx->mut_stream),
                    GF_ERR_CLEAR_CUDA_BUFFER);
        if (!implicit_ids) {
            index_del(ids_[i], row);
            ids[row].clear();
        } else if (!over.empty()) {
            over.erase(row);
        }
        live[row] = 0;
        --live_count;
        empty.push_back(row);
        found[i] = 1;
        any = true;
    }
    if (any) GF_CUDA(cudaStreamSynchronize(ctx->mut_stream), GF_ERR_CLEAR_CUDA_BUFFER);
    return GF_OK;
}

int Block::remove_rows(const std::vector<int> &rows, std::vector<unsigned long long> *zero_later)
{
    if (rows.empty()) return GF_OK;
    std::lock_guard<std::mutex> g(mu);
    Context *ctx = cache->ctx;
    const size_t rowb = (size_t)dims * precision;
    for (int row : rows) {
        if (row < 0 || row >= next_index || !live[row]) continue;
        if (zero_later) {  // the caller zeroes all rows of the Delete with one kernel (element offset in the slab)
            zero_later->push_back((unsigned long long)((dev - (uint8_t *)cache->slab->ptr) / 4) +
                                  (unsigned long long)row * dims);
        } else {
            GF_CUDA(cudaMemsetAsync(dev + (size_t)row * rowb, 0, rowb, ctx->mut_stream), GF_ERR_CLEAR_CUDA_BUFFER);
            if (shadow && precision == 4)
                GF_CUDA(cudaMemsetAsync(shadow + (size_t)row * dims * shadow_eb, 0, (size_t)dims * shadow_eb,
                                        ctx->mut_stream),
                        GF_ERR_CLEAR_CUDA_BUFFER);
        }
        if (!implicit_ids) {
            index_del(ids[row], row);
            ids[row].clear();
        } else if (!over.empty()) {
            over.erase(row);
        }
        live[row] = 0;
        --live_count;
        empty.push_back(row);
    }
    return GF_OK;
}

// ------------------------------------------------------------------ Set
bool Set::find_implicit(const std::string &id, int *pos, int *row) const
{
    if (implicit_ranges.empty()) return false;
    const size_t pl = implicit_prefix.size();
    if (id.size() != pl + 11 || id.compare(0, pl, implicit_prefix) != 0) return false;
    int64_t ord = 0;
    for (size_t i = pl; i < id.size(); ++i) {
        if (id[i] < '0' || id[i] > '9') return false;
        ord = ord * 10 + (id[i] - '0');
    }
    size_t lo = 0, hi = implicit_ranges.size();
    while (lo < hi) {  // last range with first <= ord
        size_t mid = (lo + hi) / 2;
        if (implicit_ranges[mid].first <= ord) lo = mid + 1; else hi = mid;
    }
    if (lo == 0) return false;
    const ImplicitRange &r = implicit_ranges[lo - 1];
    if (ord >= r.first + r.count) return false;
    const Block *b = blocks[r.pos];
    const int rr = (int)(ord - r.first);
    if (!b->implicit_ids || rr >= b->next_index || !b->live[rr]) return false;
    if (!b->over.empty() && b->over.count(rr)) return false;  // that row now carries an arbitrary id
    *pos = r.pos;
    *row = rr;
    return true;
}

// First block (Set.Blocks order) holding the id, last row inside it (set.go:167-170, block.go:140-144).
bool Set::find(const std::string &id, int *pos, int *row) const
{
    loc_sync();
    if (has_dups) {
        for (size_t p = 0; p < blocks.size(); ++p) {
            int r = blocks[p]->find_last(id);
            if (r >= 0) {
                *pos = (int)p;
                *row = r;
                return true;
            }
        }
        return false;
    }
    if (find_implicit(id, pos, row)) return true;
    auto it = loc.find(id);
    if (it == loc.end()) return false;
    *pos = (int)(it->second >> 32);
    *row = (int)(it->second & 0xffffffffu);
    return true;
}

void Set::enter_dup_mode()
{
    if (has_dups) return;
    has_dups = true;  // ids are not unique: fall back to the reference's block-by-block walk
    loc.clear();
    for (Block *b : blocks) b->rebuild_index();
}

void Set::note_id(const std::string &id, int pos, int row)
{
    if (has_dups) return;
    int p, r;
    if (find_implicit(id, &p, &r) || !loc.emplace(id, ((uint64_t)pos << 32) | (uint32_t)row).second) enter_dup_mode();
}

// an implicit-id block is about to take arbitrary ids: give it explicit ids and move them to the locator
void Set::materialize_block(int pos)
{
    Block *b = blocks[pos];
    if (!b->implicit_ids) return;
    b->materialize_ids();
    for (size_t i = 0; i < implicit_ranges.size(); ++i)
        if (implicit_ranges[i].pos == pos) {
            implicit_ranges.erase(implicit_ranges.begin() + i);
            break;
        }
    if (has_dups) {
        b->rebuild_index();
        return;
    }
    for (int r = 0; r < b->next_index; ++r) {
        if (!b->live[r]) continue;
        const uint64_t where = ((uint64_t)pos << 32) | (uint32_t)r;
        auto ins = loc.emplace(b->ids[r], where);
        if (!ins.second && ins.first->second != where) {  // (an overridden row is in the locator already)
            enter_dup_mode();
            return;
        }
    }
}

bool Set::slot_to_block(uint32_t slot, int *block_pos, int *row) const
{
    if (cap <= 0) return false;
    uint64_t gb = slot / (uint32_t)cap;
    if ((int)(gb % (uint64_t)shard_world) != shard_rank) return false;
    uint64_t pos = gb / (uint64_t)shard_world;
    if (pos >= blocks.size()) return false;
    *block_pos = (int)pos;
    *row = (int)(slot % (uint32_t)cap);
    return true;
}

bool Set::locate_slot(uint32_t slot, const Block **b, int *row) const
{
    if (cap <= 0) return false;
    if (is_multi()) {
        const uint64_t gb = slot / (uint32_t)cap;
        const Set *sh = shards[(size_t)(gb % shards.size())];
        int bp;
        if (!sh->slot_to_block(slot, &bp, row)) return false;
        *b = sh->blocks[bp];
        return true;
    }
    int bp;
    if (!slot_to_block(slot, &bp, row)) return false;
    *b = blocks[bp];
    return true;
}

static std::atomic<uint64_t> g_set_version{0};
uint64_t gf::next_set_version() { return ++g_set_version; }

int Set::rebuild_segments()
{
    const std::vector<Segment> old_segs = h_segs;
    const int old_tile_rows = tile_rows;
    const bool old_shadow_ok = shadow_ok, old_tensor_ok = tensor_ok;
    h_segs.clear();
    seg_block.clear();
    int tr = scan_tile_rows(dims);
    aligned16 = (dims % 4 == 0) && (cache->block_size % 16 == 0) && precision == 4;
    tile_rows = (aligned16 && tr > 0) ? tr : 64;
    uint64_t gblocks = (uint64_t)blocks.size() * shard_world;
    if (gblocks * (uint64_t)cap >= 0xFFFFFFFFull) {
        set_error_detail("set too large for 32-bit slots");
        return GF_ERR_INVALID_ARGUMENT;
    }
    uint32_t tiles = 0;
    scanned_rows = 0;
    for (size_t p = 0; p < blocks.size(); ++p) {
        Block *b = blocks[p];
        if (b->next_index == 0) continue;  // block.go:186-189
        Segment s;
        s.base = (const float *)b->dev;
        s.rows = (uint32_t)b->next_index;
        s.slot_base = slot_base_of((int)p);
        s.tile_start = tiles;
        s.pad = 0;
        tiles += (s.rows + tile_rows - 1) / tile_rows;
        scanned_rows += b->next_index;
        h_segs.push_back(s);
        seg_block.push_back((int)p);
    }
    total_tiles = tiles;
    // An Add that only refilled tombstones leaves the table as it was: the device copy, the version and every
    // recorded graph stay valid (under churn a re-capture per Add cost more than the Add)
    if (d_segs != nullptr && old_tile_rows == tile_rows && old_segs.size() == h_segs.size() && !h_segs.empty() &&
        memcmp(old_segs.data(), h_segs.data(), h_segs.size() * sizeof(Segment)) == 0) {
        shadow_ok = old_shadow_ok;
        tensor_ok = old_tensor_ok;
        return GF_OK;
    }
    version = next_set_version();  // process-wide unique: a cached graph can never match a different table
    if (h_segs.empty()) return GF_OK;
    int rc = ctx->bind();
    if (rc) return rc;
    if ((int)h_segs.size() > d_segs_cap) {
        if (d_segs) cudaFree(d_segs);
        d_segs = nullptr;
        int want = std::max((int)h_segs.size() * 2, 64);
        // one allocation: Segment[want] | gtile_start[want + 1] | slot_base[want]
        size_t bytes = (size_t)want * sizeof(Segment) + (size_t)(2 * want + 2) * sizeof(uint32_t);
        GF_CUDA(cudaMalloc((void **)&d_segs, bytes), GF_ERR_ALLOCATE_GPU_BUFFER);
        d_segs_cap = want;
        d_gtile_start = reinterpret_cast<uint32_t *>(d_segs + want);
        d_seg_slot_base = d_gtile_start + want + 1;
    }
    // host image of the three arrays, one upload
    const size_t n = h_segs.size();
    std::vector<uint32_t> aux(2 * n + 1);
    uint32_t gt = 0;
    for (size_t i = 0; i < n; ++i) {
        aux[i] = gt;
        gt += (h_segs[i].rows + 127) / 128;
        aux[n + 1 + i] = h_segs[i].slot_base;
    }
    aux[n] = gt;
    total_gtiles = gt;
    tensor_ok = precision == 4 && dims % 32 == 0 && dims >= 32 && (cache->block_size % ((int64_t)dims * 4)) == 0 &&
                ((uintptr_t)cache->slab->ptr % 128) == 0;
    // one K block of the shadow pass is a 128-byte swizzle row: 64 bf16 or 128 int8 elements
    shadow_ok = tensor_ok && cache->shadow != nullptr && dims % (128 / cache->shadow_eb) == 0 &&
                ((uintptr_t)cache->shadow->ptr % 128) == 0;
    GF_CUDA(cudaMemcpyAsync(d_segs, h_segs.data(), n * sizeof(Segment), cudaMemcpyHostToDevice, ctx->mut_stream),
            GF_ERR_WRITE_CUDA_BUFFER);
    GF_CUDA(cudaMemcpyAsync(d_gtile_start, aux.data(), (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice,
                            ctx->mut_stream),
            GF_ERR_WRITE_CUDA_BUFFER);
    GF_CUDA(cudaMemcpyAsync(d_seg_slot_base, aux.data() + n + 1, n * sizeof(uint32_t), cudaMemcpyHostToDevice,
                            ctx->mut_stream),
            GF_ERR_WRITE_CUDA_BUFFER);
    GF_CUDA(cudaStreamSynchronize(ctx->mut_stream), GF_ERR_WRITE_CUDA_BUFFER);
    return GF_OK;
}

// The tensor path's certificate needs an upper bound of |row| over the set: an atomicMax over the
// rows that were just written (deletes never lower it: the bound stays conservative).
int Set::note_rows(const float *d_rows, int64_t rows)
{
    if (precision != 4 || rows <= 0) return GF_OK;
    if (!d_small) {
        GF_CUDA(cudaMalloc((void **)&d_small, 64), GF_ERR_ALLOCATE_GPU_BUFFER);
        GF_CUDA(cudaMemsetAsync(d_small, 0, 64, ctx->mut_stream), GF_ERR_CUDA);
    }
    if (cache->shadow != nullptr && dims % 4 == 0 && (cache->shadow_eb == 2 || dims % 128 == 0)) {
        // same element index in both slabs: row j of block i sits at the same (i, j) in the shadow
        const size_t elem = (size_t)(d_rows - (const float *)cache->slab->ptr);
        if (cache->shadow_eb == 2) {
            GF_CUDA(shadow_rows_launch(d_rows, rows, dims, (uint16_t *)cache->shadow->ptr + elem, d_small, ctx->mut_stream),
                    GF_ERR_CUDA);
        } else {
            const size_t block_floats = (size_t)cache->block_size / 4;
            const size_t blk = elem / block_floats, row0 = (elem % block_floats) / (size_t)dims;
            float *sc = (float *)cache->scales->ptr + blk * ((size_t)cache->block_size / 512) + row0;
            GF_CUDA(shadow_rows_i8_launch(d_rows, rows, dims, (uint8_t *)cache->shadow->ptr + elem, sc, d_small,
                                          ctx->mut_stream),
                    GF_ERR_CUDA);
        }
    } else {
        GF_CUDA(rownorm_range_launch(d_rows, rows, dims, d_small, ctx->mut_stream), GF_ERR_CUDA);
    }
    ctx->stats.other_launches++;
    return GF_OK;
}

// Mutations wait for every search already enqueued on the device (device-resident
// searches return before their kernels finish).
static int quiesce(Context *ctx)
{
    int rc = ctx->bind();
    if (rc) return rc;
    GF_CUDA(cudaDeviceSynchronize(), GF_ERR_CUDA);
    return GF_OK;
}

// Stage `rows` float32 rows on the device and quantise their shadow there (no set lock held): the context's staging
// area is rows | shadow rows | row scales.  Returns the three device pointers; the caller holds ctx->add_mu.
static int stage_rows(Set *s, const uint8_t *values, int64_t rows, float **d_rows, uint8_t **d_shadow_rows,
                      float **d_scales)
{
    Context *ctx = s->ctx;
    Cache *cache = s->cache;
    const int dims = s->dims;
    const size_t rowb = (size_t)dims * 4;
    const bool sh = cache->shadow != nullptr && dims % 4 == 0 && (cache->shadow_eb == 2 || dims % 128 == 0);
    const size_t shb = sh ? (size_t)dims * cache->shadow_eb : 0;
    const size_t o_sh = ((size_t)rows * rowb + 255) / 256 * 256;
    const size_t o_sc = o_sh + ((size_t)rows * shb + 255) / 256 * 256;
    const size_t need = o_sc + (size_t)rows * 4 + 256;
    if (need > ctx->d_add_cap) {
        if (ctx->d_add) {
            GF_CUDA(cudaStreamSynchronize(ctx->mut_stream), GF_ERR_CUDA);
            cudaFree(ctx->d_add);
            ctx->d_add = nullptr;
            ctx->d_add_cap = 0;
        }
        GF_CUDA(cudaMalloc((void **)&ctx->d_add, need), GF_ERR_ALLOCATE_GPU_BUFFER);
        ctx->d_add_cap = need;
    }
    *d_rows = (float *)ctx->d_add;
    *d_shadow_rows = sh ? ctx->d_add + o_sh : nullptr;
    *d_scales = (sh && cache->shadow_eb == 1) ? (float *)(ctx->d_add + o_sc) : nullptr;
    int rc = ctx->upload(*d_rows, values, (size_t)rows * rowb);  // pinned ring, ends with a stream sync
    if (rc) return rc;
    if (sh && cache->shadow_eb == 2) {
        GF_CUDA(shadow_rows_launch(*d_rows, rows, dims, *d_shadow_rows, s->d_small, ctx->mut_stream), GF_ERR_CUDA);
    } else if (sh) {
        GF_CUDA(shadow_rows_i8_launch(*d_rows, rows, dims, *d_shadow_rows, *d_scales, s->d_small, ctx->mut_stream), GF_ERR_CUDA);
    } else {
        GF_CUDA(rownorm_range_launch(*d_rows, rows, dims, s->d_small, ctx->mut_stream), GF_ERR_CUDA);
    }
    ctx->stats.other_launches++;
    GF_CUDA(cudaStreamSynchronize(ctx->mut_stream), GF_ERR_CUDA);
    return GF_OK;
}

// set.go:77-116.  Three phases (see gf_set::mut_mu): (A) no set lock: rows to a device staging area through the
// pinned ring, shadow rows + row scales + the certificate's bounds computed there; (B) exclusive lock: blocks
// acquired, rows placed device-to-device (tombstoned slots first, in free-list order, then the tail), ids and live
// flags set, segment table rebuilt; (C) no set lock: the id locator learns the new ids.  Searches wait for (B) only.
int Set::add(int64_t n, const uint8_t *values, const std::vector<std::string> &ids_)
{
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    if (n == 0) return GF_OK;
    if (is_multi()) return multi_add(n, values, ids_);
    std::lock_guard<std::mutex> mg(mut_mu);
    loc_sync();
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;  // DestroySet won
    if (precision != 4) return add_plain(n, values, ids_);
    int rc = ctx->bind();
    if (rc) return rc;
    if (!d_small) {  // (searches test this pointer: publish it under the exclusive lock)
        unsigned int *p = nullptr;
        GF_CUDA(cudaMalloc((void **)&p, 64), GF_ERR_ALLOCATE_GPU_BUFFER);
        GF_CUDA(cudaMemset(p, 0, 64), GF_ERR_CUDA);
        ExclusiveGuard g(lock);
        d_small = p;
    }
    // cache.go:111-113 before anything is written: the whole request must fit
    {
        int64_t empty = 0;
        for (Block *b : blocks) empty += b->margin();
        if (n > empty) {
            const int64_t block_length = cache->block_size / ((int64_t)dims * precision);
            if (block_length <= 0) return GF_ERR_NOT_ENOUGH_BLOCKS;
            const int64_t block_num = (n - empty + block_length - 1) / block_length;
            std::lock_guard<std::mutex> cg(cache->mu);
            int64_t free_blocks = 0;
            for (auto &ub : cache->all_blocks)
                if (!ub->is_owned()) ++free_blocks;
            if (free_blocks < block_num) return GF_ERR_NOT_ENOUGH_BLOCKS;
        }
    }
    const size_t rowb = (size_t)dims * 4;
    const int64_t chunk_rows = std::max<int64_t>(1, ((int64_t)64 << 20) / (int64_t)rowb);  // 64 MiB of rows per chunk
    std::vector<std::string> ids(ids_);  // moved into the blocks chunk by chunk
    for (int64_t c0 = 0; c0 < n; c0 += chunk_rows) {
        const int64_t cn = std::min(chunk_rows, n - c0);
        std::lock_guard<std::mutex> ag(ctx->add_mu);
        float *d_rows = nullptr, *d_scales = nullptr;
        uint8_t *d_shadow_rows = nullptr;
        const auto t_a = std::chrono::steady_clock::now();
        rc = stage_rows(this, values + (size_t)c0 * rowb, cn, &d_rows, &d_shadow_rows, &d_scales);  // (A)
        if (rc) return rc;
        const auto t_b = std::chrono::steady_clock::now();
        struct Placed {
            int bp;
            std::vector<int> rows;
        };
        std::vector<Placed> placed;
        loc_sync();  // (the previous chunk's ids: (B) reads has_dups and may move a block's ids into the locator)
        {
            ExclusiveGuard g(lock);  // (B)
            rc = quiesce(ctx);
            if (rc) return rc;
            int64_t empty = 0;
            for (Block *b : blocks) empty += b->margin();
            if (cn > empty) {
                const int64_t block_length = cache->block_size / ((int64_t)dims * precision);
                const int64_t block_num = (cn - empty + block_length - 1) / block_length;
                std::vector<Block *> got;
                {
                    std::lock_guard<std::mutex> cg(cache->mu);
                    for (auto &ub : cache->all_blocks) {  // cache.go:104-115
                        if ((int64_t)got.size() == block_num) break;
                        if (!ub->is_owned()) got.push_back(ub.get());
                    }
                    if ((int64_t)got.size() < block_num) return GF_ERR_NOT_ENOUGH_BLOCKS;
                    for (size_t gi = 0; gi < got.size(); ++gi) {
                        Block *b = got[gi];
                        rc = b->acquire(name, dims, precision, batch);
                        if (rc) {  // hand back what this call took: nothing may stay owned outside `blocks`
                            for (size_t u = 0; u < gi; ++u) got[u]->release();
                            return rc;
                        }
                        b->set = this;
                        b->use_index = has_dups;
                    }
                }
                blocks.insert(blocks.end(), got.begin(), got.end());
            }
            int64_t offset = 0, remain = cn;
            for (size_t bp = 0; bp < blocks.size() && remain > 0; ++bp) {
                Block *b = blocks[bp];
                const int64_t length = std::min<int64_t>(b->margin(), remain);
                if (length <= 0) continue;
                // (an implicit-id block records arbitrary ids in its overlay; when the insert decides to turn it
                // into explicit strings the locator has to learn its generated ids first)
                if (b->implicit_ids && b->over.size() + (size_t)length > (size_t)b->capacity() / 4) materialize_block((int)bp);
                placed.push_back(Placed{(int)bp, {}});
                placed.back().rows.reserve((size_t)length);
                const size_t shb = d_shadow_rows ? (size_t)dims * cache->shadow_eb : 0;
                rc = b->insert_staged(length, d_rows + (size_t)offset * dims,
                                      d_shadow_rows ? d_shadow_rows + (size_t)offset * shb : nullptr,
                                      d_scales ? d_scales + offset : nullptr, ids, c0 + offset, &placed.back().rows);
                if (rc) {
                    rebuild_segments();
                    return rc;
                }
                offset += length;
                remain -= length;
            }
            rc = rebuild_segments();  // (ends with a sync of the mutation stream: the copies above are done)
            if (rc) return rc;
            GF_CUDA(cudaStreamSynchronize(ctx->mut_stream), GF_ERR_WRITE_CUDA_BUFFER);
        }
        const auto t_c = std::chrono::steady_clock::now();
        // (C) the locator learns the ids -- from the blocks' own copies, on a helper thread for anything but a
        // small request: the caller gets its Add back (and stages the next one) meanwhile.  Searches never read the
        // locator; every mutation, Read and DestroySet joins the job first (loc_sync).
        auto learn = [this](const std::vector<Placed> &pl, size_t rows_) {
            if (!has_dups) loc.reserve(loc.size() + rows_);
            for (const Placed &p : pl)
                for (int row : p.rows) note_id(blocks[(size_t)p.bp]->id_at(row), p.bp, row);
        };
        if (cn >= 512) {
            std::lock_guard<std::mutex> jg(loc_job_mu);
            loc_job = std::async(std::launch::async, [learn, pl = std::move(placed), cn]() { learn(pl, (size_t)cn); });
        } else {
            learn(placed, (size_t)cn);
        }
        const auto t_d = std::chrono::steady_clock::now();
        auto ns = [](std::chrono::steady_clock::time_point x, std::chrono::steady_clock::time_point y) {
            return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(y - x).count();
        };
        ctx->stats.add_stage_ns += ns(t_a, t_b);
        ctx->stats.add_place_ns += ns(t_b, t_c);
        ctx->stats.add_index_ns += ns(t_c, t_d);
        ctx->stats.add_rows += (uint64_t)cn;
    }
    return GF_OK;
}

// the reference's own sequence for the precisions the search path does not serve (whole Add under the exclusive lock)
int Set::add_plain(int64_t n, const uint8_t *values, const std::vector<std::string> &ids_)
{
    ExclusiveGuard g(lock);
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    int rc = quiesce(ctx);
    if (rc) return rc;
    int64_t empty = 0;
    for (Block *b : blocks) empty += b->margin();
    if (n > empty) {
        int64_t remain = n - empty;
        int64_t block_length = cache->block_size / ((int64_t)dims * precision);
        if (block_length <= 0) return GF_ERR_NOT_ENOUGH_BLOCKS;
        int64_t block_num = (remain + block_length - 1) / block_length;
        std::vector<Block *> got;
        {
            std::lock_guard<std::mutex> cg(cache->mu);
            for (auto &ub : cache->all_blocks) {  // cache.go:104-115
                if ((int64_t)got.size() == block_num) break;
                if (!ub->is_owned()) got.push_back(ub.get());
            }
            if ((int64_t)got.size() < block_num) return GF_ERR_NOT_ENOUGH_BLOCKS;
            for (size_t gi = 0; gi < got.size(); ++gi) {
                Block *b = got[gi];
                rc = b->acquire(name, dims, precision, batch);
                if (rc) {
                    for (size_t u = 0; u < gi; ++u) got[u]->release();
                    return rc;
                }
                b->set = this;
                b->use_index = has_dups;
            }
        }
        blocks.insert(blocks.end(), got.begin(), got.end());
    }
    const size_t rowb = (size_t)dims * precision;
    int64_t offset = 0, remain = n;
    for (size_t bp = 0; bp < blocks.size(); ++bp) {
        Block *b = blocks[bp];
        int64_t length = std::min<int64_t>(b->margin(), remain);
        if (length > 0) {
            if (b->implicit_ids && b->over.size() + (size_t)length > (size_t)b->capacity() / 4) materialize_block((int)bp);
            std::vector<std::string> sub(ids_.begin() + offset, ids_.begin() + offset + length);
            const std::vector<int> refilled(b->empty.begin(),
                                            b->empty.begin() + std::min<size_t>(b->empty.size(), (size_t)length));
            const int tail0 = b->next_index;
            rc = b->insert(length, values + (size_t)offset * rowb, sub);
            if (rc) {
                rebuild_segments();
                return rc;
            }
            for (size_t i = 0; i < refilled.size(); ++i) note_id(sub[i], (int)bp, refilled[i]);
            for (int64_t i = (int64_t)refilled.size(); i < length; ++i)
                note_id(sub[i], (int)bp, tail0 + (int)(i - (int64_t)refilled.size()));
            offset += length;
            remain -= length;
        }
        if (remain <= 0) break;
    }
    return rebuild_segments();
}

int Set::add_synthetic(int64_t n, const gf::SynthSpec &seed, int64_t first_ordinal, int normalize, const std::string &prefix)
{
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    if (precision != 4) return GF_ERR_UNSUPPORTED;
    if (n == 0) return GF_OK;
    if (is_multi()) return multi_add_synthetic(n, seed, first_ordinal, normalize, prefix);
    std::lock_guard<std::mutex> mg(mut_mu);
    loc_sync();
    ExclusiveGuard g(lock);
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    int rc = quiesce(ctx);
    if (rc) return rc;
    // synthetic rows only append (never refill tombstones): count tail room
    int64_t room = 0;
    for (Block *b : blocks) room += b->capacity() - b->next_index;
    if (n > room) {
        int64_t remain = n - room;
        int64_t block_length = cache->block_size / ((int64_t)dims * precision);
        int64_t block_num = (remain + block_length - 1) / block_length;
        std::vector<Block *> got;
        std::lock_guard<std::mutex> cg(cache->mu);
        for (auto &ub : cache->all_blocks) {
            if ((int64_t)got.size() == block_num) break;
            if (!ub->is_owned()) got.push_back(ub.get());
        }
        if ((int64_t)got.size() < block_num) return GF_ERR_NOT_ENOUGH_BLOCKS;
        for (size_t gi = 0; gi < got.size(); ++gi) {
            Block *b = got[gi];
            rc = b->acquire(name, dims, precision, batch);
            if (rc) {
                for (size_t u = 0; u < gi; ++u) got[u]->release();
                return rc;
            }
            b->set = this;
            b->use_index = has_dups;
            // implicit ids: do not keep capacity() empty strings per block
            std::vector<std::string>().swap(b->ids);
            b->implicit_ids = true;
            b->id_prefix = prefix;
            b->id_first = 0;
        }
        blocks.insert(blocks.end(), got.begin(), got.end());
    }
    if (!implicit_ranges.empty() && implicit_prefix != prefix) enter_dup_mode();  // two id families: walk the blocks
    implicit_prefix = prefix;
    int64_t done = 0;
    for (size_t bp = 0; bp < blocks.size(); ++bp) {
        Block *b = blocks[bp];
        int64_t length = std::min<int64_t>(b->capacity() - b->next_index, n - done);
        if (length > 0) {
            if (b->implicit_ids && b->next_index == 0) b->id_first = first_ordinal + done;
            const int tail0 = b->next_index;
            // generated ids that do not continue this block's own range end its implicit life: the locator has
            // to learn the ids it holds before Block::insert_synthetic turns them into strings
            if (b->implicit_ids && b->next_index > 0 &&
                !(b->id_prefix == prefix && b->id_first + b->next_index == first_ordinal + done))
                materialize_block((int)bp);
            const bool was_implicit = b->implicit_ids;
            rc = b->insert_synthetic(length, seed, first_ordinal + done, normalize, prefix);
            if (rc) return rc;
            note_rows((const float *)b->dev + (size_t)tail0 * dims, length);
            if (b->implicit_ids) {
                bool merged = false;
                for (ImplicitRange &r : implicit_ranges)
                    if (r.pos == (int)bp) {
                        r.count = b->next_index;
                        merged = true;
                    }
                if (!merged) {
                    ImplicitRange r{b->id_first, b->next_index, (int)bp};
                    auto it = std::lower_bound(implicit_ranges.begin(), implicit_ranges.end(), r,
                                               [](const ImplicitRange &x, const ImplicitRange &y) { return x.first < y.first; });
                    if ((it != implicit_ranges.end() && it->first < r.first + r.count) ||
                        (it != implicit_ranges.begin() && (it - 1)->first + (it - 1)->count > r.first))
                        enter_dup_mode();  // overlapping ordinal ranges
                    implicit_ranges.insert(it, r);
                }
                if (!loc.empty() && !has_dups) {  // an explicit id may collide with a generated one
                    char buf[32];
                    for (int64_t i = 0; i < length && !has_dups; ++i) {
                        snprintf(buf, sizeof buf, "%011lld", (long long)(first_ordinal + done + i));
                        if (loc.count(prefix + buf)) enter_dup_mode();
                    }
                }
            } else {
                if (was_implicit) materialize_block((int)bp);
                for (int64_t i = 0; i < length; ++i) note_id(b->ids[tail0 + i], (int)bp, tail0 + (int)i);
            }
            done += length;
        }
        if (done >= n) break;
    }
    GF_CUDA(cudaStreamSynchronize(ctx->mut_stream), GF_ERR_WRITE_CUDA_BUFFER);
    return rebuild_segments();
}

// set.go:163-193
int Set::remove(const std::vector<std::string> &ids_, std::vector<uint8_t> &found)
{
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    found.assign(ids_.size(), 0);
    if (ids_.empty()) return GF_OK;
    if (is_multi()) return multi_remove(ids_, found);
    std::lock_guard<std::mutex> mg(mut_mu);
    loc_sync();
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    // (A) no set lock: which rows?  set.go:163-193 walks the blocks and block.go:136-147 scans every IDs[]; the
    // locator answers the same question (first block holding the id, last row in it) in O(1) per id, and searches
    // never look at it
    std::unordered_map<std::string, size_t> first;
    first.reserve(ids_.size() * 2);
    for (size_t i = 0; i < ids_.size(); ++i) first.emplace(ids_[i], i);  // a repeated id is looked up once
    std::unordered_map<int, std::vector<int>> per_block;
    for (size_t i = 0; i < ids_.size(); ++i) {
        if (first[ids_[i]] != i) continue;
        int pos, row;
        if (!find(ids_[i], &pos, &row)) continue;
        per_block[pos].push_back(row);
        found[i] = 1;
        if (!has_dups && (!blocks[pos]->implicit_ids || blocks[pos]->over.count(row))) loc.erase(ids_[i]);
    }
    if (per_block.empty()) return GF_OK;
    size_t nrows = 0;
    for (auto &kv : per_block) nrows += kv.second.size();
    const bool batched = nrows > 4 && precision == 4;  // one zeroing kernel instead of two memsets per row
    std::vector<unsigned long long> zero_off;
    if (batched) zero_off.reserve(nrows);
    // (B) exclusive: tombstone (ids, live flags, free lists) and zero the rows on the device
    ExclusiveGuard g(lock);
    int rc = quiesce(ctx);
    if (rc) return rc;
    for (auto &kv : per_block) {
        rc = blocks[kv.first]->remove_rows(kv.second, batched ? &zero_off : nullptr);
        if (rc) return rc;
    }
    if (batched) {
        rc = ctx->zero_rows((float *)cache->slab->ptr, cache->shadow ? cache->shadow->ptr : nullptr, cache->shadow_eb,
                            zero_off, dims);
        if (rc) return rc;
    } else {
        GF_CUDA(cudaStreamSynchronize(ctx->mut_stream), GF_ERR_CLEAR_CUDA_BUFFER);
    }
    return GF_OK;  // NextIndex never shrinks (block.go:159-160): the segment table is unchanged
}

// interface.go:47-50 (TODO in the reference, set.go:195-198): overwrite in place
int Set::update(int64_t n, const uint8_t *values, const std::vector<std::string> &ids_, std::vector<uint8_t> &found)
{
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    found.assign(ids_.size(), 0);
    if (n == 0) return GF_OK;
    if (is_multi()) return multi_update(n, values, ids_, found);
    std::lock_guard<std::mutex> mg(mut_mu);
    loc_sync();
    ExclusiveGuard g(lock);
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    int rc = quiesce(ctx);
    if (rc) return rc;
    const size_t rowb = (size_t)dims * precision;
    for (int64_t i = 0; i < n; ++i) {
        int pos, row;
        if (!find(ids_[i], &pos, &row)) continue;
        Block *b = blocks[pos];
        rc = ctx->upload(b->dev + (size_t)row * rowb, values + (size_t)i * rowb, rowb);
        if (rc) return rc;
        note_rows((const float *)b->dev + (size_t)row * dims, 1);
        found[i] = 1;
    }
    GF_CUDA(cudaStreamSynchronize(ctx->mut_stream), GF_ERR_WRITE_CUDA_BUFFER);
    return GF_OK;
}

// interface.go:52-55 (TODO in the reference, set.go:216-219)
int Set::read(const std::vector<std::string> &ids_, uint8_t *out, std::vector<uint8_t> &found)
{
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    found.assign(ids_.size(), 0);
    if (is_multi()) return multi_read(ids_, out, found);
    std::lock_guard<std::mutex> mg(mut_mu);  // (the id locator is only stable between mutations)
    loc_sync();
    SharedGuard g(lock);
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    int rc = ctx->bind();
    if (rc) return rc;
    const size_t rowb = (size_t)dims * precision;
    for (size_t i = 0; i < ids_.size(); ++i) {
        memset(out + i * rowb, 0, rowb);
        int pos, row;
        if (!find(ids_[i], &pos, &row)) continue;
        GF_CUDA(cudaMemcpy(out + i * rowb, blocks[pos]->dev + (size_t)row * rowb, rowb, cudaMemcpyDeviceToHost),
                GF_ERR_READ_CUDA_BUFFER);
        ctx->stats.d2h_bytes += rowb;
        found[i] = 1;
    }
    return GF_OK;
}

// Not in the reference (NextIndex never shrinks, block.go:159-160, so deleted rows are scanned for
// ever): pack the live rows of every block to its front, in order, and shrink NextIndex.  The
// relative slot order of live rows is kept, so results change only where tombstones used to take
// places in a block's top-limit (block.go:236-243).
int Set::compact(int64_t *reclaimed)
{
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    if (precision != 4) return GF_ERR_UNSUPPORTED;
    if (is_multi()) return multi_compact(reclaimed);
    std::lock_guard<std::mutex> mg(mut_mu);
    loc_sync();
    ExclusiveGuard g(lock);
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    int rc = quiesce(ctx);
    if (rc) return rc;
    int64_t freed = 0;
    float *d_tmp = nullptr;
    int *d_idx = nullptr;
    for (size_t bp = 0; bp < blocks.size(); ++bp) {
        Block *b = blocks[bp];
        if (b->next_index == 0 || b->live_count == b->next_index) continue;
        materialize_block((int)bp);
        std::vector<int> rows;
        rows.reserve((size_t)b->live_count);
        for (int r = 0; r < b->next_index; ++r)
            if (b->live[r]) rows.push_back(r);
        const int n_live = (int)rows.size();
        if (!d_tmp) {
            GF_CUDA(cudaMalloc((void **)&d_tmp, (size_t)cache->block_size), GF_ERR_ALLOCATE_GPU_BUFFER);
            GF_CUDA(cudaMalloc((void **)&d_idx, (size_t)cap * sizeof(int) + 16), GF_ERR_ALLOCATE_GPU_BUFFER);
        }
        std::lock_guard<std::mutex> bg(b->mu);
        if (n_live > 0) {
            GF_CUDA(cudaMemcpyAsync(d_idx, rows.data(), (size_t)n_live * sizeof(int), cudaMemcpyHostToDevice,
                                    ctx->mut_stream),
                    GF_ERR_WRITE_CUDA_BUFFER);
            GF_CUDA(gather_rows_launch((const float *)b->dev, d_idx, n_live, dims, d_tmp, ctx->mut_stream),
                    GF_ERR_CUDA);
            GF_CUDA(cudaMemcpyAsync(b->dev, d_tmp, (size_t)n_live * dims * 4, cudaMemcpyDeviceToDevice,
                                    ctx->mut_stream),
                    GF_ERR_WRITE_CUDA_BUFFER);
            ctx->stats.other_launches++;
        }
        GF_CUDA(cudaMemsetAsync(b->dev + (size_t)n_live * dims * 4, 0, (size_t)(b->next_index - n_live) * dims * 4,
                                ctx->mut_stream),
                GF_ERR_CLEAR_CUDA_BUFFER);
        if (b->shadow) {  // the packed rows moved: convert them again, clear the freed tail
            if (n_live > 0) {
                rc = note_rows((const float *)b->dev, n_live);
                if (rc) return rc;
            }
            GF_CUDA(cudaMemsetAsync(b->shadow + (size_t)n_live * dims * b->shadow_eb, 0,
                                    (size_t)(b->next_index - n_live) * dims * b->shadow_eb, ctx->mut_stream),
                    GF_ERR_CLEAR_CUDA_BUFFER);
        }
        GF_CUDA(cudaStreamSynchronize(ctx->mut_stream), GF_ERR_CUDA);  // rows[] is reused by the next block
        for (int j = 0; j < n_live; ++j) {
            if (rows[j] != j) b->ids[j] = std::move(b->ids[rows[j]]);
            if (!has_dups) loc[b->ids[j]] = ((uint64_t)bp << 32) | (uint32_t)j;
        }
        for (int r = n_live; r < b->next_index; ++r) b->ids[r].clear();
        std::fill(b->live.begin(), b->live.begin() + n_live, (uint8_t)1);
        std::fill(b->live.begin() + n_live, b->live.end(), (uint8_t)0);
        freed += b->next_index - n_live;
        b->next_index = n_live;
        b->live_count = n_live;
        b->empty.clear();
        if (b->use_index) b->rebuild_index();
    }
    if (d_tmp) cudaFree(d_tmp);
    if (d_idx) cudaFree(d_idx);
    if (reclaimed) *reclaimed = freed;
    return rebuild_segments();
}

int Set::export_rows(uint8_t *out_values, uint64_t values_cap, char *out_ids, uint64_t ids_cap, uint64_t *out_off,
                     int64_t *out_rows, uint64_t *out_id_bytes)
{
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    if (is_multi()) return multi_export(out_values, values_cap, out_ids, ids_cap, out_off, out_rows, out_id_bytes);
    SharedGuard g(lock);
    int rc = ctx->bind();
    if (rc) return rc;
    const size_t rowb = (size_t)dims * precision;
    int64_t rows = 0;
    uint64_t idb = 0;
    for (const Block *b : blocks)
        for (int r = 0; r < b->next_index; ++r)
            if (b->live[r]) {
                ++rows;
                idb += b->id_at(r).size();
            }
    if (out_rows) *out_rows = rows;
    if (out_id_bytes) *out_id_bytes = idb;
    if (!out_values && !out_ids && !out_off) return GF_OK;  // size query
    if (!out_values || !out_ids || !out_off) return GF_ERR_INVALID_ARGUMENT;
    if (values_cap < (uint64_t)rows * rowb || ids_cap < idb) return GF_ERR_BUFFER_WRITE_OUT_OF_RANGE;
    std::vector<uint8_t> tmp;
    int64_t k = 0;
    uint64_t io = 0;
    out_off[0] = 0;
    for (const Block *b : blocks) {
        if (b->next_index == 0) continue;
        tmp.resize((size_t)b->next_index * rowb);
        GF_CUDA(cudaMemcpy(tmp.data(), b->dev, tmp.size(), cudaMemcpyDeviceToHost), GF_ERR_READ_CUDA_BUFFER);
        ctx->stats.d2h_bytes += tmp.size();
        for (int r = 0; r < b->next_index; ++r) {
            if (!b->live[r]) continue;
            memcpy(out_values + (size_t)k * rowb, tmp.data() + (size_t)r * rowb, rowb);
            const std::string id = b->id_at(r);
            memcpy(out_ids + io, id.data(), id.size());
            io += id.size();
            out_off[++k] = io;
        }
    }
    return GF_OK;
}

// set.go:202-214
int Set::destroy()
{
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    if (is_multi()) return multi_destroy();
    std::lock_guard<std::mutex> mg(mut_mu);
    loc_sync();
    ExclusiveGuard g(lock);
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    int rc = quiesce(ctx);
    if (rc) return rc;
    destroyed = true;
    for (Block *b : blocks) {
        rc = b->release();
        if (rc) return rc;
    }
    blocks.clear();
    loc.clear();
    implicit_ranges.clear();
    has_dups = false;
    h_segs.clear();
    seg_block.clear();
    total_tiles = 0;
    scanned_rows = 0;
    if (d_segs) {
        cudaFree(d_segs);
        d_segs = nullptr;
        d_segs_cap = 0;
    }
    if (dev_ws) {
        ctx->release_ws(dev_ws);
        dev_ws = nullptr;
    }
    if (d_small) {
        cudaFree(d_small);
        d_small = nullptr;
    }
    d_gtile_start = d_seg_slot_base = nullptr;
    total_gtiles = 0;
    tensor_ok = false;
    return GF_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// partial-list bytes one scan pass of qp queries needs
static size_t partial_bytes(const ScanPlan &p, int qp, int K)
{
    return (size_t)p.groups * p.grid_x * qp * K * sizeof(unsigned long long);
}

bool Set::tensor_eligible(int q, int K, bool per_segment, const unsigned long long *d_upper, int path) const
{
    if (path == 1 || per_segment || d_upper != nullptr || !tensor_ok || d_small == nullptr) return false;
    if (K > 512) return false;                               // K' = K + slack must fit the 1024-key re-score
    if (total_gtiles < 2 * 256) return false;                // small sets: the sample would be the set
    if (path >= 2 && path <= 5) return true;
    // the shadow pass streams a half (bf16) or a quarter (int8) of the bytes: it wins for every batch size
    if (shadow_ok) return true;
    // one scan pass cannot hold the batch, or K is beyond what the per-CTA sorted lists do cheaply
    return q > scan_max_q(dims) || K > 32;
}

// The tensor path is ~12 small launches around one big kernel, and the gaps between them were 20 % of
// a 10-query step: everything after the query preparation is recorded once per (set state, q, K) as a
// CUDA graph and replayed.  The graph only touches workspace-owned buffers (the caller's query and
// result pointers stay outside: one eager prep kernel before, two small copies after), recording
// happens on the workspace's own stream (the caller's may be the legacy default stream, which cannot
// be captured), replay goes into the caller's stream.
int Set::topk_device(Workspace *ws, cudaStream_t stream, const float *d_queries, int q, int64_t q_stride,
                     int64_t l_stride, int K, bool per_segment, const unsigned long long *d_upper,
                     unsigned long long *d_out, int path, unsigned int *d_flags, bool bake_io)
{
    const bool tensor = tensor_eligible(q, K, per_segment, d_upper, path);
    if (!tensor || q > kGemmMaxQ || ws->graphs_broken || ctx->profiling.load())
        return topk_eager(ws, stream, d_queries, q, q_stride, l_stride, K, per_segment, d_upper, d_out, path, d_flags);
    Stats &st = ctx->stats;
    const bool device_fixup = d_flags == nullptr;  // nobody to report an uncertified query to: repair it here
    const int mode = (use_shadow(path) ? cache->shadow_eb : 0) | (device_fixup ? 4 : 0) | (two_pass_only(path) ? 8 : 0);
    // bake_io (Search's own staging buffers, stable per workspace): the query preparation and the copy of
    // keys + flags are nodes of the graph too -- ONE stream operation per call
    bake_io = bake_io && d_flags != nullptr && q_stride == dims && l_stride == 1 && stream == ws->stream;
    // a multi-device set's step ends with the gather-to-root exchange: baked into the graph like the buffers
    const void *xkey = (bake_io && ws->xchg != nullptr) ? (const void *)ws->xchg->peers : nullptr;
    // The one-pass chain is three launches that write straight into the caller's buffers: replaying a graph would
    // need an eager preparation before it and a copy of the results behind it -- as many stream operations, and
    // the copy sits on the device's critical path.  Only Search's own stable buffers (bake_io) are worth a graph.
    if (!bake_io && onepass_applies(q, K, path, device_fixup))
        return topk_eager(ws, stream, d_queries, q, q_stride, l_stride, K, per_segment, d_upper, d_out, path, d_flags);
    const size_t res_b = align_up((size_t)q * K * 8, 256);
    for (auto &g : ws->graphs) {
        if (g.set == this && g.version == version && g.q == q && g.K == K && g.mode == mode &&
            g.src == (bake_io ? (const void *)d_queries : nullptr) && g.dst == (bake_io ? (void *)d_out : nullptr) &&
            g.xkey == xkey) {
            g.stamp = ++ws->graph_clock;
            st.scan_launches += g.d_scan;
            st.gemm_launches += g.d_gemm;
            st.merge_launches += g.d_merge;
            st.other_launches += g.d_other + (bake_io ? 0 : 1);
            st.tensor_passes += g.d_tensor_passes;
            st.rows_scanned += g.d_rows;
            if (bake_io) {
                GF_CUDA(cudaGraphLaunch(g.exec, stream), GF_ERR_CUDA);
                ws->xchg_used = g.xchg_in_graph;
                return GF_OK;
            }
            int rc = topk_tensor(ws, stream, d_queries, q, q_stride, l_stride, K, nullptr, nullptr, 1, path, device_fixup);
            if (rc) return rc;
            GF_CUDA(cudaGraphLaunch(g.exec, stream), GF_ERR_CUDA);
            // d_out / d_flags may be device memory or pinned host memory
            if (d_flags && (uint8_t *)d_flags == (uint8_t *)d_out + res_b &&
                (uint8_t *)ws->t_flags == (uint8_t *)ws->t_res + res_b) {
                GF_CUDA(cudaMemcpyAsync(d_out, ws->t_res, res_b + (size_t)q * 4, cudaMemcpyDefault, stream), GF_ERR_CUDA);
            } else {
                GF_CUDA(cudaMemcpyAsync(d_out, ws->t_res, (size_t)q * K * 8, cudaMemcpyDefault, stream), GF_ERR_CUDA);
                if (d_flags)
                    GF_CUDA(cudaMemcpyAsync(d_flags, ws->t_flags, (size_t)q * 4, cudaMemcpyDefault, stream), GF_ERR_CUDA);
            }
            return GF_OK;
        }
    }
    // first call of this shape: run it eagerly (sizes the scratch arenas, sets kernel attributes) ...
    int rc = topk_eager(ws, stream, d_queries, q, q_stride, l_stride, K, per_segment, d_upper, d_out, path, d_flags);
    if (rc) return rc;
    // ... then record the part after the query preparation for the next calls
    GF_CUDA(cudaStreamSynchronize(stream), GF_ERR_CUDA);
    const uint64_t s0 = st.scan_launches, g0 = st.gemm_launches, m0 = st.merge_launches, o0 = st.other_launches,
                   t0 = st.tensor_passes, r0 = st.rows_scanned;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(ws->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        ws->graphs_broken = true;
        return GF_OK;
    }
    int crc = GF_OK;
    if (bake_io) crc = topk_tensor(ws, ws->stream, d_queries, q, q_stride, l_stride, K, nullptr, nullptr, 1, path, device_fixup);
    ws->wrote_direct = false;
    ws->direct_out = bake_io ? d_out : nullptr;
    ws->direct_flags = bake_io ? d_flags : nullptr;
    if (crc == GF_OK)
        crc = topk_tensor(ws, ws->stream, d_queries, q, q_stride, l_stride, K, nullptr, nullptr, 2, path, device_fixup);
    ws->direct_out = nullptr;
    ws->direct_flags = nullptr;
    const bool graph_xchg = ws->xchg_used;  // (the eager first run above used it too, or neither did)
    if (crc == GF_OK && bake_io && !ws->wrote_direct) {
        const bool contig = (uint8_t *)d_flags == (uint8_t *)d_out + res_b &&
                            (uint8_t *)ws->t_flags == (uint8_t *)ws->t_res + res_b;
        cudaError_t me = cudaMemcpyAsync(d_out, ws->t_res, contig ? res_b + (size_t)q * 4 : (size_t)q * K * 8,
                                         cudaMemcpyDefault, ws->stream);
        if (me == cudaSuccess && !contig)
            me = cudaMemcpyAsync(d_flags, ws->t_flags, (size_t)q * 4, cudaMemcpyDefault, ws->stream);
        if (me != cudaSuccess) crc = GF_ERR_CUDA;
    }
    cudaError_t ce = cudaStreamEndCapture(ws->stream, &graph);
    Workspace::CachedGraph cg{};  // what the recording pass "launched" did not run: take its counters back
    cg.d_scan = st.scan_launches - s0;
    cg.d_gemm = st.gemm_launches - g0;
    cg.d_merge = st.merge_launches - m0;
    cg.d_other = st.other_launches - o0;
    cg.d_tensor_passes = st.tensor_passes - t0;
    cg.d_rows = st.rows_scanned - r0;
    st.scan_launches -= cg.d_scan;
    st.gemm_launches -= cg.d_gemm;
    st.merge_launches -= cg.d_merge;
    st.other_launches -= cg.d_other;
    st.tensor_passes -= cg.d_tensor_passes;
    st.rows_scanned -= cg.d_rows;
    cudaGraphExec_t exec = nullptr;
    if (crc == GF_OK && ce == cudaSuccess && graph != nullptr) ce = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (crc != GF_OK || ce != cudaSuccess || exec == nullptr) {
        cudaGetLastError();
        ws->graphs_broken = true;
        return GF_OK;  // the eager result above stands
    }
    if (ws->graphs.size() >= 16) {  // evict the least recently used shape
        size_t victim = 0;
        for (size_t i = 1; i < ws->graphs.size(); ++i)
            if (ws->graphs[i].stamp < ws->graphs[victim].stamp) victim = i;
        cudaGraphExecDestroy(ws->graphs[victim].exec);
        ws->graphs.erase(ws->graphs.begin() + victim);
    }
    cg.set = this;
    cg.version = version;
    cg.q = q;
    cg.K = K;
    cg.mode = mode;
    cg.src = bake_io ? (const void *)d_queries : nullptr;
    cg.dst = bake_io ? (void *)d_out : nullptr;
    cg.xkey = xkey;
    cg.xchg_in_graph = graph_xchg;
    cg.exec = exec;
    cg.stamp = ++ws->graph_clock;
    ws->graphs.push_back(cg);
    return GF_OK;
}

bool Set::onepass_applies(int q, int K, int path, bool device_fixup, int *kp_out) const
{
    const int seb = use_shadow(path) ? cache->shadow_eb : 0;
    const int kp = seb == 1   ? std::min(1024, std::max(K + 96, 3 * K + 64))
                   : seb == 2 ? std::min(1024, std::max(K + 64, 2 * K + 32))
                              : std::min(1024, std::max(K + 32, K + K / 2));
    // groups at or above the cut (all of them are re-scored, like the two-pass chain re-scores everything a small
    // batch collects): deep enough that the certificate's margin, score[K] - score[cut], clears delta on small sets
    const int kp1 = std::max(kp, 6 * K + 64);
    if (kp_out) *kp_out = kp1;
    // The GROUPTOP pass is 1 - 3 % slower per tile than COLLECT and its tail reads 8 bytes per 32 rows and query,
    // the chain around it ~17 us shorter.  Measured (one-pass vs two-pass, ms per step): 4 M rows 0.312 / 0.337 at
    // one query, 0.340 / 0.338 at ten; 10 M rows 0.748 / 0.766 and 0.806 / 0.781.  Path 2 / 3 (tests, sweeps) pins
    // the one-pass chain at any size.
    // (round 2, with a step's tail running next to the following step's pass: 4 M rows 0.312 / 0.342 at ten queries,
    // 8 M rows 0.598 / 0.648; sixteen queries 0.341 / 0.350 and 0.670 / 0.659; 12.5 M rows 0.985 / 0.988 at ten --
    // profiles/r2n_onepass_sweep.txt)
    const int64_t row_limit = q <= 10 ? 12000000 : (q <= 16 ? 6000000 : 3000000);
    const bool size_ok = path == 2 || path == 3 || scanned_rows <= row_limit;
    return !two_pass_only(path) && !device_fixup && q <= kGtMaxQ && K <= kGtMaxK && kp1 <= kGtCap / 2 &&
           size_ok && (uint64_t)total_gtiles * 4u < (1ull << 27);
}

int Set::topk_eager(Workspace *ws, cudaStream_t stream, const float *d_queries, int q, int64_t q_stride,
                    int64_t l_stride, int K, bool per_segment, const unsigned long long *d_upper,
                    unsigned long long *d_out, int path, unsigned int *d_flags)
{
    if (tensor_eligible(q, K, per_segment, d_upper, path))
        return topk_tensor(ws, stream, d_queries, q, q_stride, l_stride, K, d_out, d_flags, 0, path, d_flags == nullptr);
    if (d_flags) GF_CUDA(cudaMemsetAsync(d_flags, 0, (size_t)q * sizeof(unsigned int), stream), GF_ERR_CUDA);
    return topk_scan(ws, stream, d_queries, q, q_stride, l_stride, K, per_segment, d_upper, d_out);
}

// Tensor-core candidate pass + exact re-score + certificate, in passes of <= 256 queries:
//   prep (pad + |q|^2) -> TILEMAX over a strided sample of tiles -> per-query threshold ->
//   COLLECT over all tiles (threshold + atomic append) -> best K' candidates -> canonical re-score
//   -> sort, emit K, certify.  Uncertified queries are re-run by one fix-up scan pass on the device
//   (up to kFixupMaxQ per call); what is still flagged is reported through d_flags.
int Set::topk_tensor(Workspace *ws, cudaStream_t stream, const float *d_queries, int q, int64_t q_stride,
                     int64_t l_stride, int K, unsigned long long *d_out, unsigned int *d_flags, int phase, int path,
                     bool device_fixup)
{
    const int nseg = (int)h_segs.size();
    const int seb = use_shadow(path) ? cache->shadow_eb : 0;  // 0 float32 rows (tf32), 2 bf16 shadow, 1 int8 shadow
    const bool bf16 = seb != 0;                                // (a shadow pass: measured-residual certificate)
    // K' re-scored per query when the batch is too large to re-score everything collected: the certificate
    // needs exact[K] > approx[K'] + delta; delta ~ 2.2e-3 with tf32, ~3.5e-3 with bf16, ~8e-3 .. 2e-2 with int8
    const int kp = seb == 1   ? std::min(1024, std::max(K + 96, 3 * K + 64))
                   : seb == 2 ? std::min(1024, std::max(K + 64, 2 * K + 32))
                              : std::min(1024, std::max(K + 32, K + K / 2));
    const unsigned cap = kGemmCandCap;
    // sample: G strided tiles; threshold = m-th largest tile maximum so that ~target rows pass
    // whole waves of the grid: one tile per SM on a 1 M-row set, up to 6 on the largest
    int waves = (int)std::max<uint32_t>(1, std::min<uint32_t>(6, total_gtiles / (48u * (uint32_t)ctx->sms)));
    int G = std::min(kGemmMaxSampleTiles, waves * ctx->sms);
    G = (int)std::min<uint32_t>((uint32_t)G, total_gtiles);
    const uint32_t stride = total_gtiles / (uint32_t)G;
    // rows per query that should pass the COLLECT threshold: enough that K' of them are there despite the
    // sampling noise of the threshold estimate, few enough that appends stay rare
    const double target = std::max(512.0, 2.0 * kp + 64.0);
    int m = (int)std::ceil(target * G / (double)total_gtiles);
    m = std::max(2, std::min(m, G / 2));

    const int nq_max = gemm_pad_queries(std::min(q, kGemmMaxQ));
    // small batches re-score EVERY collected candidate (~`target` rows per query) and skip the selection
    auto rescore_all = [&](int qp_) { return (size_t)qp_ * (size_t)target * (size_t)dims * 4 <= ((size_t)32 << 20); };
    const bool all = rescore_all(std::min(q, kGemmMaxQ));  // one decision per call
    // The large-batch chain ends with ONE tail launch (collect_tail_kernel) per group of up to kTailBatchPasses passes:
    // thresholds, counters, norms and candidate buffers of a group's queries stay side by side until then.
    const int tail_span = all ? nq_max : std::min((q + kGemmMaxQ - 1) / kGemmMaxQ, kTailBatchPasses) * kGemmMaxQ;
    size_t off = 0;
    auto carve = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) / 256 * 256;
        return o;
    };
    const size_t o_ticket = carve(4);  // always offset 0: zeroed when the arena is allocated, left zero by every call
    const size_t o_qpad = carve((size_t)nq_max * dims * 4);
    const size_t o_qn = carve((size_t)tail_span * 4);
    const size_t o_qbf = carve(bf16 ? (size_t)nq_max * dims * seb : 0);
    const size_t o_qres = carve((size_t)tail_span * 4);
    const size_t o_qsc = carve((size_t)nq_max * 4);
    const size_t o_tmax = carve((size_t)kGemmMaxSampleTiles * nq_max * 4);
    const size_t o_thr = carve((size_t)tail_span * 4);
    const size_t o_cnt = carve((size_t)tail_span * 4);
    const size_t o_cand = carve((size_t)tail_span * cap * 8);
    // selection depth of the large-batch chain: the tail sorts the best kp_sel candidates by approximate score and
    // decides PER QUERY how many of them the exact re-score must cover
    const int kp_sel = std::min(1024, std::max(512, 2 * kp));
    const size_t o_exact = carve(all ? (size_t)nq_max * (size_t)cap * 8 : 0);
    const size_t o_map = carve((size_t)kFixupMaxQ * 4);
    const size_t o_count = carve(4);
    const size_t o_qall = carve((size_t)q * dims * 4);  // row-major copy of all queries for the fix-up scan
    const size_t o_res = carve((size_t)q * K * 8);     // result keys when the caller's buffer must stay outside (phase 2)
    const size_t o_flags = carve((size_t)q * 4);       // right behind the keys: one copy returns both
    // One-pass path (small batches whose caller takes the certificate flags): no threshold sample -- the big
    // kernel emits the best two keys of every 32-row group + a histogram, ONE tail kernel does the rest
    // (prep -> GROUPTOP -> tail: 3 launches instead of 6).  Search paths 4 / 5 pin the threshold + COLLECT chain.
    int kp1 = kp;
    const bool onepass = onepass_applies(q, K, path, device_fixup, &kp1);
    const uint32_t gt_groups = total_gtiles * 4u;
    const size_t o_gtop = carve(onepass ? (size_t)nq_max * gt_groups * 8 : 0);
    // second scratch set of the overlapped one-pass chain (see the launch below)
    const size_t o_qpad2 = carve(onepass ? (size_t)nq_max * dims * 4 : 0);
    const size_t o_qn2 = carve(onepass ? (size_t)nq_max * 4 : 0);
    const size_t o_qbf2 = carve(onepass && bf16 ? (size_t)nq_max * dims * seb : 0);
    const size_t o_qres2 = carve(onepass ? (size_t)nq_max * 4 : 0);
    const size_t o_qsc2 = carve(onepass ? (size_t)nq_max * 4 : 0);
    const size_t o_gtop2 = carve(onepass ? (size_t)nq_max * gt_groups * 8 : 0);
    int rc = ws->reserve_tensor(off);
    if (rc) return rc;
    uint8_t *T = ws->d_tensor;
    float *d_qpad = (float *)(T + o_qpad);
    float *d_qn = (float *)(T + o_qn);
    void *d_qbf = bf16 ? (void *)(T + o_qbf) : nullptr;
    float *d_qres = bf16 ? (float *)(T + o_qres) : nullptr;
    float *d_qsc = (float *)(T + o_qsc);
    const float err_coeff = seb == 1 ? kI8ErrCoeff : seb == 2 ? kAccErrCoeff : kTf32ErrCoeff;
    float *d_tmax = (float *)(T + o_tmax);
    float *d_thr = (float *)(T + o_thr);
    unsigned int *d_cnt = (unsigned int *)(T + o_cnt);
    unsigned long long *d_cand = (unsigned long long *)(T + o_cand);
    unsigned long long *d_exact = (unsigned long long *)(T + o_exact);
    unsigned int *d_fl = (unsigned int *)(T + o_flags);
    int *d_map = (int *)(T + o_map);
    unsigned int *d_count = (unsigned int *)(T + o_count);
    unsigned int *d_ticket = (unsigned int *)(T + o_ticket);
    float *d_qall = (float *)(T + o_qall);
    ws->t_res = (unsigned long long *)(T + o_res);
    ws->t_flags = d_fl;
    if (phase == 2) d_out = ws->t_res;

    GemmLaunch g{};
    g.d_segs = d_segs;
    g.d_gtile_start = d_gtile_start;
    g.nseg = nseg;
    g.dims = dims;
    g.slab_base = (const float *)cache->slab->ptr;
    g.slab_rows = (uint64_t)(cache->slab->size / ((uint64_t)dims * 4));
    g.d_queries_padded = d_qpad;
    g.shadow_base = bf16 ? cache->shadow->ptr : nullptr;
    g.shadow_eb = seb;
    g.d_queries_shadow = d_qbf;
    g.d_row_scales = seb == 1 ? (const float *)cache->scales->ptr : nullptr;
    g.d_q_scale = d_qsc;
    g.block_size = cache->block_size;
    g.d_thr = d_thr;
    g.d_cand = d_cand;
    g.d_cnt = d_cnt;
    g.cap = cap;
    g.d_tilemax = d_tmax;
    g.sms = ctx->sms;

    const int fixq = std::min(kFixupMaxQ, scan_max_q(dims));
    bool flags_direct = false;
    for (int p0 = 0; p0 < q; p0 += kGemmMaxQ) {
        const int qp = std::min(kGemmMaxQ, q - p0);
        const int nq = gemm_pad_queries(qp);
        g.nq = nq;
        g.q = qp;
        FixupPlan fix{};  // the last pass's finalize also lists the flagged queries of the whole call
        if (device_fixup && p0 + kGemmMaxQ >= q) fix = FixupPlan{d_fl, q, fixq, d_map, d_count, d_small + 1, d_ticket};
        fix.flagged = d_small + 3;
        // (the overlapped one-pass chain launches its own preparation into the scratch set of the step's parity)
        const bool overlap = onepass && phase == 0 && d_flags != nullptr && qp == q && !ctx->profiling.load();
        // (|q|^2 and the quantisation residual go to the pass's place in its tail group's arrays)
        const int pq = (all || onepass) ? 0 : p0 % (kTailBatchPasses * kGemmMaxQ);
        if (phase != 2 && !overlap)
            GF_CUDA(prep_queries_launch(d_queries + (size_t)p0 * q_stride, q_stride, l_stride, qp, nq, dims, d_qpad,
                                        d_qn + pq, d_qall + (size_t)p0 * dims, d_qbf, seb, bf16 ? d_qres + pq : nullptr,
                                        d_qsc, stream),
                    GF_ERR_CUDA);
        if (phase == 1) continue;
        if (onepass) {
            // Overlap across steps (eager device-resident calls, one pass): step i's tail -- re-score, certificate,
            // cross-rank exchange, ~10 us and mostly latency -- needs nothing of step i+1 and step i+1's candidate pass
            // nothing of it, once their scratch is separate.  Two scratch sets alternate; the preparation kernel of
            // step i+1 does not wait for tail i (it waits for tail i-1's completion counter, the last user of its
            // set), so the pass that depends on it starts while tail i is still running; tail i+1 stays behind
            // tail i (same result rows, same exchange inbox) through the other counter.  A recorded graph bakes its
            // arguments in and keeps the plain serial chain.
            const int par = overlap ? (int)(ws->op_step & 1ull) : 0;
            unsigned int *d_done = (unsigned int *)(T + 64);
            float *o_pad = par ? (float *)(T + o_qpad2) : d_qpad;
            float *o_qn = par ? (float *)(T + o_qn2) : d_qn;
            void *o_qbf = bf16 ? (par ? (void *)(T + o_qbf2) : d_qbf) : nullptr;
            float *o_qres = bf16 ? (par ? (float *)(T + o_qres2) : d_qres) : nullptr;
            float *o_qsc = par ? (float *)(T + o_qsc2) : d_qsc;
            if (overlap) {
                GF_CUDA(prep_queries_launch(d_queries, q_stride, l_stride, qp, nq, dims, o_pad, o_qn, nullptr, o_qbf, seb,
                                            o_qres, o_qsc, stream, d_done + par, ws->done_total[par], d_done + 2 + par),
                        GF_ERR_CUDA);
                ws->prep_total[par] += (unsigned int)nq;
                g.d_wait_ctr = d_done + 2 + par;
                g.wait_expect = ws->prep_total[par];
                g.d_queries_padded = o_pad;
                g.d_queries_shadow = o_qbf;
                g.d_q_scale = o_qsc;
            }
            g.mode = kGemmGroupTop;
            g.tile_first = 0;
            g.tile_stride = 1;
            g.tile_count = total_gtiles;
            g.d_gtop = (uint2 *)(T + (par ? o_gtop2 : o_gtop));
            {
                std::pair<cudaEvent_t, cudaEvent_t> ev;
                ctx->prof_begin(stream, &ev);
                cudaError_t le = gemm_launch(g, stream);
                ctx->prof_end(stream, ev, true);
                if (le != cudaSuccess) return cuda_fail(le, "gemm_launch", GF_ERR_CUDA);
            }
            GtTail t{};
            t.d_gtop = g.d_gtop;
            t.groups = gt_groups;
            t.d_segs = d_segs;
            t.d_gtile_start = d_gtile_start;
            t.nseg = nseg;
            t.dims = dims;
            t.d_queries_padded = overlap ? o_pad : d_qpad;
            t.d_q_scale = seb == 1 ? (overlap ? o_qsc : d_qsc) : nullptr;
            t.kp = kp1;
            t.K = K;
            t.q = qp;
            t.rows_total = scanned_rows;
            t.d_qnorm2 = overlap ? o_qn : d_qn;
            t.d_maxnorm2_bits = d_small;
            t.err_coeff = err_coeff;
            t.d_qres2 = overlap ? o_qres : d_qres;
            if (overlap) {
                t.wait_ctr = d_done + (par ^ 1);
                t.wait_expect = ws->done_total[par ^ 1];
                t.done_ctr = d_done + par;
            }
            t.d_maxres2_bits = d_small + 2;
            t.d_out = d_out + (size_t)p0 * K;
            t.d_flags = d_fl + p0;
            t.d_flagged = d_small + 3;
            if (phase == 2 && ws->direct_out != nullptr && ws->direct_flags != nullptr) {
                t.d_out = ws->direct_out;  // (one pass: p0 == 0)
                t.d_flags = ws->direct_flags;
                ws->wrote_direct = true;
                if (ws->xchg != nullptr && ws->xchg->gather_root_plus1 > 0 && ws->xchg->q == qp && ws->xchg->K == K &&
                    qp == q) {
                    t.has_xchg = 1;        // a multi-device set: the recorded step ends with the gather to the root
                    t.xchg = *ws->xchg;
                    ws->xchg_used = true;
                }
            } else if (phase == 0 && d_flags != nullptr) {
                t.d_flags = d_flags;       // the caller's own flags: no copy behind the chain
                flags_direct = true;
                if (ws->xchg != nullptr && ws->xchg->q == qp && ws->xchg->K == K && qp == q) {
                    t.has_xchg = 1;        // ... and the cross-rank exchange + fold at the end of the same kernel
                    t.xchg = *ws->xchg;
                    ws->xchg_used = true;
                }
            }
            t.tiles_per_block = (uint32_t)std::max<int64_t>(1, (cache->block_size / ((int64_t)dims * 4) + 127) / 128);
            GF_CUDA(grouptop_tail_launch(t, stream), GF_ERR_CUDA);
            if (overlap) {
                ws->done_total[par] += (unsigned int)qp;
                ws->op_step++;
                g.d_wait_ctr = nullptr;
                g.d_queries_padded = d_qpad;  // (the launch descriptor is reused by later passes of a larger call)
                g.d_queries_shadow = d_qbf;
                g.d_q_scale = d_qsc;
            }
            ctx->stats.gemm_launches += 1;
            ctx->stats.other_launches += phase != 2 ? 2 : 1;  // (prep,) tail
            ctx->stats.tensor_passes++;
            ctx->stats.rows_scanned += (uint64_t)scanned_rows;
            continue;
        }
        // where this pass's queries sit in the per-query arrays of its tail group
        const int group0 = all ? p0 : (p0 / (kTailBatchPasses * kGemmMaxQ)) * (kTailBatchPasses * kGemmMaxQ);
        const int gq = p0 - group0;
        g.d_thr = d_thr + gq;
        g.d_cand = d_cand + (size_t)gq * cap;
        g.d_cnt = d_cnt + gq;
        g.tile_first = 0;
        g.mode = kGemmTileMax;
        g.tile_stride = stride;
        g.tile_count = (uint32_t)G;
        GF_CUDA(gemm_launch(g, stream), GF_ERR_CUDA);
        GF_CUDA(threshold_launch(d_tmax, G, nq, qp, m, d_thr + gq, d_cnt + gq, stream), GF_ERR_CUDA);
        g.mode = kGemmCollect;
        g.tile_stride = 1;
        g.tile_count = total_gtiles;
        {
            std::pair<cudaEvent_t, cudaEvent_t> ev;
            ctx->prof_begin(stream, &ev);
            cudaError_t le = gemm_launch(g, stream);
            ctx->prof_end(stream, ev, true);
            if (le != cudaSuccess) return cuda_fail(le, "gemm_launch", GF_ERR_CUDA);
        }
        if (all) {
            GF_CUDA(rescore_launch(d_segs, d_seg_slot_base, nseg, dims, qp, d_qpad, (int)cap, d_cand, d_cnt, d_exact,
                                   stream),
                    GF_ERR_CUDA);
            GF_CUDA(finalize_launch(d_exact, d_cand, d_cnt, d_thr, cap, (int)cap, K, qp, scanned_rows, d_qn, d_small,
                                    err_coeff, d_qres, d_small + 2, d_out + (size_t)p0 * K, d_fl + p0, 1, fix, stream),
                    GF_ERR_CUDA);
            ctx->stats.other_launches += 4;
        } else if (p0 + kGemmMaxQ >= q || gq + kGemmMaxQ >= kTailBatchPasses * kGemmMaxQ) {
            // the group's last pass: ONE tail launch for all its queries (selection, two-stage exact re-score,
            // top-K, certificate -- gf_kernels.h CollectTail)
            CollectTail t{};
            t.d_cand = d_cand;
            t.d_cnt = d_cnt;
            t.d_thr = d_thr;
            t.cap = cap;
            t.kp_sel = kp_sel;
            t.K = K;
            t.q = gq + qp;
            t.rows_total = scanned_rows;
            t.d_segs = d_segs;
            t.d_seg_slot_base = d_seg_slot_base;
            t.nseg = nseg;
            t.dims = dims;
            t.rows_per_block = (uint32_t)std::max<int64_t>(1, cache->block_size / ((int64_t)dims * 4));
            t.d_queries = d_qall + (size_t)group0 * dims;
            t.d_qnorm2 = d_qn;
            t.d_maxnorm2_bits = d_small;
            t.err_coeff = err_coeff;
            t.d_qres2 = d_qres;
            t.d_maxres2_bits = d_small + 2;
            t.d_out = d_out + (size_t)group0 * K;
            t.d_flags = d_fl + group0;
            GF_CUDA(collect_tail_launch(t, fix, stream), GF_ERR_CUDA);
            ctx->stats.other_launches += 3;
        } else {
            ctx->stats.other_launches += 2;
        }
        ctx->stats.gemm_launches += 2;
        ctx->stats.tensor_passes++;
        ctx->stats.rows_scanned += (uint64_t)scanned_rows;
    }
    if (phase == 1) {
        ctx->stats.other_launches++;
        return GF_OK;
    }
    // device-side fix-up: one exact scan pass over up to kFixupMaxQ flagged queries (listed by the last
    // finalize launch), merged straight into their rows of d_out; both launches are no-ops when everything
    // was certified
    // A caller that takes the per-query flags re-runs flagged queries itself (Search does, with the exact scan):
    // its chain ends with finalize and saves the two (normally empty) fix-up launches.
    if (device_fixup) {
        rc = topk_scan(ws, stream, d_qall, fixq, dims, 1, K, false, nullptr, d_out, d_map, d_count, true);
        if (rc) return rc;
    }
    if (d_flags && phase == 0 && !flags_direct)
        GF_CUDA(cudaMemcpyAsync(d_flags, d_fl, (size_t)q * sizeof(unsigned int), cudaMemcpyDefault, stream),
                GF_ERR_CUDA);
    return GF_OK;
}

int Set::topk_scan(Workspace *ws, cudaStream_t stream, const float *d_queries, int q, int64_t q_stride,
                   int64_t l_stride, int K, bool per_segment, const unsigned long long *d_upper,
                   unsigned long long *d_out, const int *d_q_map, const unsigned int *d_q_count, bool scatter_out)
{
    const int nseg = (int)h_segs.size();
    const int maxq = scan_max_q(dims);
    if (maxq < 1) return GF_ERR_UNSUPPORTED;
    size_t out_off = 0;
    for (int p0 = 0; p0 < q; p0 += maxq) {
        const int qp = std::min(maxq, q - p0);
        ScanPlan plan;
        GF_CUDA(scan_make_plan(ctx->sms, dims, qp, K, total_tiles, nseg, per_segment, aligned16, &plan), GF_ERR_CUDA);
        // the scratch for the per-CTA lists lives at the tail of the workspace arena
        size_t need = partial_bytes(plan, qp, K);
        if (need > ws->d_cap) return GF_ERR_ALLOCATE_GPU_BUFFER;  // reserved by the caller
        unsigned long long *d_partial = (unsigned long long *)(ws->d_buf + ws->d_cap - align_up(need, 256));
        const unsigned long long *up = nullptr;  // global: [q]; per segment: [pass][segment][qp]
        if (d_upper) up = per_segment ? d_upper + out_off / K : d_upper + p0;
        std::pair<cudaEvent_t, cudaEvent_t> ev;
        ctx->prof_begin(stream, &ev);
        cudaError_t le = scan_launch(plan, d_segs, nseg, total_tiles, d_queries + (size_t)p0 * q_stride, qp, q_stride,
                                     l_stride, K, up, d_partial, stream, d_q_map, d_q_count);
        ctx->prof_end(stream, ev);
        if (le != cudaSuccess) return cuda_fail(le, "scan_launch", GF_ERR_CUDA);
        ctx->stats.scan_launches++;
        ctx->stats.rows_scanned += (uint64_t)scanned_rows;
        GF_CUDA(merge_launch(d_partial, plan.groups, plan.grid_x, qp, K, d_out + out_off, stream, d_q_count,
                             scatter_out ? d_q_map : nullptr),
                GF_ERR_CUDA);
        ctx->stats.merge_launches++;
        out_off += (size_t)plan.groups * qp * K;
    }
    return GF_OK;
}

// worst-case scratch for topk_device with K <= kMaxK
static size_t scratch_bytes(const Set *s, int q, int K, bool per_segment)
{
    const int nseg = (int)s->h_segs.size();
    const int maxq = std::max(1, scan_max_q(s->dims));
    const int qp = std::min(q, maxq);
    size_t lists = per_segment ? (size_t)std::max(nseg, s->ctx->sms * 2 + nseg) : (size_t)s->ctx->sms;
    return align_up(lists * qp * K * sizeof(unsigned long long), 256) + 256;
}

struct KeyList {
    std::vector<unsigned long long> keys;
    bool exhausted = false;
};

// a search call waits ~0.2 ms for its stream: poll instead of paying a blocking wait's wake-up
// The host call's results land in pinned memory: keys (8-byte words, never all ones: the score image of a NaN is 0) and
// flags (0 / 1).  Both areas are filled with all-ones before the launch and polled until no such word is left -- every
// word is written exactly once per call by an aligned store, so the answer is complete the moment the last sentinel is
// gone, a few microseconds before the stream itself reports completion (kernel retirement + the driver's own
// semaphore).  The stream is still consulted now and then: an error ends the wait, and so does its completion.
cudaError_t wait_results(cudaStream_t s, const volatile unsigned long long *keys, size_t nkeys,
                         const volatile unsigned int *flags, size_t nflags)
{
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(2);
    size_t ki = 0, fi = 0;  // everything before these is known to be written
    for (int i = 0;; ++i) {
        while (fi < nflags && flags[fi] != 0xFFFFFFFFu) ++fi;
        if (fi == nflags) {
            while (ki < nkeys && keys[ki] != ~0ull) ++ki;
            if (ki == nkeys) {
                std::atomic_thread_fence(std::memory_order_acquire);
                return cudaSuccess;
            }
        }
        if ((i & 63) == 63) {
            cudaError_t e = cudaStreamQuery(s);
            if (e != cudaErrorNotReady) return e;
            if (std::chrono::steady_clock::now() >= deadline) return cudaStreamSynchronize(s);
        }
    }
}

static cudaError_t wait_stream(cudaStream_t s)
{
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(2);
    for (int i = 0;; ++i) {
        cudaError_t e = cudaStreamQuery(s);
        if (e != cudaErrorNotReady) return e;
        if ((i & 15) == 15 && std::chrono::steady_clock::now() >= deadline) return cudaStreamSynchronize(s);
    }
}

bool Set::direct_applies(int limit, int q, int *kb) const
{
    if (is_multi()) return multi_direct_applies(limit, q, kb);
    if (precision != 4 || scanned_rows <= 0 || limit <= 0 || q <= 0 || q > kGemmMaxQ) return false;
    const int64_t k_total = std::min<int64_t>(limit, scanned_rows);
    if (k_total > kMaxK) return false;
    *kb = (int)k_total;
    return tensor_eligible(q, (int)k_total, false, nullptr, search_path);
}

int Set::direct_keys(Workspace *ws, int kb, int q, const uint8_t *const *parts, const int *part_q, int nparts,
                     unsigned long long **h_keys)
{
    if (is_multi()) return multi_direct_keys(ws, kb, q, parts, part_q, nparts, h_keys);
    const size_t qbytes = (size_t)q * dims * 4;
    const size_t out_keys = (size_t)q * kb;
    const size_t d_need = align_up(qbytes, 256) + align_up(out_keys * 8, 256) + scratch_bytes(this, q, kb, false);
    // (a second keys + flags area: the refused queries' second tier answers into it)
    const size_t h_need = align_up(qbytes, 256) + 2 * (align_up(out_keys * 8, 256) + align_up((size_t)q * 4, 256));
    int rc = ws->reserve(d_need, h_need);
    if (rc) return rc;
    const auto t_a = std::chrono::steady_clock::now();
    float *d_q = (float *)ws->d_buf;
    unsigned long long *d_out = (unsigned long long *)(ws->d_buf + align_up(qbytes, 256));
    float *h_q = (float *)ws->h_buf;
    unsigned long long *h_out = (unsigned long long *)(ws->h_buf + align_up(qbytes, 256));
    unsigned int *h_flags = (unsigned int *)((uint8_t *)h_out + align_up(out_keys * 8, 256));
    unsigned long long *h_out2 = (unsigned long long *)((uint8_t *)h_flags + align_up((size_t)q * 4, 256));
    unsigned int *h_flags2 = (unsigned int *)((uint8_t *)h_out2 + align_up(out_keys * 8, 256));
    size_t off = 0;
    for (int i = 0; i < nparts; ++i) {
        memcpy((uint8_t *)h_q + off, parts[i], (size_t)part_q[i] * dims * 4);
        off += (size_t)part_q[i] * dims * 4;
    }
    ctx->stats.h2d_bytes += qbytes;
    // The kernels read the queries from the pinned staging buffer and the winners come back in ONE copy
    // (keys + flags): three stream operations per call.
    const bool poll_results = out_keys <= 8192;  // (large answers: filling and scanning the area costs more than it saves)
    if (poll_results) {
        memset(h_out, 0xFF, out_keys * 8);  // sentinels (wait_results)
        memset(h_flags, 0xFF, (size_t)q * 4);
    }
    const auto t_b = std::chrono::steady_clock::now();
    rc = topk_device(ws, ws->stream, h_q, q, dims, 1, kb, false, nullptr, h_out, search_path, h_flags, true);
    if (rc) return rc;
    const auto t_c = std::chrono::steady_clock::now();
    GF_CUDA(poll_results ? wait_results(ws->stream, h_out, out_keys, h_flags, (size_t)q) : wait_stream(ws->stream),
            GF_ERR_READ_CUDA_BUFFER);
    const auto t_d = std::chrono::steady_clock::now();
    ctx->stats.host_stage_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t_b - t_a).count();
    ctx->stats.host_launch_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t_c - t_b).count();
    ctx->stats.host_wait_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t_d - t_c).count();
    ctx->stats.host_calls++;
    ctx->stats.d2h_bytes += out_keys * 8 + (size_t)q * 4;
    // Tensor-path queries the certificate refused: ONLY those are re-run, as one compact batch (the refused query
    // vectors are packed at the front of the device staging area), in tiers:
    //   2  more than one scan pass worth of them (> scan_max_q) and a shadow pass refused them: ONE kind::tf32 candidate
    //      pass over the float32 rows for the whole batch -- its error bound is ~4x (bf16) to ~10x (int8) tighter, so
    //      near neighbours the shadow could not separate certify here (0.30 ms per pass at 1 M x 512 whatever the
    //      batch, against 0.35 ms per EIGHT queries of the exact scan);
    //   3  what is still refused (or a small batch): the exact scan, passes of scan_max_q queries.
    std::vector<int> bad;
    for (int i = 0; i < q; ++i)
        if (h_flags[i]) bad.push_back(i);
    auto pack = [&](const std::vector<int> &idx) -> int {
        for (size_t j = 0; j < idx.size(); ++j)
            GF_CUDA(cudaMemcpyAsync(d_q + j * dims, h_q + (size_t)idx[j] * dims, (size_t)dims * 4,
                                    cudaMemcpyHostToDevice, ws->stream),
                    GF_ERR_WRITE_INPUT_BUFFER);
        return GF_OK;
    };
    if (!bad.empty()) {
        ctx->stats.uncertified_queries += bad.size();
        const int sp = search_path.load();
        if ((int)bad.size() > scan_max_q(dims) && sp != 3 && sp != 5 && use_shadow(sp) &&
            tensor_eligible((int)bad.size(), kb, false, nullptr, 5)) {
            const int nb = (int)bad.size();
            rc = pack(bad);
            if (rc) return rc;
            rc = topk_device(ws, ws->stream, d_q, nb, dims, 1, kb, false, nullptr, h_out2, 5, h_flags2, false);
            if (rc) return rc;
            GF_CUDA(wait_stream(ws->stream), GF_ERR_READ_CUDA_BUFFER);
            std::vector<int> still;
            for (int j = 0; j < nb; ++j) {
                if (h_flags2[j]) {
                    still.push_back(bad[j]);
                    continue;
                }
                memcpy(h_out + (size_t)bad[j] * kb, h_out2 + (size_t)j * kb, (size_t)kb * 8);
            }
            ctx->stats.tier2_queries += (uint64_t)(nb - (int)still.size());
            bad.swap(still);
        }
    }
    if (!bad.empty()) {
        const int nb = (int)bad.size();
        // d_q: [q][dims] staging; d_out: [q][kb].  Refused queries in d_q[0 .. nb), results into d_out[0 .. nb)
        rc = pack(bad);
        if (rc) return rc;
        rc = topk_scan(ws, ws->stream, d_q, nb, dims, 1, kb, false, nullptr, d_out);
        if (rc) return rc;
        // (h_q is free again: the winners of the re-run land there before they are scattered)
        unsigned long long *h_fix = (unsigned long long *)h_q;
        if ((size_t)nb * kb * 8 <= align_up(qbytes, 256)) {
            GF_CUDA(cudaMemcpyAsync(h_fix, d_out, (size_t)nb * kb * 8, cudaMemcpyDeviceToHost, ws->stream),
                    GF_ERR_READ_CUDA_BUFFER);
            GF_CUDA(cudaStreamSynchronize(ws->stream), GF_ERR_READ_CUDA_BUFFER);
            for (int j = 0; j < nb; ++j) memcpy(h_out + (size_t)bad[j] * kb, h_fix + (size_t)j * kb, (size_t)kb * 8);
        } else {
            for (int j = 0; j < nb; ++j)
                GF_CUDA(cudaMemcpyAsync(h_out + (size_t)bad[j] * kb, d_out + (size_t)j * kb, (size_t)kb * 8,
                                        cudaMemcpyDeviceToHost, ws->stream),
                        GF_ERR_READ_CUDA_BUFFER);
            GF_CUDA(cudaStreamSynchronize(ws->stream), GF_ERR_READ_CUDA_BUFFER);
        }
    }
    *h_keys = h_out;
    return GF_OK;
}

bool Set::resolve_winners(const unsigned long long *keys, int kb, float threshold, Result *r) const
{
    int n = 0;
    while (n < kb && keys[n] != 0ull) ++n;
    for (int j = 0; j < n; ++j) {  // tombstones among the winners?  (block.go:236-243 selects before it drops them)
        const Block *b;
        int row;
        if (!locate_slot(key_slot(keys[j]), &b, &row) || !b->live[row]) return false;
    }
    for (int j = 0; j < n; ++j) {
        const float sc = key_score(keys[j]);
        if (!(sc >= threshold)) continue;  // set.go:145
        const Block *b;
        int row;
        locate_slot(key_slot(keys[j]), &b, &row);
        r->push(sc, key_slot(keys[j]), b->id_at(row));
    }
    r->end_query();
    return true;
}

// set.go:118-159 + block.go:184-249
int Set::search_now(float threshold, int limit, int q, const uint8_t *queries, Result **out)
{
    if (is_multi()) return multi_search_now(threshold, limit, q, queries, out);
    std::unique_ptr<Result> res(new Result());
    res->begin(q);
    SharedGuard g(lock);
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    ctx->stats.searches++;
    if (q == 0) {
        *out = res.release();
        return GF_OK;
    }
    if (scanned_rows == 0 || limit <= 0) {
        for (int i = 0; i < q; ++i) res->end_query();
        *out = res.release();
        return GF_OK;
    }
    if (precision != 4) return GF_ERR_UNSUPPORTED;
    int rc = ctx->bind();
    if (rc) return rc;
    const int nseg = (int)h_segs.size();
    const int64_t k_total = std::min<int64_t>(limit, scanned_rows);
    const int kb_max = (int)std::min<int64_t>(k_total, kMaxK);

    Workspace *ws = ctx->acquire_ws();
    if (!ws) return GF_ERR_CUDA;
    struct Release {
        Context *c;
        Workspace *w;
        ~Release() { c->release_ws(w); }
    } rel{ctx, ws};

    int kb_direct = 0;
    if (direct_applies(limit, q, &kb_direct)) {
        unsigned long long *h_keys = nullptr;
        rc = direct_keys(ws, kb_direct, q, &queries, &q, 1, &h_keys);
        if (rc) return rc;
        bool quirk = false;
        const auto t_r0 = std::chrono::steady_clock::now();
        for (int i = 0; i < q && !quirk; ++i) quirk = !resolve_winners(h_keys + (size_t)i * kb_direct, kb_direct, threshold, res.get());
        ctx->stats.host_resolve_ns +=
            (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_r0).count();
        if (!quirk) {
            *out = res.release();
            return GF_OK;
        }
        res.reset(new Result());  // a tombstone among the winners: the generic code below takes the per-block path
        res->begin(q);
    }

    const size_t qbytes = (size_t)q * dims * 4;
    const size_t out_keys = (size_t)q * kb_max;
    size_t d_need = align_up(qbytes, 256) + align_up((size_t)q * 8, 256) + align_up(out_keys * 8, 256) +
                    align_up((size_t)q * 4, 256) + scratch_bytes(this, q, kb_max, false);
    size_t h_need = align_up(qbytes, 256) + align_up((size_t)q * 8, 256) + align_up(out_keys * 8, 256) +
                    align_up((size_t)q * 4, 256);
    rc = ws->reserve(d_need, h_need);
    if (rc) return rc;
    float *d_q = (float *)ws->d_buf;
    unsigned long long *d_upper = (unsigned long long *)(ws->d_buf + align_up(qbytes, 256));
    unsigned long long *d_out = (unsigned long long *)((uint8_t *)d_upper + align_up((size_t)q * 8, 256));
    unsigned int *d_flags = (unsigned int *)((uint8_t *)d_out + align_up(out_keys * 8, 256));
    float *h_q = (float *)ws->h_buf;
    unsigned long long *h_upper = (unsigned long long *)(ws->h_buf + align_up(qbytes, 256));
    unsigned long long *h_out = (unsigned long long *)((uint8_t *)h_upper + align_up((size_t)q * 8, 256));
    unsigned int *h_flags = (unsigned int *)((uint8_t *)h_out + align_up(out_keys * 8, 256));

    memcpy(h_q, queries, qbytes);
    GF_CUDA(cudaMemcpyAsync(d_q, h_q, qbytes, cudaMemcpyHostToDevice, ws->stream), GF_ERR_WRITE_INPUT_BUFFER);
    ctx->stats.h2d_bytes += qbytes;

    // ---- set-wide top-k_total, in bands of <= kMaxK keys
    std::vector<KeyList> lists((size_t)q);
    int64_t got = 0;
    bool first_band = true;
    while (got < k_total) {
        const int kb = (int)std::min<int64_t>(kMaxK, k_total - got);
        if (!first_band) {
            for (int i = 0; i < q; ++i) h_upper[i] = lists[i].exhausted ? 0ull : lists[i].keys.back();
            GF_CUDA(cudaMemcpyAsync(d_upper, h_upper, (size_t)q * 8, cudaMemcpyHostToDevice, ws->stream),
                    GF_ERR_WRITE_INPUT_BUFFER);
        }
        rc = topk_device(ws, ws->stream, d_q, q, dims, 1, kb, false, first_band ? nullptr : d_upper, d_out,
                         search_path, d_flags);
        if (rc) return rc;
        GF_CUDA(cudaMemcpyAsync(h_out, d_out, (size_t)q * kb * 8, cudaMemcpyDeviceToHost, ws->stream),
                GF_ERR_READ_CUDA_BUFFER);
        GF_CUDA(cudaMemcpyAsync(h_flags, d_flags, (size_t)q * 4, cudaMemcpyDeviceToHost, ws->stream),
                GF_ERR_READ_CUDA_BUFFER);
        GF_CUDA(cudaStreamSynchronize(ws->stream), GF_ERR_READ_CUDA_BUFFER);
        ctx->stats.d2h_bytes += (size_t)q * kb * 8 + (size_t)q * 4;
        // tensor-path queries that could neither be certified nor fixed on the device: exact scan, one by one
        for (int i = 0; i < q; ++i) {
            if (!h_flags[i]) continue;
            ctx->stats.uncertified_queries++;
            rc = topk_scan(ws, ws->stream, d_q + (size_t)i * dims, 1, dims, 1, kb, false, nullptr,
                           d_out + (size_t)i * kb);
            if (rc) return rc;
            GF_CUDA(cudaMemcpyAsync(h_out + (size_t)i * kb, d_out + (size_t)i * kb, (size_t)kb * 8,
                                    cudaMemcpyDeviceToHost, ws->stream),
                    GF_ERR_READ_CUDA_BUFFER);
            GF_CUDA(cudaStreamSynchronize(ws->stream), GF_ERR_READ_CUDA_BUFFER);
        }
        bool all_done = true;
        for (int i = 0; i < q; ++i) {
            if (lists[i].exhausted) continue;
            int n = 0;
            while (n < kb && h_out[(size_t)i * kb + n] != 0ull) ++n;
            lists[i].keys.insert(lists[i].keys.end(), h_out + (size_t)i * kb, h_out + (size_t)i * kb + n);
            if (n < kb) lists[i].exhausted = true;
            if (!lists[i].exhausted) all_done = false;
        }
        got += kb;
        first_band = false;
        if (all_done) break;
    }

    // ---- tombstones among the winners?  (block.go:236-243 selects before it drops them)
    bool quirk = false;
    for (int i = 0; i < q && !quirk; ++i) {
        for (unsigned long long k : lists[i].keys) {
            int bp, row;
            if (!slot_to_block(key_slot(k), &bp, &row) || !blocks[bp]->live[row]) {
                quirk = true;
                break;
            }
        }
    }

    if (!quirk) {
        for (int i = 0; i < q; ++i) {
            for (unsigned long long k : lists[i].keys) {
                float sc = key_score(k);
                if (!(sc >= threshold)) continue;  // set.go:145
                int bp, row;
                slot_to_block(key_slot(k), &bp, &row);
                res->push(sc, key_slot(k), blocks[bp]->id_at(row));
            }
            res->end_query();
        }
        *out = res.release();
        return GF_OK;
    }

    // ---- exact per-block path: top-limit per block INCLUDING tombstones, drop them, threshold,
    // merge, top-limit (block.go:237-243, set.go:138-157)
    ctx->stats.quirk_fallbacks++;
    const int maxq = std::max(1, scan_max_q(dims));
    int64_t max_rows = 0;
    for (const Segment &s : h_segs) max_rows = std::max<int64_t>(max_rows, s.rows);
    const int64_t kseg_total = std::min<int64_t>(limit, max_rows);
    const int kseg = (int)std::min<int64_t>(kseg_total, kMaxK);
    const size_t seg_keys = (size_t)nseg * q * kseg;
    d_need = align_up(qbytes, 256) + align_up((size_t)nseg * q * 8, 256) + align_up(seg_keys * 8, 256) +
             scratch_bytes(this, q, kseg, true);
    h_need = align_up(qbytes, 256) + align_up((size_t)nseg * q * 8, 256) + align_up(seg_keys * 8, 256);
    rc = ws->reserve(d_need, h_need);
    if (rc) return rc;
    d_q = (float *)ws->d_buf;
    d_upper = (unsigned long long *)(ws->d_buf + align_up(qbytes, 256));
    d_out = (unsigned long long *)((uint8_t *)d_upper + align_up((size_t)nseg * q * 8, 256));
    h_q = (float *)ws->h_buf;
    h_upper = (unsigned long long *)(ws->h_buf + align_up(qbytes, 256));
    h_out = (unsigned long long *)((uint8_t *)h_upper + align_up((size_t)nseg * q * 8, 256));
    memcpy(h_q, queries, qbytes);
    GF_CUDA(cudaMemcpyAsync(d_q, h_q, qbytes, cudaMemcpyHostToDevice, ws->stream), GF_ERR_WRITE_INPUT_BUFFER);

    // lists per (query, segment); layout of every band on the device: [pass][segment][qp][kb]
    std::vector<KeyList> seg_lists((size_t)q * nseg);
    got = 0;
    first_band = true;
    while (got < kseg_total) {
        const int kb = (int)std::min<int64_t>(kMaxK, kseg_total - got);
        if (!first_band) {
            size_t off = 0;
            for (int p0 = 0; p0 < q; p0 += maxq) {
                int qp = std::min(maxq, q - p0);
                for (int sgi = 0; sgi < nseg; ++sgi)
                    for (int j = 0; j < qp; ++j) {
                        KeyList &kl = seg_lists[(size_t)(p0 + j) * nseg + sgi];
                        h_upper[off++] = kl.exhausted ? 0ull : kl.keys.back();
                    }
            }
            GF_CUDA(cudaMemcpyAsync(d_upper, h_upper, (size_t)nseg * q * 8, cudaMemcpyHostToDevice, ws->stream),
                    GF_ERR_WRITE_INPUT_BUFFER);
        }
        rc = topk_device(ws, ws->stream, d_q, q, dims, 1, kb, true, first_band ? nullptr : d_upper, d_out);
        if (rc) return rc;
        GF_CUDA(cudaMemcpyAsync(h_out, d_out, (size_t)nseg * q * kb * 8, cudaMemcpyDeviceToHost, ws->stream),
                GF_ERR_READ_CUDA_BUFFER);
        GF_CUDA(cudaStreamSynchronize(ws->stream), GF_ERR_READ_CUDA_BUFFER);
        ctx->stats.d2h_bytes += (size_t)nseg * q * kb * 8;
        bool all_done = true;
        size_t off = 0;
        for (int p0 = 0; p0 < q; p0 += maxq) {
            int qp = std::min(maxq, q - p0);
            for (int sgi = 0; sgi < nseg; ++sgi)
                for (int j = 0; j < qp; ++j) {
                    KeyList &kl = seg_lists[(size_t)(p0 + j) * nseg + sgi];
                    const unsigned long long *src = h_out + off;
                    off += kb;
                    if (kl.exhausted) continue;
                    int n = 0;
                    while (n < kb && src[n] != 0ull) ++n;
                    kl.keys.insert(kl.keys.end(), src, src + n);
                    if (n < kb) kl.exhausted = true;
                    if (!kl.exhausted) all_done = false;
                }
        }
        got += kb;
        first_band = false;
        if (all_done) break;
    }
    for (int i = 0; i < q; ++i) {
        std::vector<unsigned long long> merged;
        for (int sgi = 0; sgi < nseg; ++sgi) {
            const KeyList &kl = seg_lists[(size_t)i * nseg + sgi];
            size_t take = std::min<size_t>(kl.keys.size(), (size_t)limit);  // MaxNFloat32(vec, limit)
            for (size_t j = 0; j < take; ++j) {
                unsigned long long k = kl.keys[j];
                int bp, row;
                if (!slot_to_block(key_slot(k), &bp, &row) || !blocks[bp]->live[row]) continue;  // block.go:239
                if (!(key_score(k) >= threshold)) continue;                                        // set.go:145
                merged.push_back(k);
            }
        }
        std::sort(merged.begin(), merged.end(), std::greater<unsigned long long>());  // set.go:155
        if (merged.size() > (size_t)limit) merged.resize((size_t)limit);
        for (unsigned long long k : merged) {
            int bp, row;
            slot_to_block(key_slot(k), &bp, &row);
            res->push(key_score(k), key_slot(k), blocks[bp]->id_at(row));
        }
        res->end_query();
    }
    *out = res.release();
    return GF_OK;
}

int Set::search(float threshold, int limit, int q, const uint8_t *queries, Result **out)
{
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    if (q > batch) return GF_ERR_OUT_OF_BATCH;  // set.go:119-122
    if (coalesce_wait_us > 0 && q > 0 && limit > 0) return search_coalesced(threshold, limit, q, queries, out);
    return search_now(threshold, limit, q, queries, out);
}

// The analogue of the reference's SearchQueue (set.go:36, cache.go:61: every Search of a set feeds
// one channel).  Concurrent callers are folded into ONE pass over the rows: the first caller to
// arrive leads, waits at most coalesce_wait_us for company, runs the concatenated batch (same limit)
// with no threshold, and hands every follower its slice with its own threshold applied
// (set.go:145 filters after selection, so filtering a shared top-limit is exact).  While a batch
// runs, new callers queue up and form the next batch.
int Set::search_coalesced(float threshold, int limit, int q, const uint8_t *queries, Result **out)
{
    PendingSearch me;
    me.queries = queries;
    me.q = q;
    me.limit = limit;
    me.threshold = threshold;
    const size_t qbytes = (size_t)dims * precision;
    static thread_local const Set *tl_set = nullptr;
    static thread_local uint64_t tl_batch = 0;
    static const int hw = (int)std::max(2u, std::thread::hardware_concurrency());
    struct Inside {
        std::atomic<int> &c;
        explicit Inside(std::atomic<int> &c_) : c(c_) { ++c; }
        ~Inside() { --c; }
    } inside(co_inside);
    auto spin_until = [](const std::function<bool()> &ready, int us) {
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(us);
        for (int i = 0;; ++i) {
            if (ready()) return true;
            if ((i & 63) == 63 && std::chrono::steady_clock::now() >= deadline) return false;
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
    };
    std::unique_lock<std::mutex> lk(co_mu);
    me.returnee = (tl_set == this && tl_batch == co_batch_id && co_batch_id != 0);
    co_queue.push_back(&me);
    co_queued.store((int)co_queue.size(), std::memory_order_release);
    co_cv.notify_all();  // a parked leader re-counts its company
    bool my_retry = false;  // fast hand-off only: a tombstone among MY winners -> my own uncoalesced search
    while (!me.done.load(std::memory_order_acquire)) {
        if (co_leader_active.load()) {  // someone else is leading: wait to be served (or to take over)
            // a batch takes a few hundred microseconds: spin through it before paying a futex sleep + wake-up
            if (co_inside.load() <= hw - 4) {  // more waiters than spare cores: spinning would starve the leader
                lk.unlock();
                spin_until([&] { return me.done.load(std::memory_order_acquire) || !co_leader_active.load(); }, 600);
                if (me.done.load(std::memory_order_acquire)) break;  // served: nothing left to do under the lock
                lk.lock();
            }
            if (!me.done.load(std::memory_order_acquire) && co_leader_active.load()) co_cv.wait(lk);
            continue;
        }
        co_leader_active = true;
        // How many callers to expect: everybody served by the previous batch is about to come back,
        // plus whoever queued up meanwhile and was not in it.  Wait for them, but never longer than
        // coalesce_wait_us: two half-size batches alternating cost twice the HBM traffic of one.
        {
            int fresh = 0;
            for (PendingSearch *p : co_queue) fresh += p->returnee ? 0 : 1;
            const int target = co_last_group + fresh;
            if ((int)co_queue.size() < target) {
                lk.unlock();
                spin_until([&] { return co_queued.load(std::memory_order_acquire) >= target; }, coalesce_wait_us);
                lk.lock();
            }
        }
        // take every queued request with my limit, up to 256 queries
        std::vector<PendingSearch *> group;
        int total = 0;
        for (auto it = co_queue.begin(); it != co_queue.end();) {
            PendingSearch *p = *it;
            if (p->limit == me.limit && (p == &me || total + p->q <= kGemmMaxQ)) {
                group.push_back(p);
                total += p->q;
                it = co_queue.erase(it);
            } else {
                ++it;
            }
        }
        co_queued.store((int)co_queue.size(), std::memory_order_release);
        lk.unlock();
        if (group.size() > 1) ctx->stats.coalesced_batches++;

        // ---- fast hand-off: one tensor-path pass, raw keys published, every caller resolves its own
        bool fast = false;
        {
            SharedGuard g(lock);
            int kb = 0;
            if (!destroyed && direct_applies(limit, total, &kb) && ctx->bind() == GF_OK) {
                Workspace *ws = ctx->acquire_ws();
                if (ws) {
                    fast = true;
                    ctx->stats.searches++;
                    std::vector<const uint8_t *> parts(group.size());
                    std::vector<int> part_q(group.size());
                    for (size_t i = 0; i < group.size(); ++i) {
                        parts[i] = group[i]->queries;
                        part_q[i] = group[i]->q;
                    }
                    unsigned long long *h_keys = nullptr;
                    const int rc = direct_keys(ws, kb, total, parts.data(), part_q.data(), (int)group.size(), &h_keys);
                    std::atomic<int> resolvers{(int)group.size() - 1};
                    lk.lock();
                    const uint64_t bid = ++co_batch_id;
                    co_last_group = (int)group.size();
                    int q0 = 0;
                    const unsigned long long *mine = nullptr;
                    for (PendingSearch *p : group) {
                        p->rc = rc;
                        p->batch_id = bid;
                        if (p == &me) {
                            mine = rc == GF_OK ? h_keys + (size_t)q0 * kb : nullptr;
                        } else {
                            p->keys = rc == GF_OK ? h_keys + (size_t)q0 * kb : nullptr;
                            p->kb = kb;
                            p->resolvers = &resolvers;
                        }
                        q0 += p->q;
                    }
                    for (PendingSearch *p : group)
                        if (p != &me) p->done.store(true, std::memory_order_release);  // p may be gone right after its resolve
                    co_leader_active = false;
                    co_cv.notify_all();
                    lk.unlock();
                    if (mine) {
                        me.result = new Result();
                        me.result->begin(me.q);
                        for (int i = 0; i < me.q && !my_retry; ++i)
                            my_retry = !resolve_winners(mine + (size_t)i * kb, kb, me.threshold, me.result);
                    }
                    // the staging buffer and the read lock stay until every caller has read its keys
                    for (int i = 0; resolvers.load(std::memory_order_acquire) > 0; ++i)
                        if ((i & 1023) == 1023) std::this_thread::yield();
                    ctx->release_ws(ws);
                    me.done.store(true, std::memory_order_release);
                }
            }
        }
        if (fast) {
            break;  // me.done is set; the lock is not held
        }

        // ---- generic path (scan kernel, bands, empty sets ...): one search_now, results split by the leader
        std::vector<uint8_t> flat((size_t)total * qbytes);
        size_t off = 0;
        for (PendingSearch *p : group) {
            memcpy(flat.data() + off, p->queries, (size_t)p->q * qbytes);
            off += (size_t)p->q * qbytes;
        }
        Result *all = nullptr;
        int rc = search_now(-INFINITY, limit, total, flat.data(), &all);
        int q0 = 0;
        for (PendingSearch *p : group) {
            p->rc = rc;
            if (rc == GF_OK) {
                Result *r = new Result();
                r->begin(p->q);
                for (int i = 0; i < p->q; ++i) {
                    for (int64_t k = all->offsets[q0 + i]; k < all->offsets[q0 + i + 1]; ++k) {
                        if (!(all->scores[k] >= p->threshold)) continue;  // set.go:145
                        r->push(all->scores[k], all->slots[k],
                                all->id_bytes.substr(all->id_off[k], all->id_off[k + 1] - all->id_off[k]));
                    }
                    r->end_query();
                }
                p->result = r;
            }
            q0 += p->q;
        }
        delete all;
        lk.lock();
        const uint64_t bid = ++co_batch_id;
        co_last_group = (int)group.size();
        for (PendingSearch *p : group) p->batch_id = bid;
        for (PendingSearch *p : group)
            if (p != &me) p->done.store(true, std::memory_order_release);  // p may be gone right after this
        me.done.store(true, std::memory_order_release);
        co_leader_active = false;
        co_cv.notify_all();
    }
    if (lk.owns_lock()) lk.unlock();
    tl_set = this;
    tl_batch = me.batch_id;
    if (me.rc) {
        if (me.resolvers) me.resolvers->fetch_sub(1, std::memory_order_release);
        return me.rc;
    }
    if (me.keys != nullptr) {  // fast hand-off: build my own result; the leader holds the set's read lock meanwhile
        Result *r = new Result();
        r->begin(me.q);
        for (int i = 0; i < me.q && !my_retry; ++i)
            my_retry = !resolve_winners(me.keys + (size_t)i * me.kb, me.kb, me.threshold, r);
        me.resolvers->fetch_sub(1, std::memory_order_release);  // last access to anything the leader owns
        me.result = r;
    }
    if (my_retry) {  // block.go:236-243's tombstone quirk: rare, and exact only per block
        delete me.result;
        return search_now(threshold, limit, q, queries, out);
    }
    *out = me.result;
    return GF_OK;
}

int Set::search_device(int limit, int q, const float *d_queries, unsigned long long *d_out, int path,
                       cudaStream_t stream, unsigned int *d_flags, const ExchangeArgs *xchg, bool *xchg_used)
{
    if (xchg_used) *xchg_used = false;
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    if (is_multi()) {
        set_error_detail("device-resident searches address one device: use gf_set_search on a multi-device set");
        return GF_ERR_UNSUPPORTED;
    }
    if (q > batch) return GF_ERR_OUT_OF_BATCH;
    if (q <= 0 || limit <= 0) return GF_ERR_INVALID_ARGUMENT;
    if (limit > kMaxK) return GF_ERR_UNSUPPORTED;
    if (precision != 4) return GF_ERR_UNSUPPORTED;
    SharedGuard g(lock);
    int rc = ctx->bind();
    if (rc) return rc;
    ctx->stats.searches++;
    // one dedicated workspace per set for device-resident searches (calls must be stream ordered)
    Workspace *ws;
    {
        std::lock_guard<std::mutex> lg(co_mu);
        if (!dev_ws) dev_ws = ctx->acquire_ws();
        ws = dev_ws;
    }
    if (!ws) return GF_ERR_CUDA;
    cudaStream_t st = stream;  // NULL is the CUDA legacy default stream, as everywhere in CUDA
    if (scanned_rows == 0) {
        GF_CUDA(cudaMemsetAsync(d_out, 0, (size_t)q * limit * 8, st), GF_ERR_CUDA);
        if (d_flags) GF_CUDA(cudaMemsetAsync(d_flags, 0, (size_t)q * 4, st), GF_ERR_CUDA);
        return GF_OK;
    }
    size_t need = scratch_bytes(this, q, limit, false);
    if (need > ws->d_cap) {
        GF_CUDA(cudaStreamSynchronize(st), GF_ERR_CUDA);
        rc = ws->reserve(need, 0);
        if (rc) return rc;
    }
    ws->xchg = d_flags != nullptr ? xchg : nullptr;
    ws->xchg_used = false;
    rc = topk_device(ws, st, d_queries, q, dims, 1, limit, false, nullptr, d_out, path, d_flags);
    if (xchg_used) *xchg_used = ws->xchg_used;
    ws->xchg = nullptr;
    return rc;
}

int Set::resolve_keys(float threshold, int limit, int q, const unsigned long long *h_keys, Result **out)
{
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    std::unique_ptr<Result> res(new Result());
    res->begin(q);
    SharedGuard g(lock);
    for (int i = 0; i < q; ++i) {
        for (int j = 0; j < limit; ++j) {
            unsigned long long k = h_keys[(size_t)i * limit + j];
            if (k == 0ull) break;
            float sc = key_score(k);
            if (!(sc >= threshold)) continue;
            const Block *b;
            int row;
            if (locate_slot(key_slot(k), &b, &row)) {
                if (!b->live[row]) continue;
                res->push(sc, key_slot(k), b->id_at(row));
            } else {
                res->push(sc, key_slot(k), std::string());  // another shard owns this slot
            }
        }
        res->end_query();
    }
    *out = res.release();
    return GF_OK;
}

// block.go:184-249 on one bare block; `input` holds the dims-major transposed queries
int gf::block_search(Block *b, const Buffer *input, int batch, int limit, Result **out)
{
    std::unique_ptr<Result> res(new Result());
    if (!b->is_owned()) return GF_ERR_INVALID_ARGUMENT;
    if (b->next_index == 0) {  // block.go:186-189: `return` with ret == nil
        res->begin(0);
        *out = res.release();
        return GF_OK;
    }
    if (b->precision != 4) return GF_ERR_UNSUPPORTED;
    if (batch <= 0 || (uint64_t)batch * b->dims * 4 > input->size) return GF_ERR_INVALID_ARGUMENT;
    res->begin(batch);
    if (limit <= 0) {
        for (int i = 0; i < batch; ++i) res->end_query();
        *out = res.release();
        return GF_OK;
    }
    Context *ctx = b->cache->ctx;
    int rc = ctx->bind();
    if (rc) return rc;
    std::lock_guard<std::mutex> g(b->mu);
    ctx->stats.searches++;

    // a one-segment, un-sharded view of the block
    Set view;
    view.cache = b->cache;
    view.ctx = ctx;
    view.dims = b->dims;
    view.precision = 4;
    view.batch = batch;
    view.cap = b->capacity();
    view.blocks.push_back(b);
    Workspace *ws = ctx->acquire_ws();
    if (!ws) return GF_ERR_CUDA;
    struct Release {
        Context *c;
        Workspace *w;
        ~Release() { c->release_ws(w); }
    } rel{ctx, ws};
    {
        // build the segment table in the workspace instead of Set::rebuild_segments' own buffer
        int tr = scan_tile_rows(b->dims);
        view.aligned16 = (b->dims % 4 == 0) && (((uintptr_t)b->dev) % 16 == 0);
        view.tile_rows = (view.aligned16 && tr > 0) ? tr : 64;
        Segment s;
        s.base = (const float *)b->dev;
        s.rows = (uint32_t)b->next_index;
        s.slot_base = 0;
        s.tile_start = 0;
        s.pad = 0;
        view.h_segs.push_back(s);
        view.total_tiles = (s.rows + view.tile_rows - 1) / view.tile_rows;
        view.scanned_rows = b->next_index;
    }
    const int64_t k_total = std::min<int64_t>(limit, b->next_index);
    const int kb_max = (int)std::min<int64_t>(k_total, kMaxK);
    const size_t out_keys = (size_t)batch * kb_max;
    size_t d_need = 256 + align_up((size_t)batch * 8, 256) + align_up(out_keys * 8, 256) +
                    scratch_bytes(&view, batch, kb_max, false);
    size_t h_need = 256 + align_up((size_t)batch * 8, 256) + align_up(out_keys * 8, 256);
    rc = ws->reserve(d_need, h_need);
    if (rc) return rc;
    Segment *d_seg = (Segment *)ws->d_buf;
    unsigned long long *d_upper = (unsigned long long *)(ws->d_buf + 256);
    unsigned long long *d_out = (unsigned long long *)((uint8_t *)d_upper + align_up((size_t)batch * 8, 256));
    Segment *h_seg = (Segment *)ws->h_buf;
    unsigned long long *h_upper = (unsigned long long *)(ws->h_buf + 256);
    unsigned long long *h_out = (unsigned long long *)((uint8_t *)h_upper + align_up((size_t)batch * 8, 256));
    *h_seg = view.h_segs[0];
    GF_CUDA(cudaMemcpyAsync(d_seg, h_seg, sizeof(Segment), cudaMemcpyHostToDevice, ws->stream), GF_ERR_CUDA);
    view.d_segs = d_seg;

    std::vector<KeyList> lists((size_t)batch);
    int64_t got = 0;
    bool first_band = true;
    while (got < k_total) {
        const int kb = (int)std::min<int64_t>(kMaxK, k_total - got);
        if (!first_band) {
            for (int i = 0; i < batch; ++i) h_upper[i] = lists[i].exhausted ? 0ull : lists[i].keys.back();
            GF_CUDA(cudaMemcpyAsync(d_upper, h_upper, (size_t)batch * 8, cudaMemcpyHostToDevice, ws->stream),
                    GF_ERR_CUDA);
        }
        // dims-major queries: element l of query i at input[l*batch + i] (util.go:33-36)
        rc = view.topk_device(ws, ws->stream, (const float *)input->ptr, batch, 1, batch, kb, false,
                              first_band ? nullptr : d_upper, d_out);
        if (rc) break;
        cudaError_t e = cudaMemcpyAsync(h_out, d_out, (size_t)batch * kb * 8, cudaMemcpyDeviceToHost, ws->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ws->stream);
        if (e != cudaSuccess) {
            rc = cuda_fail(e, "block search readback", GF_ERR_READ_CUDA_BUFFER);
            break;
        }
        bool all_done = true;
        for (int i = 0; i < batch; ++i) {
            if (lists[i].exhausted) continue;
            int n = 0;
            while (n < kb && h_out[(size_t)i * kb + n] != 0ull) ++n;
            lists[i].keys.insert(lists[i].keys.end(), h_out + (size_t)i * kb, h_out + (size_t)i * kb + n);
            if (n < kb) lists[i].exhausted = true;
            if (!lists[i].exhausted) all_done = false;
        }
        got += kb;
        first_band = false;
        if (all_done) break;
    }
    view.d_segs = nullptr;  // owned by the workspace
    view.blocks.clear();
    if (rc) return rc;
    for (int i = 0; i < batch; ++i) {
        for (unsigned long long k : lists[i].keys) {
            int row = (int)key_slot(k);
            if (!b->live[row]) continue;  // block.go:239
            res->push(key_score(k), (uint64_t)row, b->id_at(row));
        }
        res->end_query();
    }
    *out = res.release();
    return GF_OK;
}
