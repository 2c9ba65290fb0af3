// gf_multi.cu -- one process, several devices, behind the same C ABI (gf_context_create_multi).
//
// The reference builds its cache on one context (cache.go:21-43, `devices[0]` in search_test.go:40).  Here a context
// may span G devices: block i of a cache lives on device i % G (cache.go:104-115 hands blocks out in AllBlocks order,
// so a growing set is striped evenly), a set is G shards (shard d = rank d of G, slots stay global:
// slot = global block * capacity + row), and Set.Search (set.go:118-159) runs every shard's scan from ONE host
// thread: each device replays its recorded chain (query preparation -> tcgen05 candidate pass -> tail), whose last
// kernel stores the shard's K winners per query into the ROOT device's inbox over NVLink peer memory; the root's tail
// waits for the G contributions, folds them and writes the merged keys into pinned host memory.  No collective
// library, no second process, one wait on the host; ids are resolved by the shard that owns each winning slot.
//
// What it restates: FeatureSet.Add's fill order (set.go:97-114) over the striped blocks, FeatureSet.Delete's
// first-block-in-order rule (set.go:163-193), the per-block tombstone rule (block.go:236-243) through the shards'
// own Search when a tombstone reaches the winners.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <functional>
#include <thread>

#include "gf_device.cuh"
#include "gf_runtime.h"

namespace gf {

MultiExt::~MultiExt()
{
    for (size_t d = 0; d < ctxs.size(); ++d) {
        cudaSetDevice(ctxs[d]->device);
        if (d < ws.size() && ws[d]) {
            cudaStreamSynchronize(ws[d]->stream);
            ctxs[d]->release_ws(ws[d]);
        }
        if (d < inbox.size() && inbox[d]) cudaFree(inbox[d]);
        if (d < d_peers.size() && d_peers[d]) cudaFree(d_peers[d]);
    }
    if (h_pin) cudaFreeHost(h_pin);
}

}  // namespace gf

using namespace gf;

namespace {

struct ExclusiveGuard {
    RWLock &l;
    explicit ExclusiveGuard(RWLock &l_) : l(l_) { l.lock(); }
    ~ExclusiveGuard() { l.unlock(); }
};
struct SharedGuard {
    RWLock &l;
    explicit SharedGuard(RWLock &l_) : l(l_) { l.lock_shared(); }
    ~SharedGuard() { l.unlock_shared(); }
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace
// (gf_runtime.cu) poll the pinned result words instead of the stream
cudaError_t wait_results(cudaStream_t s, const volatile unsigned long long *keys, size_t nkeys,
                         const volatile unsigned int *flags, size_t nflags);
namespace {
cudaError_t wait_stream(cudaStream_t s)
{
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(2);
    for (int i = 0;; ++i) {
        cudaError_t e = cudaStreamQuery(s);
        if (e != cudaErrorNotReady) return e;
        if ((i & 15) == 15 && std::chrono::steady_clock::now() >= deadline) return cudaStreamSynchronize(s);
    }
}

// one (global block order) chunk of an Add: rows [off, off + len) of the request go to shard `shard`
struct Chunk {
    int shard;
    int64_t off, len;
};

// set.go:97-114 over the striped blocks: existing blocks in global order take min(margin, remain), then new
// blocks (global block B, B + 1, ... -> shard B % G) take the rest.  room(block) = rows the block can take.
int plan_fill(const Set *P, int64_t n, const std::function<int64_t(const Block *)> &room, std::vector<Chunk> &out,
              std::vector<int> &new_blocks)
{
    const int G = (int)P->shards.size();
    size_t B = 0;
    for (const Set *sh : P->shards) B += sh->blocks.size();
    new_blocks.assign((size_t)G, 0);
    int64_t offset = 0, remain = n;
    for (size_t g = 0; g < B && remain > 0; ++g) {
        const Set *sh = P->shards[g % G];
        const size_t pos = g / G;
        if (pos >= sh->blocks.size()) return GF_ERR_INVALID_SET_STATE;  // (the stripes are kept even by construction)
        const int64_t m = std::min<int64_t>(room(sh->blocks[pos]), remain);
        if (m > 0) {
            out.push_back({(int)(g % G), offset, m});
            offset += m;
            remain -= m;
        }
    }
    if (remain > 0 && P->cap <= 0) return GF_ERR_NOT_ENOUGH_BLOCKS;
    while (remain > 0) {
        const int64_t m = std::min<int64_t>(P->cap, remain);
        out.push_back({(int)(B % G), offset, m});
        new_blocks[B % G]++;
        ++B;
        offset += m;
        remain -= m;
    }
    // every shard's cache must be able to hand out what its stripe needs (cache.go:111-113), before anything is written
    for (int d = 0; d < G; ++d) {
        if (new_blocks[d] == 0) continue;
        Cache *cc = P->shards[d]->cache;
        std::lock_guard<std::mutex> cg(cc->mu);
        int free_blocks = 0;
        for (auto &ub : cc->all_blocks)
            if (!ub->is_owned()) ++free_blocks;
        if (free_blocks < new_blocks[d]) return GF_ERR_NOT_ENOUGH_BLOCKS;
    }
    return GF_OK;
}

}  // namespace

void Set::multi_refresh()
{
    int64_t rows = 0;
    for (const Set *sh : shards) rows += sh->scanned_rows;
    scanned_rows = rows;
    version = next_set_version();
}

bool Set::multi_find(const std::string &id, int *shard, int *pos, int *row) const
{
    const int G = (int)shards.size();
    long long best = -1;
    for (int d = 0; d < G; ++d) {
        int p, r;
        if (!shards[d]->find(id, &p, &r)) continue;
        const long long g = (long long)p * G + d;
        if (best < 0 || g < best) {
            best = g;
            *shard = d;
            *pos = p;
            *row = r;
        }
    }
    return best >= 0;
}

// ------------------------------------------------------------------ mutations (parent's exclusive lock)
int Set::multi_add(int64_t n, const uint8_t *values, const std::vector<std::string> &ids_)
{
    ExclusiveGuard g(lock);
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    std::vector<Chunk> chunks;
    std::vector<int> new_blocks;
    int rc = plan_fill(this, n, [](const Block *b) { return (int64_t)b->margin(); }, chunks, new_blocks);
    if (rc) return rc;
    const size_t rowb = (size_t)dims * precision;
    const int G = (int)shards.size();
    // one Add per shard with its chunks concatenated in order: the shard's own fill loop (set.go:97-114) then puts
    // every chunk into the block the plan meant (a chunk is either a block's whole margin or the request's tail)
    std::vector<int> rcs((size_t)G, GF_OK);
    std::vector<std::thread> th;
    for (int d = 0; d < G; ++d) {
        int64_t total = 0;
        int parts = 0;
        for (const Chunk &c : chunks)
            if (c.shard == d) {
                total += c.len;
                ++parts;
            }
        if (total == 0) continue;
        th.emplace_back([&, d, total, parts]() {
            std::vector<std::string> sub;
            sub.reserve((size_t)total);
            std::vector<uint8_t> packed;
            const uint8_t *src = nullptr;
            if (parts > 1) packed.reserve((size_t)total * rowb);
            for (const Chunk &c : chunks) {
                if (c.shard != d) continue;
                sub.insert(sub.end(), ids_.begin() + c.off, ids_.begin() + c.off + c.len);
                if (parts == 1)
                    src = values + (size_t)c.off * rowb;
                else
                    packed.insert(packed.end(), values + (size_t)c.off * rowb, values + (size_t)(c.off + c.len) * rowb);
            }
            rcs[d] = shards[d]->add(total, parts == 1 ? src : packed.data(), sub);
        });
    }
    for (auto &t : th) t.join();
    multi_refresh();
    for (int r : rcs)
        if (r) return r;
    return GF_OK;
}

int Set::multi_add_synthetic(int64_t n, const gf::SynthSpec &seed, int64_t first_ordinal, int normalize, const std::string &prefix)
{
    ExclusiveGuard g(lock);
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    std::vector<Chunk> chunks;
    std::vector<int> new_blocks;
    // synthetic rows only append (Set::add_synthetic): a block's room is its tail
    int rc = plan_fill(this, n, [](const Block *b) { return (int64_t)(b->capacity() - b->next_index); }, chunks,
                       new_blocks);
    if (rc) return rc;
    const int G = (int)shards.size();
    std::vector<int> rcs((size_t)G, GF_OK);
    std::vector<std::thread> th;
    for (int d = 0; d < G; ++d)
        th.emplace_back([&, d]() {
            for (const Chunk &c : chunks) {
                if (c.shard != d || rcs[d]) continue;
                rcs[d] = shards[d]->add_synthetic(c.len, seed, first_ordinal + c.off, normalize, prefix);
            }
        });
    for (auto &t : th) t.join();
    multi_refresh();
    for (int r : rcs)
        if (r) return r;
    return GF_OK;
}

int Set::multi_remove(const std::vector<std::string> &ids_, std::vector<uint8_t> &found)
{
    ExclusiveGuard g(lock);
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    const int G = (int)shards.size();
    std::vector<std::vector<std::string>> sub((size_t)G);
    std::vector<std::vector<size_t>> who((size_t)G);
    std::unordered_map<std::string, size_t> first;
    for (size_t i = 0; i < ids_.size(); ++i) first.emplace(ids_[i], i);  // a repeated id is looked up once (block.go:136-139)
    for (size_t i = 0; i < ids_.size(); ++i) {
        if (first[ids_[i]] != i) continue;
        int d, p, r;
        if (!multi_find(ids_[i], &d, &p, &r)) continue;
        sub[d].push_back(ids_[i]);
        who[d].push_back(i);
    }
    for (int d = 0; d < G; ++d) {
        if (sub[d].empty()) continue;
        std::vector<uint8_t> f;
        int rc = shards[d]->remove(sub[d], f);
        if (rc) return rc;
        for (size_t k = 0; k < f.size(); ++k) found[who[d][k]] = f[k];
    }
    return GF_OK;
}

int Set::multi_update(int64_t n, const uint8_t *values, const std::vector<std::string> &ids_, std::vector<uint8_t> &found)
{
    ExclusiveGuard g(lock);
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    const size_t rowb = (size_t)dims * precision;
    for (int64_t i = 0; i < n; ++i) {
        int d, p, r;
        if (!multi_find(ids_[i], &d, &p, &r)) continue;
        std::vector<uint8_t> f;
        int rc = shards[d]->update(1, values + (size_t)i * rowb, {ids_[i]}, f);
        if (rc) return rc;
        found[i] = f.empty() ? 0 : f[0];
    }
    return GF_OK;
}

int Set::multi_read(const std::vector<std::string> &ids_, uint8_t *out, std::vector<uint8_t> &found)
{
    SharedGuard g(lock);
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    const size_t rowb = (size_t)dims * precision;
    for (size_t i = 0; i < ids_.size(); ++i) {
        memset(out + i * rowb, 0, rowb);
        int d, p, r;
        if (!multi_find(ids_[i], &d, &p, &r)) continue;
        std::vector<uint8_t> f;
        int rc = shards[d]->read({ids_[i]}, out + i * rowb, f);
        if (rc) return rc;
        found[i] = f.empty() ? 0 : f[0];
    }
    return GF_OK;
}

int Set::multi_compact(int64_t *reclaimed)
{
    ExclusiveGuard g(lock);
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    int64_t total = 0;
    for (Set *sh : shards) {
        int64_t r = 0;
        int rc = sh->compact(&r);
        if (rc) return rc;
        total += r;
    }
    if (reclaimed) *reclaimed = total;
    multi_refresh();
    return GF_OK;
}

int Set::multi_export(uint8_t *out_values, uint64_t values_cap, char *out_ids, uint64_t ids_cap, uint64_t *out_off,
                      int64_t *out_rows, uint64_t *out_id_bytes)
{
    SharedGuard g(lock);
    const int G = (int)shards.size();
    const size_t rowb = (size_t)dims * precision;
    std::vector<int64_t> rows((size_t)G, 0);
    std::vector<uint64_t> idb((size_t)G, 0);
    int64_t rows_total = 0;
    uint64_t idb_total = 0;
    for (int d = 0; d < G; ++d) {
        int rc = shards[d]->export_rows(nullptr, 0, nullptr, 0, nullptr, &rows[d], &idb[d]);
        if (rc) return rc;
        rows_total += rows[d];
        idb_total += idb[d];
    }
    if (out_rows) *out_rows = rows_total;
    if (out_id_bytes) *out_id_bytes = idb_total;
    if (!out_values && !out_ids && !out_off) return GF_OK;  // size query
    if (!out_values || !out_ids || !out_off) return GF_ERR_INVALID_ARGUMENT;
    if (values_cap < (uint64_t)rows_total * rowb || ids_cap < idb_total) return GF_ERR_BUFFER_WRITE_OUT_OF_RANGE;
    // every shard exports its own blocks in order; the global slot order interleaves them block by block
    std::vector<std::vector<uint8_t>> vals((size_t)G);
    std::vector<std::string> idsb((size_t)G);
    std::vector<std::vector<uint64_t>> offs((size_t)G);
    for (int d = 0; d < G; ++d) {
        vals[d].resize((size_t)rows[d] * rowb + 1);
        idsb[d].resize((size_t)idb[d] + 1);
        offs[d].assign((size_t)rows[d] + 1, 0);
        int64_t r = 0;
        uint64_t b = 0;
        int rc = shards[d]->export_rows(vals[d].data(), vals[d].size(), &idsb[d][0], idsb[d].size(), offs[d].data(), &r, &b);
        if (rc) return rc;
    }
    size_t B = 0;
    for (const Set *sh : shards) B = std::max(B, sh->blocks.size());
    std::vector<int64_t> cursor((size_t)G, 0);
    int64_t k = 0;
    uint64_t io = 0;
    out_off[0] = 0;
    for (size_t pos = 0; pos < B; ++pos)
        for (int d = 0; d < G; ++d) {
            if (pos >= shards[d]->blocks.size()) continue;
            const int live = shards[d]->blocks[pos]->live_count;
            for (int j = 0; j < live; ++j) {
                const int64_t c = cursor[d]++;
                memcpy(out_values + (size_t)k * rowb, vals[d].data() + (size_t)c * rowb, rowb);
                const uint64_t len = offs[d][c + 1] - offs[d][c];
                memcpy(out_ids + io, idsb[d].data() + offs[d][c], len);
                io += len;
                out_off[++k] = io;
            }
        }
    return GF_OK;
}

int Set::multi_destroy()
{
    ExclusiveGuard g(lock);
    if (destroyed) return GF_ERR_FEATURE_SET_NOT_FOUND;
    destroyed = true;
    int first_rc = GF_OK;
    for (Set *sh : shards) {
        int rc = sh->destroy();
        {
            std::lock_guard<std::mutex> cg(sh->cache->mu);
            sh->cache->sets.erase(sh->name);
        }
        if (rc && !first_rc) first_rc = rc;
    }
    scanned_rows = 0;
    return first_rc;
}

// ------------------------------------------------------------------ search
bool Set::multi_direct_applies(int limit, int q, int *kb) const
{
    if (!ctx->peer_ok || precision != 4 || scanned_rows <= 0 || limit <= 0 || q <= 0) return false;
    const int G = (int)shards.size();
    int k0 = 0;
    for (int d = 0; d < G; ++d) {
        int kd = 0;
        // every shard must take the tensor path and hold at least `limit` rows (a shorter shard's list would be
        // zero padded, which the fold handles, but its own k differs: leave those sets to the generic path)
        if (!shards[d]->direct_applies(limit, q, &kd)) return false;
        if (d == 0) k0 = kd;
        if (kd != k0 || kd != limit) return false;
    }
    if ((int64_t)G * k0 > kExchangeMaxKeys) return false;
    *kb = k0;
    return true;
}

static int ensure_ext(Set *P, Workspace *ws, int q, int kb)
{
    const int G = (int)P->shards.size();
    MultiExt *x = ws->multi;
    if (x == nullptr) {
        x = new MultiExt();
        ws->multi = x;
        x->ctxs.resize((size_t)G);
        x->ws.assign((size_t)G, nullptr);
        x->inbox.assign((size_t)G, nullptr);
        x->d_peers.assign((size_t)G, nullptr);
        for (int d = 0; d < G; ++d) {
            x->ctxs[d] = P->shards[d]->ctx;
            int rc = x->ctxs[d]->bind();
            if (rc) return rc;
            x->ws[d] = x->ctxs[d]->acquire_ws();
            if (!x->ws[d]) return GF_ERR_CUDA;
        }
    }
    if (q > x->max_q || kb > x->max_k) {
        // grow the inboxes (and forget every chain recorded against the old ones)
        const int mq = std::max(std::max(q, 16), x->max_q), mk = std::max(std::max(kb, 16), x->max_k);
        const size_t bytes = exchange_inbox_bytes(G, mq, mk);
        std::vector<void *> fresh((size_t)G, nullptr);
        for (int d = 0; d < G; ++d) {
            int rc = x->ctxs[d]->bind();
            if (rc) return rc;
            GF_CUDA(cudaStreamSynchronize(x->ws[d]->stream), GF_ERR_CUDA);
            x->ws[d]->drop_graphs();
            if (x->inbox[d]) cudaFree(x->inbox[d]);
            x->inbox[d] = nullptr;
            GF_CUDA(cudaMalloc(&fresh[d], bytes), GF_ERR_ALLOCATE_GPU_BUFFER);
            GF_CUDA(cudaMemset(fresh[d], 0, bytes), GF_ERR_CLEAR_CUDA_BUFFER);
            x->inbox[d] = fresh[d];
        }
        for (int d = 0; d < G; ++d) {
            x->ctxs[d]->bind();
            if (!x->d_peers[d]) GF_CUDA(cudaMalloc((void **)&x->d_peers[d], sizeof(void *) * G), GF_ERR_ALLOCATE_GPU_BUFFER);
            GF_CUDA(cudaMemcpy(x->d_peers[d], x->inbox.data(), sizeof(void *) * G, cudaMemcpyHostToDevice), GF_ERR_CUDA);
            GF_CUDA(cudaDeviceSynchronize(), GF_ERR_CUDA);
        }
        x->max_q = mq;
        x->max_k = mk;
        x->dirty = false;
    }
    if (x->dirty) {
        for (int d = 0; d < G; ++d) {
            x->ctxs[d]->bind();
            cudaDeviceSynchronize();
            cudaMemset(x->inbox[d], 0, exchange_inbox_bytes(G, x->max_q, x->max_k));
            cudaDeviceSynchronize();
        }
        x->dirty = false;
    }
    return GF_OK;
}

// Device part of one search of q queries over all shards; the merged keys [q][kb] are left in the ext's pinned
// buffer.  The parent's read lock is held by the caller.
int Set::multi_direct_keys(Workspace *ws, int kb, int q, const uint8_t *const *parts, const int *part_q, int nparts,
                           unsigned long long **h_keys)
{
    const int G = (int)shards.size();
    int rc = ensure_ext(this, ws, q, kb);
    if (rc) return rc;
    MultiExt *x = ws->multi;
    const size_t qbytes = (size_t)q * dims * 4;
    const size_t keyb = align_up((size_t)q * kb * 8, 256);
    // pinned: queries | merged keys | merged flags | per-shard keys of re-run queries
    const size_t h_need = align_up(qbytes, 256) + keyb + align_up((size_t)q * 4, 256) + (size_t)G * keyb;
    if (h_need > x->h_cap) {
        for (int d = 0; d < G; ++d) {
            x->ctxs[d]->bind();
            GF_CUDA(cudaStreamSynchronize(x->ws[d]->stream), GF_ERR_CUDA);
            x->ws[d]->drop_graphs();  // they read the queries from the old buffer
        }
        if (x->h_pin) cudaFreeHost(x->h_pin);
        x->h_pin = nullptr;
        x->h_cap = 0;
        const size_t want = std::max(h_need, (size_t)1 << 16);
        GF_CUDA(cudaHostAlloc((void **)&x->h_pin, want, cudaHostAllocPortable), GF_ERR_ALLOCATE_GPU_BUFFER);
        x->h_cap = want;
    }
    const auto t_a = std::chrono::steady_clock::now();
    float *h_q = (float *)x->h_pin;
    unsigned long long *h_out = (unsigned long long *)(x->h_pin + align_up(qbytes, 256));
    unsigned int *h_flags = (unsigned int *)((uint8_t *)h_out + keyb);
    unsigned long long *h_shard = (unsigned long long *)((uint8_t *)h_flags + align_up((size_t)q * 4, 256));
    size_t off = 0;
    for (int i = 0; i < nparts; ++i) {
        memcpy((uint8_t *)h_q + off, parts[i], (size_t)part_q[i] * dims * 4);
        off += (size_t)part_q[i] * dims * 4;
    }
    ctx->stats.h2d_bytes += qbytes * G;  // every device reads the batch over PCIe
    // sentinels: the root's tail writes every key and flag word exactly once (wait_results, gf_runtime.cu); a device
    // that timed out leaves 0xFFFFFFFF in a flag, and the wait then ends with the stream itself
    const bool poll_results = (size_t)q * kb <= 8192;  // (large answers: filling and scanning the area costs more than it saves)
    if (poll_results) {
        memset(h_out, 0xFF, (size_t)q * kb * 8);
        memset(h_flags, 0xFF, (size_t)q * 4);
    }

    // per-device scratch: local winners | local flags | device copy of the queries (re-runs) | scan lists at the tail
    std::vector<unsigned long long *> d_loc((size_t)G);
    std::vector<unsigned int *> d_lfl((size_t)G);
    std::vector<float *> d_q((size_t)G);
    std::vector<ExchangeArgs> xa((size_t)G);
    const int root = 0;
    bool launched_any = false;
    const auto t_b = std::chrono::steady_clock::now();
    auto fail = [&](int code) {
        if (launched_any) x->dirty = true;
        return code;
    };
    for (int step = 0; step < G; ++step) {
        const int d = (root + 1 + step) % G;  // the root goes last: its tail waits for everybody else's push
        Set *sh = shards[d];
        Workspace *cws = x->ws[d];
        rc = sh->ctx->bind();
        if (rc) return fail(rc);
        const int nseg = (int)sh->h_segs.size();
        const int maxq = std::max(1, scan_max_q(dims));
        const size_t lists = (size_t)sh->ctx->sms * std::min(q, maxq) * kb * 8;
        (void)nseg;
        const size_t d_need = keyb + align_up((size_t)q * 4, 256) + align_up(qbytes, 256) + align_up(lists, 256) + 256;
        rc = cws->reserve(d_need, 0);
        if (rc) return fail(rc);
        d_loc[d] = (unsigned long long *)cws->d_buf;
        d_lfl[d] = (unsigned int *)(cws->d_buf + keyb);
        d_q[d] = (float *)(cws->d_buf + keyb + align_up((size_t)q * 4, 256));
        ExchangeArgs &a = xa[d];
        a = ExchangeArgs{};
        a.peers = x->d_peers[d];
        a.rank = d;
        a.world = G;
        a.q = q;
        a.K = kb;
        a.max_q = x->max_q;
        a.max_k = x->max_k;
        a.seq = 1;
        a.local_keys = d_loc[d];
        a.local_flags = d_lfl[d];
        a.out_keys = d == root ? h_out : nullptr;
        a.out_flags = d == root ? h_flags : nullptr;
        a.timeout_ns = 5000000000ull;
        a.gather_root_plus1 = root + 1;
        cws->xchg = &a;
        cws->xchg_used = false;
        rc = sh->topk_device(cws, cws->stream, h_q, q, dims, 1, kb, false, nullptr, d_loc[d], search_path, d_lfl[d], true);
        const bool fused = cws->xchg_used;
        cws->xchg = nullptr;
        if (rc) return fail(rc);
        launched_any = true;
        if (!fused) {
            cudaError_t e = exchange_args_launch(a, cws->stream);
            if (e != cudaSuccess) return fail(cuda_fail(e, "exchange_args_launch", GF_ERR_CUDA));
            sh->ctx->stats.merge_launches++;
        }
    }
    const auto t_c = std::chrono::steady_clock::now();
    shards[root]->ctx->bind();
    cudaError_t we = poll_results ? wait_results(x->ws[root]->stream, h_out, (size_t)q * kb, h_flags, (size_t)q)
                                  : wait_stream(x->ws[root]->stream);
    if (we != cudaSuccess) return fail(cuda_fail(we, "multi-device search", GF_ERR_READ_CUDA_BUFFER));
    for (int d = 0; d < G; ++d) {  // the others pushed their winners before the root could fold: a query, rarely a wait
        if (d == root) continue;
        cudaError_t e = cudaStreamQuery(x->ws[d]->stream);
        if (e == cudaErrorNotReady) e = cudaStreamSynchronize(x->ws[d]->stream);
        if (e != cudaSuccess) return fail(cuda_fail(e, "multi-device search", GF_ERR_READ_CUDA_BUFFER));
    }
    const auto t_d = std::chrono::steady_clock::now();
    ctx->stats.host_stage_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t_b - t_a).count();
    ctx->stats.host_launch_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t_c - t_b).count();
    ctx->stats.host_wait_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t_d - t_c).count();
    ctx->stats.host_calls++;
    ctx->stats.d2h_bytes += (size_t)q * kb * 8 + (size_t)q * 4;
    // queries some shard could not certify: every shard re-runs JUST those through its exact scan, the host folds
    bool any = false;
    for (int i = 0; i < q; ++i) {
        if (h_flags[i] == 0xFFFFFFFFu) {
            set_error_detail("multi-device search: a device did not deliver its winners within 5 s");
            return fail(GF_ERR_CUDA);
        }
        any = any || h_flags[i] != 0u;
    }
    if (any) {
        std::vector<int> bad;
        for (int i = 0; i < q; ++i)
            if (h_flags[i]) bad.push_back(i);
        ctx->stats.uncertified_queries += bad.size();
        const int nb = (int)bad.size();
        for (int d = 0; d < G; ++d) {
            Set *sh = shards[d];
            Workspace *cws = x->ws[d];
            rc = sh->ctx->bind();
            if (rc) return rc;
            for (int j = 0; j < nb; ++j)  // the refused queries, packed: one compact batch per shard
                GF_CUDA(cudaMemcpyAsync(d_q[d] + (size_t)j * dims, h_q + (size_t)bad[j] * dims, (size_t)dims * 4,
                                        cudaMemcpyHostToDevice, cws->stream),
                        GF_ERR_WRITE_INPUT_BUFFER);
            rc = sh->topk_scan(cws, cws->stream, d_q[d], nb, dims, 1, kb, false, nullptr, d_loc[d]);
            if (rc) return rc;
            GF_CUDA(cudaMemcpyAsync(h_shard + (size_t)d * q * kb, d_loc[d], (size_t)nb * kb * 8, cudaMemcpyDeviceToHost,
                                    cws->stream),
                    GF_ERR_READ_CUDA_BUFFER);
        }
        for (int d = 0; d < G; ++d) {
            shards[d]->ctx->bind();
            GF_CUDA(cudaStreamSynchronize(x->ws[d]->stream), GF_ERR_READ_CUDA_BUFFER);
        }
        std::vector<unsigned long long> all((size_t)G * kb);
        for (int j = 0; j < nb; ++j) {
            for (int d = 0; d < G; ++d)
                memcpy(all.data() + (size_t)d * kb, h_shard + (size_t)d * q * kb + (size_t)j * kb, (size_t)kb * 8);
            std::sort(all.begin(), all.end(), std::greater<unsigned long long>());
            memcpy(h_out + (size_t)bad[j] * kb, all.data(), (size_t)kb * 8);
        }
    }
    *h_keys = h_out;
    return GF_OK;
}

// set.go:118-159 over the shards
int Set::multi_search_now(float threshold, int limit, int q, const uint8_t *queries, Result **out)
{
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
    int kb = 0;
    if (multi_direct_applies(limit, q, &kb)) {
        Workspace *ws = ctx->acquire_ws();
        if (!ws) return GF_ERR_CUDA;
        unsigned long long *h_keys = nullptr;
        int rc = multi_direct_keys(ws, kb, q, &queries, &q, 1, &h_keys);
        bool quirk = false;
        const auto t_r0 = std::chrono::steady_clock::now();
        if (rc == GF_OK)
            for (int i = 0; i < q && !quirk; ++i) quirk = !resolve_winners(h_keys + (size_t)i * kb, kb, threshold, res.get());
        ctx->stats.host_resolve_ns +=
            (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_r0).count();
        ctx->release_ws(ws);
        if (rc) return rc;
        if (!quirk) {
            *out = res.release();
            return GF_OK;
        }
        res.reset(new Result());  // a tombstone among the winners: block.go:236-243 is exact only per block
        res->begin(q);
        ctx->stats.quirk_fallbacks++;
    }
    // generic path (small sets, bands, tombstones among the winners): every shard answers with ITS Search -- the
    // per-block rule applied to its blocks, threshold applied, top-limit -- and the hit lists are merged
    // (MaxNFeatureResult over the union, set.go:154-157).  top-limit of a union of top-limits is exact.
    const int G = (int)shards.size();
    std::vector<Result *> part((size_t)G, nullptr);
    std::vector<int> rcs((size_t)G, GF_OK);
    std::vector<std::thread> th;
    for (int d = 0; d < G; ++d)
        th.emplace_back([&, d]() { rcs[d] = shards[d]->search_now(threshold, limit, q, queries, &part[d]); });
    for (auto &t : th) t.join();
    int rc = GF_OK;
    for (int d = 0; d < G; ++d)
        if (rcs[d] && !rc) rc = rcs[d];
    if (rc == GF_OK) {
        struct Hit {
            unsigned long long key;
            int shard;
            int64_t idx;
        };
        std::vector<Hit> hits;
        for (int i = 0; i < q; ++i) {
            hits.clear();
            for (int d = 0; d < G; ++d)
                for (int64_t k = part[d]->offsets[i]; k < part[d]->offsets[i + 1]; ++k)
                    hits.push_back({make_key(part[d]->scores[k], (uint32_t)part[d]->slots[k]), d, k});
            std::sort(hits.begin(), hits.end(), [](const Hit &a, const Hit &b) { return a.key > b.key; });
            if (hits.size() > (size_t)limit) hits.resize((size_t)limit);
            for (const Hit &h : hits) {
                const Result *p = part[h.shard];
                res->push(p->scores[h.idx], p->slots[h.idx],
                          p->id_bytes.substr(p->id_off[h.idx], p->id_off[h.idx + 1] - p->id_off[h.idx]));
            }
            res->end_query();
        }
    }
    for (Result *p : part) delete p;
    if (rc) return rc;
    *out = res.release();
    return GF_OK;
}
