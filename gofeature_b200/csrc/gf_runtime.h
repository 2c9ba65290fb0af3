// gf_runtime.h -- host object model behind the C ABI (internal, C++17).
//
// Mirrors the reference's layers: Context (unixpickle/cuda Context), Buffer
// (buffer.go), Block (block.go), Set (set.go), Cache (cache.go).  Citations are
// /root/reference/<file>:<line>.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <future>
#include <unordered_map>
#include <vector>

#include "../../include/gofeature_b200.h"
#include "gf_kernels.h"

// The opaque handle types of the C header ARE these structs.
struct gf_context;
struct gf_buffer;
struct gf_cache;
struct gf_block;
struct gf_set;
struct gf_result;

namespace gf {

using Context = ::gf_context;
using Buffer = ::gf_buffer;
using Cache = ::gf_cache;
using Block = ::gf_block;
using Set = ::gf_set;
using Result = ::gf_result;

void set_error_detail(const std::string &s);
uint64_t next_set_version();
int cuda_fail(cudaError_t e, const char *what, int status = GF_ERR_CUDA);

#define GF_CUDA(call, status)                                   \
    do {                                                        \
        cudaError_t _e = (call);                                \
        if (_e != cudaSuccess) return gf::cuda_fail(_e, #call, status); \
    } while (0)

// writer-preferring reader/writer lock: searches share, Add/Delete/Update exclude
class RWLock {
   public:
    void lock_shared();
    void unlock_shared();
    void lock();
    void unlock();

   private:
    std::mutex m_;
    std::condition_variable cv_;
    int readers_ = 0;
    int writers_waiting_ = 0;
    bool writer_ = false;
};

struct Stats {
    std::atomic<uint64_t> searches{0}, scan_launches{0}, gemm_launches{0}, merge_launches{0}, other_launches{0},
        rows_scanned{0}, quirk_fallbacks{0}, coalesced_batches{0}, h2d_bytes{0}, d2h_bytes{0}, scan_kernel_ns{0},
        scan_kernel_timed{0}, tensor_passes{0}, uncertified_queries{0}, gemm_kernel_ns{0}, gemm_kernel_timed{0},
        host_stage_ns{0}, host_launch_ns{0}, host_wait_ns{0}, host_resolve_ns{0}, host_calls{0},
        add_stage_ns{0}, add_place_ns{0}, add_index_ns{0}, add_rows{0}, tier2_queries{0};
};

// per-search scratch: a stream, device and pinned arenas
struct Workspace {
    cudaStream_t stream = nullptr;
    uint8_t *d_buf = nullptr;
    size_t d_cap = 0;
    uint8_t *h_buf = nullptr;
    size_t h_cap = 0;
    uint8_t *d_tensor = nullptr;  // scratch of the tcgen05 path (queries, tile maxima, candidates ...)
    size_t d_tensor_cap = 0;
    // CUDA graphs of whole search pipelines (one per distinct call shape), replayed to shed launch gaps
    struct CachedGraph {
        const void *set;
        uint64_t version;
        int q, K, mode;
        const void *src;  // baked-in query source / result destination (Search's pinned staging buffers), or
        void *dst;        // null: the query preparation and the result copy stay outside the graph
        const void *xkey; // baked-in cross-device exchange (its inbox table), or null
        bool xchg_in_graph;  // ... and whether the recorded chain performs it (one-pass chain) or the caller launches it
        cudaGraphExec_t exec;
        uint64_t d_scan, d_gemm, d_merge, d_other, d_tensor_passes, d_rows;  // counters one replay adds
        uint64_t stamp;
    };
    std::vector<CachedGraph> graphs;
    uint64_t graph_clock = 0;
    bool graphs_broken = false;
    unsigned long long *t_res = nullptr;  // tensor path: internal result keys / flags of the last carve
    unsigned int *t_flags = nullptr;
    // recording a graph whose results go to the caller's own (stable, possibly pinned host) buffers: the one-pass
    // chain's last kernel writes them itself and sets wrote_direct -- no copy node behind it
    unsigned long long *direct_out = nullptr;
    unsigned int *direct_flags = nullptr;
    bool wrote_direct = false;
    // a sharded caller's cross-rank step: the one-pass chain's tail ends with it (sets xchg_used) instead of a launch
    const ExchangeArgs *xchg = nullptr;
    bool xchg_used = false;
    // Overlapped one-pass chains (Set::topk_tensor): two scratch sets alternate with the step's parity so that a
    // step's tail runs next to the following step's candidate pass; done_total[p] = tail clusters launched so far on
    // set p (the device-side completion counters live at the head of d_tensor and are compared against these)
    uint64_t op_step = 0;
    unsigned int done_total[2] = {0u, 0u};
    unsigned int prep_total[2] = {0u, 0u};  // preparation CTAs launched so far per scratch set
    struct MultiExt *multi = nullptr;  // a multi-device set's per-search state (child workspaces, inboxes): gf_multi.cu
    void drop_graphs();
    int reserve(size_t d_bytes, size_t h_bytes);
    int reserve_tensor(size_t bytes);
    ~Workspace();
};

// What one in-flight search of a multi-device set (gf_context_create_multi) owns besides its parent workspace:
// a workspace on every device, an inbox on every device for the gather-to-root exchange, and one portable
// pinned buffer every device reads the queries from and the root device writes the merged winners to.
struct MultiExt {
    std::vector<::gf_context *> ctxs;
    std::vector<Workspace *> ws;
    std::vector<void *> inbox;
    std::vector<void **> d_peers;
    int max_q = 0, max_k = 0;
    uint8_t *h_pin = nullptr;
    size_t h_cap = 0;
    bool dirty = false;  // a step failed half way: clear the inboxes before the next one
    ~MultiExt();
};

}  // namespace gf

struct gf_context {
    using Workspace = gf::Workspace;
    using Stats = gf::Stats;
    int device = 0;
    int sms = 0;
    cudaStream_t mut_stream = nullptr;  // Add / Delete / Update traffic
    uint8_t *stage[2] = {nullptr, nullptr};  // pinned upload ring
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    std::mutex stage_mu;
    // Add staging on the device (rows | shadow rows | row scales of one Add chunk): filled and quantised BEFORE the
    // set's exclusive lock is taken, copied into the blocks device-to-device under it (Set::add)
    std::mutex add_mu;
    uint8_t *d_add = nullptr;
    size_t d_add_cap = 0;
    // helper threads of upload(): the copy from the caller's pageable rows into the pinned ring is what bounds Add
    // (one core copies ~8 GB/s, the link takes ~50): three parked helpers + the caller split every chunk
    struct CopyPool *copy_pool = nullptr;
    void parallel_copy(void *dst, const void *src, size_t bytes);
    Stats stats;
    // optional CUDA-event timing of the scan launches
    std::atomic<int> profiling{0};
    std::mutex prof_mu;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_pending, prof_pending_gemm, prof_free;
    void prof_begin(cudaStream_t st, std::pair<cudaEvent_t, cudaEvent_t> *ev);
    void prof_end(cudaStream_t st, const std::pair<cudaEvent_t, cudaEvent_t> &ev, bool gemm = false);
    void prof_collect();  // waits for the pending events and folds them into stats
    std::mutex ws_mu;
    std::vector<Workspace *> ws_free;
    std::vector<Workspace *> ws_all;

    // gf_context_create_multi: this context lives on devices[0] (buffers, bare blocks) and owns one child context
    // per device (children[0] is a second context on devices[0]); caches created on it stripe their blocks
    std::vector<gf_context *> children;
    bool peer_ok = false;  // every child device can store into the root device's memory
    bool is_multi() const { return !children.empty(); }
    int bind() const;  // cudaSetDevice
    Workspace *acquire_ws();
    void release_ws(Workspace *w);
    // pageable host -> device through the pinned ring, on mut_stream; returns after the copy finished
    int upload(void *dst, const void *src, size_t bytes);
    // n rows of `rowb` bytes, row i to base + slots[i]*rowb: staged through the ring, one sync at the end
    int upload_scatter(uint8_t *base, const int *slots, size_t n, size_t rowb, const uint8_t *src);
    int zero_rows(float *d_slab, void *d_shadow, int shadow_eb, const std::vector<unsigned long long> &elem_off,
                  int dims);
    ~gf_context();
};

namespace gf {
struct DeviceAlloc {
    Context *ctx;
    void *ptr;
    uint64_t size;
    bool owned;
    DeviceAlloc(Context *c, void *p, uint64_t s, bool o) : ctx(c), ptr(p), size(s), owned(o) {}
    DeviceAlloc(const DeviceAlloc &) = delete;
    DeviceAlloc &operator=(const DeviceAlloc &) = delete;
    ~DeviceAlloc();
};

}  // namespace gf

struct gf_buffer {
    gf::Context *ctx;
    std::shared_ptr<gf::DeviceAlloc> root;
    uint8_t *ptr;
    uint64_t size;
};

struct gf_result {
    int batch = 0;
    std::vector<int64_t> offsets;  // batch + 1
    std::vector<float> scores;
    std::vector<uint64_t> slots;
    std::vector<uint64_t> id_off;  // total + 1
    std::string id_bytes;
    void begin(int q);
    void push(float score, uint64_t slot, const std::string &id);
    void end_query();
};

struct gf_block {
    gf::Cache *cache = nullptr;
    int index = 0;
    int64_t block_size = 0;
    uint8_t *dev = nullptr;
    uint8_t *shadow = nullptr;  // this block's rows in the cache's shadow slab (null: no shadow)
    int shadow_eb = 0;          // bytes per shadow element: 2 bf16, 1 int8 (+ one float scale per row)
    std::mutex mu;  // block.go:19
    // feature info (block.go:24-31)
    int dims = 0, precision = 0, batch = 0;
    std::string owner;
    std::vector<int> empty;
    int next_index = 0;
    std::vector<std::string> ids;  // explicit IDs[]; unused while implicit
    std::vector<uint8_t> live;     // live[j] = (IDs[j] != "")
    int live_count = 0;
    // implicit ids "<prefix>%011d" of id_first + slot (synthetic fill).  Rows that later take an arbitrary id
    // (a tombstone refilled by Add, an explicit tail append) are kept in `over`; the block is only turned into
    // explicit strings (materialize_ids) when a quarter of it has been overridden or it is compacted
    bool implicit_ids = false;
    std::string id_prefix;
    int64_t id_first = 0;
    std::unordered_map<int, std::string> over;  // row -> explicit id, implicit blocks only
    std::unordered_map<std::string, std::vector<int>> id_index;  // id -> rows holding it (when use_index)
    bool use_index = true;  // false while a set keeps the set-wide locator for this block
    gf::Set *set = nullptr;

    int capacity() const { return (dims > 0 && precision > 0) ? (int)(block_size / ((int64_t)precision * dims)) : 0; }
    int margin() const { return (int)empty.size() + (capacity() - next_index); }
    bool is_owned() const { return !owner.empty(); }
    std::string id_at(int row) const;
    int find_last(const std::string &id) const;  // largest row holding id, -1 if none (block.go:140-144)
    void materialize_ids();
    int acquire(const std::string &owner_, int dims_, int precision_, int batch_);
    int release();
    // values: n rows of dims*precision bytes; ids[i] non-empty
    int insert(int64_t n, const uint8_t *values, const std::vector<std::string> &ids_);
    // the same with the rows (and their shadow rows / scales, when the cache keeps a shadow) already on the device:
    // device-to-device copies only; ids are MOVED out of ids_[first .. first + n)
    int insert_staged(int64_t n, const float *d_rows, const uint8_t *d_shadow_rows, const float *d_scales,
                      std::vector<std::string> &ids_, int64_t first, std::vector<int> *placed_rows);
    int insert_synthetic(int64_t n, const gf::SynthSpec &seed, int64_t first_ordinal, int normalize, const std::string &prefix);
    int remove(const std::vector<std::string> &ids_, std::vector<uint8_t> &found);
    // tombstone the rows; zero them on the device (no stream sync) or, with zero_later, only report their element
    // offsets in the slab so that the caller zeroes the whole Delete with one kernel
    int remove_rows(const std::vector<int> &rows, std::vector<unsigned long long> *zero_later = nullptr);
    void rebuild_index();
    void index_add(const std::string &id, int row);
    void index_del(const std::string &id, int row);
};

namespace gf {
// one caller of gf_set_search waiting to be served by a coalesced scan pass
struct PendingSearch {
    const uint8_t *queries;
    int q;
    int limit;
    float threshold;
    gf::Result *result = nullptr;
    int rc = 0;
    std::atomic<bool> done{false};
    bool returnee = false;  // this thread was served by the immediately preceding batch
    // fast hand-off (tensor path, one band): the leader publishes raw keys and every caller builds its own
    // result from them, in parallel, while the leader keeps the set's read lock
    const unsigned long long *keys = nullptr;  // [q][kb] inside the leader's pinned staging buffer
    int kb = 0;
    std::atomic<int> *resolvers = nullptr;     // decremented once this caller no longer reads `keys`
    uint64_t batch_id = 0;                     // the batch that served this caller
};
}

struct gf_set {
    using Segment = gf::Segment;
    using Workspace = gf::Workspace;
    using Block = gf::Block;
    using Result = gf::Result;
    gf::Cache *cache = nullptr;
    gf::Context *ctx = nullptr;
    std::string name;
    int dims = 0, precision = 0, batch = 0;
    int cap = 0;  // BlockFeatureNum (cache.go:59)
    std::vector<Block *> blocks;
    gf::RWLock lock;
    // Mutations (Add / Delete / Update / Compact / Destroy, and Read, which walks the id locator) are serialised by
    // mut_mu; the exclusive side of `lock` is only held while something a SEARCH reads changes (device rows inside
    // NextIndex, ids / live flags, the segment table) -- host staging, H2D, shadow quantisation and the id locator
    // are worked on outside it, so searches stall for microseconds, not for the length of the upload
    std::mutex mut_mu;
    std::atomic<bool> destroyed{false};  // read before any lock is taken (TSAN: tools/race_test)
    int shard_rank = 0, shard_world = 1;
    // device segment table of the scanned rows, rebuilt after every mutation that moves NextIndex
    std::vector<Segment> h_segs;
    std::vector<int> seg_block;  // segment -> position in blocks
    Segment *d_segs = nullptr;
    int d_segs_cap = 0;
    // tcgen05 path: 128-row tile prefix per segment, slot base per segment, |row|^2 bound, counters
    uint32_t *d_gtile_start = nullptr;
    uint32_t *d_seg_slot_base = nullptr;
    uint32_t total_gtiles = 0;
    unsigned int *d_small = nullptr;  // [0] max |row|^2 bits, [1] queries the flag-less device call could not repair, [2] max |shadow(row) - row|^2 bits, [3] queries the certificate refused
    bool tensor_ok = false;
    bool shadow_ok = false;   // the shadow of this set's rows is maintained and usable by the tensor path
    std::atomic<int> search_path{0};      // path the host Search takes (gf_set_set_search_path): 0 choose, 1 scan, 2 tensor, 3 tensor tf32, 4 / 5 = 2 / 3 without the one-pass chain
    uint64_t version = 0;  // bumped whenever the segment table or the shard identity changes
    uint32_t total_tiles = 0;
    int64_t scanned_rows = 0;
    int tile_rows = 0;
    bool aligned16 = false;
    Workspace *dev_ws = nullptr;  // scratch of gf_set_search_device (stream-ordered calls)
    // set-wide id locator: O(1) Delete / Update / Read instead of block.go:140-144's scan of every block.
    // Exact while ids are unique; the first duplicate id switches to the reference's block walk.
    struct ImplicitRange {
        int64_t first;
        int count;
        int pos;
    };
    std::unordered_map<std::string, uint64_t> loc;  // explicit id -> (block position << 32 | row)
    std::vector<ImplicitRange> implicit_ranges;     // sorted by first; ids "<implicit_prefix>%011d"
    std::string implicit_prefix;
    bool has_dups = false;
    // Add's last phase -- the locator learns the new ids -- runs on a helper thread while the caller returns (and stages
    // its next Add); everything that reads or writes the locator, the blocks' id tables or `has_dups` joins it first
    mutable std::mutex loc_job_mu;
    mutable std::future<void> loc_job;
    void loc_sync() const
    {
        std::lock_guard<std::mutex> g(loc_job_mu);
        if (loc_job.valid()) loc_job.get();
    }
    bool find(const std::string &id, int *pos, int *row) const;
    bool find_implicit(const std::string &id, int *pos, int *row) const;
    void note_id(const std::string &id, int pos, int row);
    void enter_dup_mode();
    void materialize_block(int pos);
    // coalescer
    std::atomic<int> coalesce_wait_us{0};
    std::mutex co_mu;
    std::condition_variable co_cv;
    std::vector<gf::PendingSearch *> co_queue;
    std::atomic<bool> co_leader_active{false};
    std::atomic<int> co_queued{0};  // co_queue.size(), readable without the lock
    std::atomic<int> co_inside{0};  // callers currently inside search_coalesced (spin only when cores are spare)
    uint64_t co_batch_id = 0;   // batches served so far
    int co_last_group = 0;      // requests in the last batch

    // ---- a set of a multi-device cache: no blocks of its own, one shard per device (shard d = rank d of
    // shards.size(), owned by the child cache); every public operation goes through the parent (gf_multi.cu)
    std::vector<gf_set *> shards;
    bool is_multi() const { return !shards.empty(); }
    void multi_refresh();  // scanned_rows / version after a mutation
    int multi_add(int64_t n, const uint8_t *values, const std::vector<std::string> &ids_);
    int multi_add_synthetic(int64_t n, const gf::SynthSpec &seed, int64_t first_ordinal, int normalize, const std::string &prefix);
    int multi_remove(const std::vector<std::string> &ids_, std::vector<uint8_t> &found);
    int multi_update(int64_t n, const uint8_t *values, const std::vector<std::string> &ids_, std::vector<uint8_t> &found);
    int multi_read(const std::vector<std::string> &ids_, uint8_t *out, std::vector<uint8_t> &found);
    int multi_search_now(float threshold, int limit, int q, const uint8_t *queries, Result **out);
    int multi_direct_keys(Workspace *ws, int kb, int q, const uint8_t *const *parts, const int *part_q, int nparts,
                          unsigned long long **h_keys);
    bool multi_direct_applies(int limit, int q, int *kb) const;
    int multi_destroy();
    int multi_compact(int64_t *reclaimed);
    int multi_export(uint8_t *out_values, uint64_t values_cap, char *out_ids, uint64_t ids_cap, uint64_t *out_off,
                     int64_t *out_rows, uint64_t *out_id_bytes);
    // owner shard of an id: the first block in GLOBAL block order holding it (set.go:167-170)
    bool multi_find(const std::string &id, int *shard, int *pos, int *row) const;
    // the block and row of a (global) slot, whichever shard owns it; false: no such block
    bool locate_slot(uint32_t slot, const Block **b, int *row) const;

    uint32_t slot_base_of(int block_pos) const
    {
        return (uint32_t)(((uint64_t)block_pos * shard_world + shard_rank) * (uint64_t)cap);
    }
    int rebuild_segments();  // caller holds the exclusive lock
    int add(int64_t n, const uint8_t *values, const std::vector<std::string> &ids_);
    int add_plain(int64_t n, const uint8_t *values, const std::vector<std::string> &ids_);
    int add_synthetic(int64_t n, const gf::SynthSpec &seed, int64_t first_ordinal, int normalize, const std::string &prefix);
    int remove(const std::vector<std::string> &ids_, std::vector<uint8_t> &found);
    int update(int64_t n, const uint8_t *values, const std::vector<std::string> &ids_, std::vector<uint8_t> &found);
    int read(const std::vector<std::string> &ids_, uint8_t *out, std::vector<uint8_t> &found);
    int search(float threshold, int limit, int q, const uint8_t *queries, Result **out);
    int search_now(float threshold, int limit, int q, const uint8_t *queries, Result **out);
    int search_coalesced(float threshold, int limit, int q, const uint8_t *queries, Result **out);
    // xchg (optional, needs d_flags): the sharded caller's cross-rank step; *xchg_used tells whether the search's
    // last kernel performed it (one-pass chain) or the caller still has to launch exchange_merge_launch
    int search_device(int limit, int q, const float *d_queries, unsigned long long *d_out, int path,
                      cudaStream_t stream, unsigned int *d_flags = nullptr, const gf::ExchangeArgs *xchg = nullptr,
                      bool *xchg_used = nullptr);
    int resolve_keys(float threshold, int limit, int q, const unsigned long long *h_keys, Result **out);
    // device part of a one-band tensor-path search of q queries gathered from `parts` (pointer, count):
    // the keys [q][kb] are left in the workspace's pinned buffer (*h_keys).  Read lock held by the caller.
    int direct_keys(Workspace *ws, int kb, int q, const uint8_t *const *parts, const int *part_q, int nparts,
                    unsigned long long **h_keys);
    bool direct_applies(int limit, int q, int *kb) const;
    // winners of ONE query (zero-terminated or kb keys) -> r, threshold applied (set.go:145); false when a
    // winner is a tombstone: block.go:236-243 then needs the per-block path
    bool resolve_winners(const unsigned long long *keys, int kb, float threshold, Result *r) const;
    int destroy();
    int compact(int64_t *reclaimed);
    int export_rows(uint8_t *out_values, uint64_t values_cap, char *out_ids, uint64_t ids_cap, uint64_t *out_off,
                    int64_t *out_rows, uint64_t *out_id_bytes);  // SURVEY 8f N3  // pack live rows, shrink NextIndex
    // top-K over the scanned rows into d_out; read lock held by the caller.
    //   per_segment: d_out is [pass][segment][qp][K] (qp = queries of that pass), else [q][K]
    //   path: 0 = choose (tensor cores on large sets: any q with the bf16 shadow, q >= 9 without),
    //         1 = fused scan only, 2 = tensor path, 3 = tensor path on the fp32 rows (kind::tf32) even with a shadow
    //   d_flags (optional, [q]): 1 where a tensor-path query could not be certified nor fixed on the device
    int topk_device(Workspace *ws, cudaStream_t stream, const float *d_queries, int q, int64_t q_stride,
                    int64_t l_stride, int K, bool per_segment, const unsigned long long *d_upper,
                    unsigned long long *d_out, int path = 1, unsigned int *d_flags = nullptr, bool bake_io = false);
    int topk_eager(Workspace *ws, cudaStream_t stream, const float *d_queries, int q, int64_t q_stride,
                   int64_t l_stride, int K, bool per_segment, const unsigned long long *d_upper,
                   unsigned long long *d_out, int path, unsigned int *d_flags);
    int topk_scan(Workspace *ws, cudaStream_t stream, const float *d_queries, int q, int64_t q_stride,
                  int64_t l_stride, int K, bool per_segment, const unsigned long long *d_upper,
                  unsigned long long *d_out, const int *d_q_map = nullptr, const unsigned int *d_q_count = nullptr,
                  bool scatter_out = false);  // scatter_out: query i's result goes to row d_q_map[i] of d_out
    //   phase 0: everything, results to d_out / d_flags;  1: query preparation only;
    //   2: everything after the preparation, results to the workspace (ws->t_res / ws->t_flags)
    int topk_tensor(Workspace *ws, cudaStream_t stream, const float *d_queries, int q, int64_t q_stride,
                    int64_t l_stride, int K, unsigned long long *d_out, unsigned int *d_flags, int phase = 0,
                    int path = 0, bool device_fixup = true);
    bool tensor_eligible(int q, int K, bool per_segment, const unsigned long long *d_upper, int path) const;
    // rows were written on the device: refresh their bf16 shadow and the certificate's bounds (mutation stream)
    int note_rows(const float *d_rows, int64_t rows);
    bool use_shadow(int path) const { return shadow_ok && path != 3 && path != 5; }
    static bool two_pass_only(int path) { return path == 4 || path == 5; }  // pin the threshold + COLLECT chain
    // the one-pass chain (prep -> GROUPTOP -> tail) serves this call; kp_out = groups at or above the histogram cut
    bool onepass_applies(int q, int K, int path, bool device_fixup, int *kp_out = nullptr) const;
    bool slot_to_block(uint32_t slot, int *block_pos, int *row) const;
};

struct gf_cache {
    gf::Context *ctx = nullptr;
    int64_t block_size = 0;
    std::shared_ptr<gf::DeviceAlloc> slab;
    // optional shadow of the slab, same row geometry: bf16 (shadow_eb = 2, half the bytes) or int8 with one float
    // scale per row (shadow_eb = 1, a quarter of the bytes; scales: block b, row j at b * (block_size / 512) + j)
    std::shared_ptr<gf::DeviceAlloc> shadow;
    std::shared_ptr<gf::DeviceAlloc> scales;
    int shadow_eb = 0;
    std::vector<std::unique_ptr<gf::Block>> all_blocks;
    std::mutex mu;  // cache.go:17
    std::unordered_map<std::string, gf::Set *> sets;
    std::vector<std::unique_ptr<gf::Set>> owned_sets;  // live and destroyed (handles stay valid)
    // multi-device cache: block i of the cache = block i / G of children[i % G] (cache.go:104-115 hands blocks out
    // in AllBlocks order, so a growing set is striped evenly over the devices); no slab of its own
    std::vector<gf_cache *> children;
    int multi_blocks = 0;
    bool is_multi() const { return !children.empty(); }
    gf::Block *multi_block_at(int i) const
    {
        const int G = (int)children.size();
        return children[i % G]->all_blocks[i / G].get();
    }
};

namespace gf {

// searches on one bare block (Block.Search through the C ABI)
int block_search(Block *b, const Buffer *input, int batch, int limit, Result **out);

}  // namespace gf
