// gf_scan.cu -- fused dot-product + streaming top-K scan (small query batches), sm_100a.
//
// Replaces, for one whole set in ONE launch, what the reference does per block:
// cublasSgemm (block.go:196-209) + D2H of the full score matrix (block.go:214) +
// strided gather and full sort on the CPU (block.go:231-237, util.go:76-93).
//
// Shape of the work: every feature row (dims float32, contiguous, block.go:99) is read
// exactly once from HBM; Q <= 8 queries sit in registers; scores never leave the SM.
// The kernel is HBM-bound: algorithmic bytes per pass = rows*dims*4 (+ Q*dims*4 + lists).
//
//   producer warp (1 elected thread): walks this CTA's tiles, one 1-D bulk copy
//       (cp.async.bulk -> SASS UBLKCP, the TMA engine) of tile_rows*dims*4 = 28 KB per tile
//       into a `stages`-deep shared-memory ring, completion on an mbarrier (complete_tx).
//   7 consumer warps: wait on the tile's "full" barrier, each takes tile_rows/7 rows:
//       128-bit conflict-free LDS, FFMA into the canonical 4 partial sums per thread,
//       xor-butterfly, then a gate against the CTA's current K-th best key.  The rare
//       survivor is appended to a per-(CTA,query) list under a shared-memory lock; a
//       full list is bitonic-sorted by the inserting warp and cut back to K.
//   end: each list is sorted and its K best keys written to d_partial; merge_kernel
//       (gf_merge.cu) folds the per-CTA lists into the final K.
//
// Grid: persistent, one CTA per SM (148 on B200), tiles strided over CTAs; in per-segment
// mode (tombstone parity path, block.go:236-243) gridDim.y = blocks and every CTA stays
// inside one block.
#include <cstdlib>

#include "gf_device.cuh"
#include "gf_kernels.h"

namespace gf {

namespace {

struct ScanArgs {
    const Segment *segs;
    int nseg;
    uint32_t total_tiles;
    const float *queries;
    int q;
    long long q_stride;
    long long l_stride;
    int K;
    int list_cap;
    int stages;
    int tile_rows;
    int dims;
    int per_segment;
    int warp_lists;
    const int *q_map;             // optional: query i of this launch is row q_map[i] of `queries`
    const unsigned int *q_count;  // optional: live query count lives on the device (0 -> nothing to do)
    const unsigned long long *upper;
    unsigned long long *partial;
};

struct TileMeta {
    uint32_t rows;
    uint32_t slot0;
};

constexpr int kConsumerWarps = kScanConsumerWarps;
constexpr int kConsumerThreads = kConsumerWarps * 32;
// iterations after which a mailbox / tile wait gives up with a trap (each iteration is >= ~100 ns: many seconds, against
// waits of microseconds) -- a launch failure the host reports, instead of a GPU that has to be reset
constexpr unsigned int kSpinLimit = 1u << 26;
// threads of the bulk-copy kernel: QG groups of 7 consumer warps + the producer warp
__host__ __device__ constexpr int scan_threads(int qg) { return (kConsumerWarps * qg + 1) * 32; }
static_assert(scan_threads(1) == kScanThreads, "one query group: 7 consumer warps + the producer");

__device__ __forceinline__ void tile_range(const ScanArgs &a, uint32_t &first, uint32_t &end, uint32_t &stride,
                                           uint32_t &group)
{
    if (a.per_segment) {
        group = blockIdx.y;
        uint32_t b = a.segs[group].tile_start;
        end = (group + 1 < (uint32_t)a.nseg) ? a.segs[group + 1].tile_start : a.total_tiles;
        first = b + blockIdx.x;
    } else {
        group = 0;
        first = blockIdx.x;
        end = a.total_tiles;
    }
    stride = gridDim.x;
}

// gate + insert for one (row, query) score; warp-uniform arguments
__device__ __forceinline__ void offer(ListCtl *ctl, unsigned long long *buf, int cap, int K, float score,
                                      uint32_t slot, unsigned long long upper, int lane)
{
    unsigned long long key = make_key(score, slot);
    // (one lane reads the gate and broadcasts it: other warps change it meanwhile, and list_insert is warp-collective)
    unsigned long long tau = 0ull;
    if (lane == 0) tau = *(volatile unsigned long long *)&ctl->tau;
    tau = __shfl_sync(0xffffffffu, tau, 0);
    if (key > tau && key < upper) list_insert(ctl, buf, cap, K, key, lane);
}

// --------------------------------------------------------------------------------------
// Bulk-copy (TMA) pipeline kernel.  D in {128,256,512,1024}; CH = D/128 chunks per lane.
// --------------------------------------------------------------------------------------
//
// QG = query groups.  QT queries in registers cost QT * CH * 4 registers per thread: at QT * CH = 32 (8 queries of
// 512-d) that is 128 of the 255 a thread can have, the CTA is 8 warps and each scheduler has two warps to hide the
// FFMA2 / shuffle / shared-memory latencies of a 250-instruction dependent chain per row pair (ncu: issue active
// 40 %, 5.4 TB/s).  With QG = 2 the same tile is read by two groups of 7 consumer warps, each holding HALF of the
// queries (QW = QT / QG, 128 registers per thread, 15 warps): the arithmetic per SM is unchanged, twice as many
// warps interleave on it, and every feature byte still crosses HBM once (shared memory is read twice).
// WL: owner-computes lists + mailboxes (compile-time: the locked-list build carries none of their code in its loop)
template <int D, int QT, int QG, bool WL>
__global__ void __launch_bounds__(scan_threads(QG), 1) scan_tma_kernel(const ScanArgs a)
{
    constexpr int CH = D / 128;
    constexpr int QW = QT / QG;                      // queries a consumer warp holds
    constexpr int kWarpsAll = kConsumerWarps * QG;   // consumer warps of the CTA
    constexpr int TILE_ROWS = kScanTileBytes / (D * 4);
    constexpr int RPW = TILE_ROWS / kConsumerWarps;  // rows per warp per tile
    constexpr int RB = (RPW >= 2 && QW * CH <= 32) ? 2 : 1;  // rows in flight per warp
    static_assert(RPW >= 1 && RPW % RB == 0, "tile split");
    static_assert(QT % QG == 0 && QW >= 1, "query groups");

    extern __shared__ __align__(128) uint8_t smem[];
    const int stages = a.stages;
    float *tiles = reinterpret_cast<float *>(smem);
    uint8_t *p = smem + (size_t)stages * kScanTileBytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(p);
    uint64_t *empty = full + stages;
    TileMeta *meta = reinterpret_cast<TileMeta *>(empty + stages);
    ListCtl *ctl = reinterpret_cast<ListCtl *>(meta + stages);
    unsigned long long *lists = reinterpret_cast<unsigned long long *>(ctl + QT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_wait();  // launched as a link of the search chain (gf_device.cuh): nothing global is touched before this
    if (threadIdx.x == 0) pdl_launch_dependents();
    int q_live = a.q;
    if (a.q_count != nullptr) {
        const unsigned int c = *a.q_count;
        if (c == 0u) return;  // fix-up launch with nothing to fix
        q_live = c < (unsigned)a.q ? (int)c : a.q;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kWarpsAll);
        }
        mbar_fence_init();
    }
    if (threadIdx.x < QT) list_init(&ctl[threadIdx.x]);
    // lists: [query][list_cap]; owner-computes mode: one mailbox per query behind them, then the done counter
    const int list_words = QT * a.list_cap;
    Mailbox *mbox = reinterpret_cast<Mailbox *>(lists + (size_t)QT * a.list_cap);
    unsigned int *done_warps = reinterpret_cast<unsigned int *>(mbox + QT);  // consumer warps that have posted everything
    for (int i = threadIdx.x; i < list_words; i += blockDim.x) lists[i] = 0ull;
    if (WL)
        for (int i = threadIdx.x; i < (int)(QT * sizeof(Mailbox) / 8) + 1; i += blockDim.x)
            reinterpret_cast<unsigned long long *>(mbox)[i] = 0ull;  // (+ the done counter behind the mailboxes)
    __syncthreads();

    uint32_t first, end, stride, group;
    tile_range(a, first, end, stride, group);

    if (warp == kWarpsAll) {
        // ===================== producer: one thread feeds the ring =====================
        if (lane == 0 && first < end) {
            const uint64_t policy = l2_policy_evict_first();
            uint32_t segi = a.per_segment ? group : 0;
            Segment cur = a.segs[segi];
            uint32_t next_start = (segi + 1 < (uint32_t)a.nseg) ? a.segs[segi + 1].tile_start : a.total_tiles;
            int s = 0;
            uint32_t ph = 0;
            for (uint32_t t = first; t < end; t += stride) {
                mbar_wait(&empty[s], ph ^ 1);
                while (t >= next_start) {
                    ++segi;
                    cur = a.segs[segi];
                    next_start = (segi + 1 < (uint32_t)a.nseg) ? a.segs[segi + 1].tile_start : a.total_tiles;
                }
                uint32_t row0 = (t - cur.tile_start) * TILE_ROWS;
                uint32_t rows = min((uint32_t)TILE_ROWS, cur.rows - row0);
                uint32_t bytes = rows * (uint32_t)(D * 4);
                meta[s].rows = rows;
                meta[s].slot0 = cur.slot_base + row0;
                mbar_arrive_expect_tx(&full[s], bytes);
                bulk_g2s_hint(tiles + (size_t)s * (kScanTileBytes / 4), cur.base + (size_t)row0 * D, bytes, &full[s],
                              policy);
                if (++s == stages) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        return;
    }

    // ===================== consumers =====================
    // warp = qg * kConsumerWarps + rw: row share rw of every tile, queries [qg * QW, qg * QW + QW)
    const int qg = warp / kConsumerWarps, rw = warp - qg * kConsumerWarps;
    const int q0 = qg * QW;
    float4 qreg[QW][CH];
#pragma unroll
    for (int qi = 0; qi < QW; ++qi) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            float v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int e = 128 * j + 4 * lane + c;
                const int qsrc = (q0 + qi < q_live && a.q_map != nullptr) ? a.q_map[q0 + qi] : q0 + qi;
                v[c] = (q0 + qi < q_live) ? a.queries[(long long)qsrc * a.q_stride + (long long)e * a.l_stride] : 0.0f;
            }
            qreg[qi][j] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    // V = RB*QW scores are reduced together; afterwards this lane holds score index vi = b*QW + qi
    constexpr int V = RB * QW;
    constexpr int LOG2V = (V == 1) ? 0 : (V == 2) ? 1 : (V == 4) ? 2 : (V == 8) ? 3 : 4;
    constexpr int GROUP = 32 >> LOG2V;  // lanes that end up with the same score
    const int vi = lane >> (5 - LOG2V);
    const int my_b = vi / QW, my_q = q0 + vi % QW;
    const bool my_live = my_q < q_live;

    const int cap = a.list_cap, K = a.K;
    int s = 0;
    uint32_t ph = 0;
    for (uint32_t t = first; t < end; t += stride) {
        if (WL) {
            // candidates the other warps posted for the queries this warp owns -- also WHILE it waits for the tile: a
            // sender blocked on one of these mailboxes may not have released its stage yet (tiles of several row pairs
            // per warp, 128 / 256-d), and the producer, hence this wait, would then depend on it for ever (rows sorted
            // by score: tests/test_gpu_round2b.py)
            for (unsigned int spins = 0;; ++spins) {
                for (int oq = warp; oq < QT; oq += kWarpsAll)
                    mailbox_drain(&mbox[oq], &ctl[oq], lists + (size_t)oq * cap, cap, K, lane);
                if (__all_sync(0xffffffffu, mbar_try_wait(&full[s], ph))) break;
                if (spins == kSpinLimit) __trap();  // (seconds: a wait that long is a bug -- fail the launch, do not hang the GPU)
            }
        } else {
            mbar_wait(&full[s], ph);
        }
        const uint32_t vrows = meta[s].rows;
        const uint32_t slot0 = meta[s].slot0;
        // one-compare pre-gate, refreshed per tile; a stale (lower) value only lets extra keys through
        const float tau_l = *(volatile float *)&ctl[my_q].tauf;
        const float4 *tile = reinterpret_cast<const float4 *>(tiles + (size_t)s * (kScanTileBytes / 4));
#pragma unroll
        for (int rb = 0; rb < RPW; rb += RB) {
            const int r0 = rw * RPW + rb;
            float4 x[RB][CH];
#pragma unroll
            for (int b = 0; b < RB; ++b)
#pragma unroll
                for (int j = 0; j < CH; ++j) x[b][j] = tile[(size_t)(r0 + b) * (D / 4) + j * 32 + lane];
            if (rb + RB >= RPW) {
                // this warp's last rows of the tile are in registers: hand the stage back to the producer now,
                // the arithmetic, the butterfly and the gate no longer need shared memory
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
            float tsum[V];
#pragma unroll
            for (int b = 0; b < RB; ++b) {
#pragma unroll
                for (int qi = 0; qi < QW; ++qi) {
                    // P[4l..4l+3] as two packed pairs: fma.rn.f32x2 (FFMA2) is two independent IEEE fmas
                    float2 p01 = make_float2(0.f, 0.f), p23 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < CH; ++j) {
                        p01 = __ffma2_rn(make_float2(x[b][j].x, x[b][j].y),
                                         make_float2(qreg[qi][j].x, qreg[qi][j].y), p01);
                        p23 = __ffma2_rn(make_float2(x[b][j].z, x[b][j].w),
                                         make_float2(qreg[qi][j].z, qreg[qi][j].w), p23);
                    }
                    tsum[b * QW + qi] = __fadd_rn(__fadd_rn(p01.x, p01.y), __fadd_rn(p23.x, p23.y));
                }
            }
            const float sc = canonical_reduce_many<V>(tsum, lane);
            const bool pass = my_live && ((uint32_t)(r0 + my_b) < vrows) && !(sc < tau_l);
            unsigned bal = __ballot_sync(0xffffffffu, pass);
            while (bal) {  // rare: a score reached this CTA's current K-th best
                const int leader = __ffs(bal) - 1;
                const int lvi = leader >> (5 - LOG2V);
                const float ls = __shfl_sync(0xffffffffu, sc, leader);
                const unsigned gmask = (GROUP == 32) ? 0xffffffffu : (((1u << GROUP) - 1u) << (lvi * GROUP));
                bal &= ~gmask;
                const int lb = lvi / QW, lq = q0 + lvi % QW;
                const unsigned long long up = a.upper ? a.upper[(size_t)group * a.q + lq] : ~0ull;
                if (WL) {
                    const unsigned long long key = make_key(ls, slot0 + r0 + lb);
                    unsigned long long *lst = lists + (size_t)lq * cap;
                    // (a peek at the owner's K-th key: it only grows, a stale value lets an extra candidate through;
                    // read by one lane and broadcast -- the owner changes it meanwhile, and lanes that disagreed
                    // would split the warp around the collective code below)
                    unsigned long long tau0 = 0ull;
                    if (lane == 0) tau0 = *(volatile unsigned long long *)&ctl[lq].tau;
                    tau0 = __shfl_sync(0xffffffffu, tau0, 0);
                    if (key > tau0 && key < up) {
                        if (lq % kWarpsAll == warp) {
                            owner_list_insert(&ctl[lq], lst, cap, K, key, lane);
                        } else {
                            Mailbox *mb = &mbox[lq];
                            unsigned int idx = 0;
                            if (lane == 0) idx = atomicAdd(&mb->tail, 1u);
                            idx = __shfl_sync(0xffffffffu, idx, 0);
                            // mailbox full: serve this warp's own mailboxes meanwhile (somebody may be waiting on them)
                            for (unsigned int spins = 0;; ++spins) {
                                if (spins == kSpinLimit) __trap();
                                unsigned int hd = 0;
                                if (lane == 0) hd = *(volatile unsigned int *)&mb->head;
                                hd = __shfl_sync(0xffffffffu, hd, 0);
                                if (idx - hd < (unsigned)kMailboxCap) break;
                                for (int oq = warp; oq < QT; oq += kWarpsAll)
                                    mailbox_drain(&mbox[oq], &ctl[oq], lists + (size_t)oq * cap, cap, K, lane);
                            }
                            if (lane == 0) *(volatile unsigned long long *)&mb->slot[idx % kMailboxCap] = key;
                        }
                    }
                } else
                    offer(&ctl[lq], lists + (size_t)lq * cap, cap, K, ls, slot0 + r0 + lb, up, lane);
            }
        }
        if (++s == stages) {
            s = 0;
            ph ^= 1;
        }
    }

    if (WL) {
        // No plain barrier here: a warp that is done keeps serving its mailboxes until EVERY warp has posted its last
        // candidate -- a sender up to a ring's depth behind could otherwise fill the mailbox of an owner that already
        // waits at the barrier (rows sorted by score make every row a candidate) and block for ever.
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            atomicAdd(done_warps, 1u);
        }
        bool all_done;
        unsigned int spins = 0;
        do {
            if (++spins == kSpinLimit) __trap();
            unsigned int dn = 0;
            if (lane == 0) dn = *(volatile unsigned int *)done_warps;  // (read BEFORE the drain it covers)
            all_done = __shfl_sync(0xffffffffu, dn, 0) >= (unsigned)kWarpsAll;
            __threadfence_block();
            for (int oq = warp; oq < QT; oq += kWarpsAll)
                mailbox_drain(&mbox[oq], &ctl[oq], lists + (size_t)oq * cap, cap, K, lane);
        } while (!all_done);
    } else {
        named_bar_sync(1, kWarpsAll * 32);
    }
    for (int qi = warp; qi < q_live && qi < QT; qi += kWarpsAll) {  // (the owner of qi flushes it)
        unsigned long long *out =
            a.partial + (((size_t)group * gridDim.x + blockIdx.x) * (size_t)a.q + (size_t)qi) * (size_t)K;
        if (WL) {
            for (int i = lane; i < K; i += 32) out[i] = lists[(size_t)qi * cap + i];
        } else {
            list_flush(&ctl[qi], lists + (size_t)qi * cap, cap, K, out, lane);
        }
    }
}

// --------------------------------------------------------------------------------------
// Generic kernel: any dims, any alignment, plain global loads.  Correct, not tuned --
// it serves the odd geometries (the reference's own unit test uses dims = 5).
// --------------------------------------------------------------------------------------
template <int QT>
__global__ void __launch_bounds__(kConsumerThreads) scan_generic_kernel(const ScanArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int dims = a.dims;
    float *qs = reinterpret_cast<float *>(smem);  // QT x dims
    size_t qbytes = ((size_t)QT * dims * 4 + 15) & ~(size_t)15;
    ListCtl *ctl = reinterpret_cast<ListCtl *>(smem + qbytes);
    unsigned long long *lists = reinterpret_cast<unsigned long long *>(ctl + QT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_wait();
    if (threadIdx.x == 0) pdl_launch_dependents();
    int q_live = a.q;
    if (a.q_count != nullptr) {
        const unsigned int c = *a.q_count;
        if (c == 0u) return;
        q_live = c < (unsigned)a.q ? (int)c : a.q;
    }
    for (int i = threadIdx.x; i < QT * dims; i += blockDim.x) {
        int qi = i / dims, e = i - qi * dims;
        const int qsrc = (qi < q_live && a.q_map != nullptr) ? a.q_map[qi] : qi;
        qs[i] = (qi < q_live) ? a.queries[(long long)qsrc * a.q_stride + (long long)e * a.l_stride] : 0.0f;
    }
    if (threadIdx.x < QT) list_init(&ctl[threadIdx.x]);
    for (int i = threadIdx.x; i < QT * a.list_cap; i += blockDim.x) lists[i] = 0ull;
    __syncthreads();

    uint32_t first, end, stride, group;
    tile_range(a, first, end, stride, group);
    const int cap = a.list_cap, K = a.K, TR = a.tile_rows;
    uint32_t segi = a.per_segment ? group : 0;

    for (uint32_t t = first; t < end; t += stride) {
        while (segi + 1 < (uint32_t)a.nseg && t >= a.segs[segi + 1].tile_start) ++segi;
        const Segment cur = a.segs[segi];
        const uint32_t row0 = (t - cur.tile_start) * TR;
        const uint32_t vrows = min((uint32_t)TR, cur.rows - row0);
        for (uint32_t r = warp; r < vrows; r += kConsumerWarps) {
            const float *x = cur.base + (size_t)(row0 + r) * dims;
            float P[QT][4];
#pragma unroll
            for (int qi = 0; qi < QT; ++qi) P[qi][0] = P[qi][1] = P[qi][2] = P[qi][3] = 0.f;
            for (int j = 0; j < dims; j += 128) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    int e = j + 4 * lane + c;
                    if (e < dims) {
                        float xv = __ldg(x + e);
#pragma unroll
                        for (int qi = 0; qi < QT; ++qi) P[qi][c] = __fmaf_rn(xv, qs[qi * dims + e], P[qi][c]);
                    }
                }
            }
#pragma unroll
            for (int qi = 0; qi < QT; ++qi) {
                float sc = canonical_reduce(P[qi][0], P[qi][1], P[qi][2], P[qi][3]);
                if (qi < q_live) {
                    unsigned long long up = a.upper ? a.upper[(size_t)group * a.q + qi] : ~0ull;
                    offer(&ctl[qi], lists + (size_t)qi * cap, cap, K, sc, cur.slot_base + row0 + r, up, lane);
                }
            }
        }
    }
    __syncthreads();
    for (int qi = warp; qi < q_live && qi < QT; qi += kConsumerWarps) {
        unsigned long long *out =
            a.partial + (((size_t)group * gridDim.x + blockIdx.x) * (size_t)a.q + (size_t)qi) * (size_t)K;
        list_flush(&ctl[qi], lists + (size_t)qi * cap, cap, K, out, lane);
    }
}

constexpr size_t kSmemLimit = 232448;  // 227 KB opt-in per CTA on sm_100

int next_pow2(int v)
{
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

template <typename Kern>
cudaError_t set_smem_attr(Kern kern, size_t bytes)
{
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int D, int QT>
cudaError_t launch_tma(const ScanPlan &plan, const ScanArgs &args, cudaStream_t stream)
{
    if constexpr (QT * D > 4096) {
        return cudaErrorInvalidValue;
    } else {
        // the register-heaviest shapes (QT * D / 128 = 32: 8 queries of 512-d, 4 of 1024-d) split the queries over
        // two groups of consumer warps -- with the owner-computes lists; the locked lists keep one group (twice the
        // warps would queue for the same locks)
        dim3 grid(plan.grid_x, plan.groups);
        if constexpr (QT * D == 4096 && QT >= 2) {
            if (plan.warp_lists) {
                cudaError_t e2 = set_smem_attr(scan_tma_kernel<D, QT, 2, true>, plan.smem_bytes);
                if (e2 != cudaSuccess) return e2;
                return launch_chain(scan_tma_kernel<D, QT, 2, true>, grid, dim3(scan_threads(2)), plan.smem_bytes, stream, args);
            }
        }
        if (plan.warp_lists) {
            cudaError_t e1 = set_smem_attr(scan_tma_kernel<D, QT, 1, true>, plan.smem_bytes);
            if (e1 != cudaSuccess) return e1;
            return launch_chain(scan_tma_kernel<D, QT, 1, true>, grid, dim3(scan_threads(1)), plan.smem_bytes, stream, args);
        }
        cudaError_t e = set_smem_attr(scan_tma_kernel<D, QT, 1, false>, plan.smem_bytes);
        if (e != cudaSuccess) return e;
        return launch_chain(scan_tma_kernel<D, QT, 1, false>, grid, dim3(scan_threads(1)), plan.smem_bytes, stream, args);
    }
}

template <int D>
cudaError_t launch_tma_q(const ScanPlan &plan, const ScanArgs &args, cudaStream_t stream)
{
    switch (plan.qt) {
        case 1: return launch_tma<D, 1>(plan, args, stream);
        case 2: return launch_tma<D, 2>(plan, args, stream);
        case 4: return launch_tma<D, 4>(plan, args, stream);
        case 8: return launch_tma<D, 8>(plan, args, stream);
    }
    return cudaErrorInvalidValue;
}

template <int QT>
cudaError_t launch_generic(const ScanPlan &plan, const ScanArgs &args, cudaStream_t stream)
{
    cudaError_t e = set_smem_attr(scan_generic_kernel<QT>, plan.smem_bytes);
    if (e != cudaSuccess) return e;
    dim3 grid(plan.grid_x, plan.groups);
    return launch_chain(scan_generic_kernel<QT>, grid, dim3(kConsumerThreads), plan.smem_bytes, stream, args);
}

}  // namespace

int scan_tile_rows(int dims)
{
    if (dims == 128 || dims == 256 || dims == 512 || dims == 1024) return kScanTileBytes / (dims * 4);
    return 0;
}

int scan_max_q(int dims)
{
    if (scan_tile_rows(dims)) return dims <= 512 ? 8 : 4;
    // generic path: queries live in shared memory
    int byq = (int)(96 * 1024 / ((size_t)dims * 4));
    if (byq >= 8) return 8;
    if (byq >= 4) return 4;
    if (byq >= 2) return 2;
    return byq >= 1 ? 1 : 0;
}

cudaError_t scan_make_plan(int device_sms, int dims, int q, int K, uint32_t total_tiles, int nseg, bool per_segment,
                           bool aligned16, ScanPlan *plan)
{
    if (q < 1 || q > kScanMaxQ || K < 1 || K > kMaxK || dims < 1) return cudaErrorInvalidValue;
    ScanPlan p{};
    p.dims = dims;
    p.qt = q <= 1 ? 1 : q <= 2 ? 2 : q <= 4 ? 4 : 8;
    p.list_cap = list_is_sorted_mode(K) ? ((K + 31) / 32) * 32 : next_pow2(2 * K);
    size_t lists_bytes = (size_t)p.qt * p.list_cap * 8;
    int tr = scan_tile_rows(dims);
    p.tma = aligned16 && tr > 0 && p.qt * dims <= 4096;
    // Owner-computes lists for the sorted lists (K <= 128) -- except the consumer-bound shapes (8 queries of 512-d, 4 of
    // 1024-d) at K <= 16: there inserts are few, and an owner warp that inserts for everybody is the warp the ring waits
    // for.  Same box, 1 M x 512, ms per pass, locked / owner-computes: 8 queries top-10 0.371 / 0.402, top-32 0.508 /
    // 0.464, top-100 1.31 / 0.69; 4 queries top-10 0.353 / 0.335; 1 query top-10 0.335 / 0.320, top-100 0.81 / 0.43.
    p.warp_lists = p.tma && list_is_sorted_mode(K) && !(p.qt * dims == 4096 && p.qt >= 2 && K <= 16);
    if (p.warp_lists) lists_bytes += (size_t)p.qt * sizeof(Mailbox) + 16;
    if (p.tma) {
        p.tile_rows = tr;
        size_t fixed = 1024;
        if (lists_bytes + fixed + 2 * (size_t)kScanTileBytes > kSmemLimit) return cudaErrorInvalidValue;
        int st = (int)((kSmemLimit - fixed - lists_bytes) / kScanTileBytes);
        p.stages = st > 7 ? 7 : st;
        p.smem_bytes = (size_t)p.stages * kScanTileBytes + fixed + lists_bytes;
    } else {
        p.tile_rows = 64;
        p.stages = 0;
        size_t qbytes = ((size_t)p.qt * dims * 4 + 15) & ~(size_t)15;
        p.smem_bytes = qbytes + 256 + lists_bytes;
        if (p.smem_bytes > kSmemLimit) return cudaErrorInvalidValue;
    }
    p.groups = per_segment ? nseg : 1;
    uint32_t per_group_tiles = per_segment ? (total_tiles + nseg - 1) / (nseg > 0 ? nseg : 1) : total_tiles;
    int want = per_segment ? (device_sms * 2 + nseg - 1) / (nseg > 0 ? nseg : 1) : device_sms;
    if (want < 1) want = 1;
    if ((uint32_t)want > per_group_tiles) want = (int)per_group_tiles;
    if (want < 1) want = 1;
    p.grid_x = want;
    *plan = p;
    return cudaSuccess;
}

cudaError_t scan_launch(const ScanPlan &plan, const Segment *d_segs, int nseg, uint32_t total_tiles,
                        const float *d_queries, int q, int64_t q_stride, int64_t l_stride, int K,
                        const unsigned long long *d_upper, unsigned long long *d_partial, cudaStream_t stream,
                        const int *d_q_map, const unsigned int *d_q_count)
{
    ScanArgs a{};
    a.q_map = d_q_map;
    a.q_count = d_q_count;
    a.segs = d_segs;
    a.nseg = nseg;
    a.total_tiles = total_tiles;
    a.queries = d_queries;
    a.q = q;
    a.q_stride = q_stride;
    a.l_stride = l_stride;
    a.K = K;
    a.list_cap = plan.list_cap;
    a.stages = plan.stages;
    a.tile_rows = plan.tile_rows;
    a.dims = plan.dims;
    a.per_segment = (plan.groups != 1) ? 1 : 0;
    a.warp_lists = plan.warp_lists ? 1 : 0;
    a.upper = d_upper;
    a.partial = d_partial;
    if (plan.tma) {
        switch (plan.dims) {
            case 128: return launch_tma_q<128>(plan, a, stream);
            case 256: return launch_tma_q<256>(plan, a, stream);
            case 512: return launch_tma_q<512>(plan, a, stream);
            case 1024: return launch_tma_q<1024>(plan, a, stream);
        }
        return cudaErrorInvalidValue;
    }
    switch (plan.qt) {
        case 1: return launch_generic<1>(plan, a, stream);
        case 2: return launch_generic<2>(plan, a, stream);
        case 4: return launch_generic<4>(plan, a, stream);
        case 8: return launch_generic<8>(plan, a, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace gf
