// gf_device.cuh -- device-side helpers shared by the sm_100a kernels.
//
// Candidate keys, the canonical float32 dot-product order, mbarrier / bulk-copy (TMA)
// PTX wrappers and the per-CTA candidate lists.  The canonical order is restated
// on the CPU in oracle/gf_oracle.c (gf_dot_canonical_scalar); the two must stay
// bit-identical.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gf {

// ------------------------------------------------------------------ programmatic dependent launch
// The search pipeline is a chain of ~9 short kernels around one long one; each is launched with
// programmatic stream serialization so that its launch latency and CTA start-up overlap the tail of its
// predecessor.  Contract: EVERY kernel of the chain executes pdl_wait() in every CTA before it reads
// anything another kernel wrote (and before it exits) -- completion is then transitive along the chain.
// Both instructions are no-ops in a kernel that was launched normally.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ------------------------------------------------------------------ keys
// (orderable image of the score) << 32 | (0xFFFFFFFF - slot).  Larger = better:
// score descending, then slot ascending.  NaN -> image 0 (worst), -0.0 -> +0.0.
// The reference's own tie order is unspecified (unstable sort.Slice, util.go:85).
__host__ __device__ __forceinline__ uint32_t score_image(float s)
{
    if (s != s) return 0u;
    s = s + 0.0f;
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(s);
#else
    union { float f; uint32_t u; } c; c.f = s; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float image_score(uint32_t im)
{
    uint32_t u = (im & 0x80000000u) ? (im & 0x7FFFFFFFu) : ~im;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float s, uint32_t slot)
{
    return ((uint64_t)score_image(s) << 32) | (uint64_t)(0xFFFFFFFFu - slot);
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return image_score((uint32_t)(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_slot(uint64_t k) { return 0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFu); }

#ifdef __CUDACC__
// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP); completion is
// signalled as `bytes` transaction bytes on `bar`.  src/dst 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// same, with an L2 eviction-priority policy (createpolicy) for stream-once data
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                              uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------ canonical reduction
// Thread l of a warp holds P[4l..4l+3] in (p0,p1,p2,p3); returns the full sum in every lane:
//   T[l] = (P[4l]+P[4l+1]) + (P[4l+2]+P[4l+3]); then xor-butterfly 16,8,4,2,1; then +0.0f.
__device__ __forceinline__ float canonical_reduce(float p0, float p1, float p2, float p3)
{
    float t = __fadd_rn(__fadd_rn(p0, p1), __fadd_rn(p2, p3));
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, off));
    return __fadd_rn(t, 0.0f);
}

// Same sum for V independent values at once (V a power of two <= 32): at each of the first
// log2(V) butterfly levels a lane keeps half of its values and trades the other half, so the
// pairs added are exactly those of canonical_reduce (T[l] + T[l ^ off]) but V values cost
// V-1 + (5 - log2 V) shuffles instead of 5 V.  On return lane l holds the finished sum of value
// index  l >> (5 - log2 V)  (input t[] holds the per-thread sums T of every value).
template <int V>
__device__ __forceinline__ float canonical_reduce_many(float (&t)[V], int lane)
{
    static_assert(V >= 1 && V <= 32 && (V & (V - 1)) == 0, "V must be a power of two");
    int off = 16;
#pragma unroll
    for (int n = V; n > 1; n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = up ? t[i] : t[i + n / 2];
            const float keep = up ? t[i + n / 2] : t[i];
            t[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, off));
        }
        off >>= 1;
    }
    float s = t[0];
#pragma unroll
    for (; off >= 1; off >>= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, off));
    return __fadd_rn(s, 0.0f);
}

// ------------------------------------------------------------------ per-CTA candidate lists
// One list per query lives in shared memory: an unsorted append buffer of `cap` keys
// (a power of two, >= 2*K) plus the gate tau = K-th best key seen at the last compaction
// (0 while fewer than `cap` keys have ever been appended).  A warp that finds key > tau
// takes the list's lock, appends, and when the buffer is full sorts it and keeps K.
struct ListCtl {
    unsigned long long tau;  // gate; keys <= tau are rejected without taking the lock
    int cnt;                 // keys in the buffer
    int lock;                // 0 free, 1 held
    float tauf;              // score of tau (-inf while tau == 0): the one-compare pre-gate
    int pad;
};

__device__ __forceinline__ void list_init(ListCtl *ctl)
{
    ctl->tau = 0ull;
    ctl->cnt = 0;
    ctl->lock = 0;
    ctl->tauf = __int_as_float(0xff800000);  // -inf
    ctl->pad = 0;
}

// bitonic sort (descending) of n (power of two) keys in shared memory by ONE warp
__device__ __forceinline__ void warp_sort_desc(unsigned long long *buf, int n, int lane)
{
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < n; i += 32) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = buf[i], b = buf[ixj];
                    bool desc = ((i & k) == 0);
                    if (desc ? (a < b) : (a > b)) {
                        buf[i] = b;
                        buf[ixj] = a;
                    }
                }
            }
            __syncwarp();
        }
    }
}

// K <= kSortedMaxK: the list is kept SORTED (descending, zero padded) in cap = 32*E slots, lane l
// owning slots [l*E, l*E+E), so tau is exact after every insert and the number of inserts per
// list is the minimum a streaming top-K can have, K*(1 + ln(rows/K)).  Larger K: unsorted append
// buffer of cap >= 2K keys, sorted and cut back to K whenever it fills (amortised O(log^2) per key).
constexpr int kSortedMaxK = 128;

__host__ __device__ __forceinline__ bool list_is_sorted_mode(int K) { return K <= kSortedMaxK; }

__device__ __forceinline__ float tau_score(unsigned long long tau)
{
    return tau == 0ull ? __int_as_float(0xff800000) : key_score(tau);
}

// lock held by this warp, key > tau already checked
template <int E>
__device__ __forceinline__ void sorted_insert(ListCtl *ctl, unsigned long long *buf, int K, unsigned long long key,
                                              int lane)
{
    unsigned long long my[E];
    unsigned p = 0;  // slots holding a key greater than `key`: the insert position (the list is a sorted prefix)
#pragma unroll
    for (int e = 0; e < E; ++e) {
        my[e] = *(volatile unsigned long long *)&buf[lane * E + e];
        p += __popc(__ballot_sync(0xffffffffu, my[e] > key));
    }
    const unsigned long long up = __shfl_up_sync(0xffffffffu, my[E - 1], 1);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const unsigned i = (unsigned)(lane * E + e);
        const unsigned long long prev = (e > 0) ? my[e > 0 ? e - 1 : 0] : up;
        const unsigned long long nv = (i < p) ? my[e] : ((i == p) ? key : prev);
        if (i >= p) *(volatile unsigned long long *)&buf[i] = nv;
        if (i == (unsigned)(K - 1)) {
            *(volatile unsigned long long *)&ctl->tau = nv;
            *(volatile float *)&ctl->tauf = tau_score(nv);
        }
    }
}

// Whole warp calls this with a warp-uniform key.
__device__ __forceinline__ void list_insert(ListCtl *ctl, unsigned long long *buf, int cap, int K,
                                            unsigned long long key, int lane)
{
    if (lane == 0) {
        // back off while another warp holds the list: spinning ATOMS on one bank would starve the
        // holder's own shared-memory traffic
        unsigned ns = 32;
        while (atomicCAS(&ctl->lock, 0, 1) != 0) {
            __nanosleep(ns);
            if (ns < 512) ns <<= 1;
        }
    }
    __syncwarp();
    __threadfence_block();
    unsigned long long tau = *(volatile unsigned long long *)&ctl->tau;
    if (key > tau) {
        if (list_is_sorted_mode(K)) {
            switch (cap >> 5) {
                case 1: sorted_insert<1>(ctl, buf, K, key, lane); break;
                case 2: sorted_insert<2>(ctl, buf, K, key, lane); break;
                case 3: sorted_insert<3>(ctl, buf, K, key, lane); break;
                default: sorted_insert<4>(ctl, buf, K, key, lane); break;
            }
        } else {
            int n = *(volatile int *)&ctl->cnt;
            if (lane == 0) buf[n] = key;
            n += 1;
            __syncwarp();
            if (n == cap) {
                warp_sort_desc(buf, cap, lane);
                n = K;
                if (lane == 0) {
                    unsigned long long kth = buf[K - 1];
                    *(volatile unsigned long long *)&ctl->tau = kth;
                    *(volatile float *)&ctl->tauf = tau_score(kth);
                }
            }
            if (lane == 0) *(volatile int *)&ctl->cnt = n;
        }
    }
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        atomicExch(&ctl->lock, 0);
    }
    __syncwarp();
}

// Owner-computes lists (scan_tma_kernel, K <= 128).  The shared lists above serialise the warps of a CTA behind one lock
// per query (atomicCAS + back-off + two fences per insert), and a consumer warp that is late releasing its stage stalls
// the whole bulk-copy ring: measured on 1 M x 512 rows the gate + insert path cost 11 % of the scan's bandwidth at one
// query and 23 % at eight (6.83 -> 6.11 and 6.69 -> 5.16 TB/s), nearly all of it in the first tiles, where every row
// still passes and seven warps queue for the same lock.  Here every query's list (sorted, K <= 128: sorted_insert's layout)
// belongs to ONE consumer warp, which alone reads and writes it -- no lock, no fence; the other warps post their rare
// candidates to the owner's mailbox (one shared-memory atomic + one store) and go on streaming, the owner drains its
// mailboxes once per tile.  The gate is the owner's exact K-th key, at most a tile old.
constexpr int kMailboxCap = 64;
struct Mailbox {
    unsigned int tail;  // next index a sender takes (atomicAdd)
    unsigned int head;  // next index the owner consumes (written by the owner only)
    unsigned long long slot[kMailboxCap];  // 0 = empty
};

// owner only: sorted insert into its list (cap = 32 E slots, lane l owns slots [lE, lE + E)); ctl->tau / tauf = the new
// K-th key and its score, which the other warps read as their gate
__device__ __forceinline__ void owner_list_insert(ListCtl *ctl, unsigned long long *buf, int cap, int K,
                                                  unsigned long long key, int lane)
{
    if (!(key > *(volatile unsigned long long *)&ctl->tau)) return;  // (warp-uniform: only this warp writes tau)
    switch (cap >> 5) {
        case 1: sorted_insert<1>(ctl, buf, K, key, lane); break;
        case 2: sorted_insert<2>(ctl, buf, K, key, lane); break;
        case 3: sorted_insert<3>(ctl, buf, K, key, lane); break;
        default: sorted_insert<4>(ctl, buf, K, key, lane); break;
    }
}

// owner only: consume everything posted so far
__device__ __forceinline__ void mailbox_drain(Mailbox *mb, ListCtl *ctl, unsigned long long *buf, int cap, int K, int lane)
{
    // (every value other warps change is read by ONE lane and broadcast: lanes that saw different values would leave the
    // warp-collective code below at different times)
    unsigned int t = 0, h = 0;
    if (lane == 0) {
        t = *(volatile unsigned int *)&mb->tail;
        h = *(volatile unsigned int *)&mb->head;
    }
    t = __shfl_sync(0xffffffffu, t, 0);
    h = __shfl_sync(0xffffffffu, h, 0);
    while (h != t) {
        // An index that was taken but not stored yet ends this drain instead of being waited for: its sender may itself
        // sit in a full-mailbox wait, draining ITS mailbox and coming across a slot that this warp owes -- two owners
        // spinning on each other's unwritten slots never get back to the stores they owe (the hang of the first
        // version on rows sorted by score).  Every caller comes back: the tile wait, the full-mailbox wait and the end
        // of the kernel are loops around this function.
        unsigned long long key = 0ull;
        if (lane == 0) key = *(volatile unsigned long long *)&mb->slot[h % kMailboxCap];
        key = __shfl_sync(0xffffffffu, key, 0);
        if (key == 0ull) return;
        if (lane == 0) {
            *(volatile unsigned long long *)&mb->slot[h % kMailboxCap] = 0ull;
            __threadfence_block();
            *(volatile unsigned int *)&mb->head = h + 1;
        }
        ++h;
        owner_list_insert(ctl, buf, cap, K, key, lane);
    }
}

// Write the list's best K keys (descending, zero padded) to out[0..K).  One warp, no contention.
__device__ __forceinline__ void list_flush(ListCtl *ctl, unsigned long long *buf, int cap, int K,
                                           unsigned long long *out, int lane)
{
    if (!list_is_sorted_mode(K)) {
        int n = ctl->cnt;
        for (int i = n + lane; i < cap; i += 32) buf[i] = 0ull;
        __syncwarp();
        warp_sort_desc(buf, cap, lane);
    }
    for (int i = lane; i < K; i += 32) out[i] = buf[i];
}
#endif  // __CUDACC__

}  // namespace gf
