// gf_exchange.cuh -- the cross-rank step of a sharded search as a CTA-wide device function, shared by
// exchange_merge_kernel (gf_merge.cu) and the one-pass chain's tail (gf_gemm.cu), which ends with it.
#pragma once
#include "gf_device.cuh"
#include "gf_kernels.h"

namespace gf {

constexpr int kMergeBuf = 2048;  // keys held in shared memory; kMaxK <= kMergeBuf/2

// bitonic sort (descending) of n (power of two, <= kMergeBuf) keys in shared memory by the CTA
__device__ __forceinline__ void block_sort_desc(unsigned long long *buf, int n)
{
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
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
            __syncthreads();
        }
    }
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// The whole CTA (blockDim.x >= world): hand the K winners `local` of query qi (shared or global memory) and the
// certificate flag to every rank (or to the gather root only) over NVLink peer memory, wait for everybody's, fold
// the world x K keys.  buf: kMergeBuf keys of shared memory; s_flag: one shared word.
//
// Protocol: every 8-byte word that crosses the link validates ITSELF -- its upper half is the step's 32-bit sequence
// tag, its lower half the payload (a key travels as two words: score image, slot image; the flag as one).  An
// aligned 8-byte store is atomic, so the receiver simply polls its own inbox until a word carries this step's tag:
// no fence between data and flag, no second hop (the fence.sys behind the remote stores alone cost ~6.8 us of a
// 21 us exchange at 2 GPUs; measured with `trace`).  Inbox (per rank), two slots alternating with the sequence
// parity so a fast peer can run one step ahead:
//   keys  [2][world][max_q][max_k][2] u64      flags [2][world][max_q] u64
// Gather mode (a.gather_root_plus1 = r + 1): only rank r receives, folds and writes out_keys / out_flag; it zeroes
// every word it consumed, so the next step may carry the same tag (its arguments are baked into a recorded graph).
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void exchange_fold(const ExchangeArgs &a, int qi, const unsigned long long *local,
                                              unsigned int local_flag, unsigned long long *buf, unsigned int *s_flag,
                                              unsigned long long *out_keys, unsigned int *out_flag)
{
    const int K = a.K, world = a.world;
    const int slot = (int)(a.seq & 1ull);
    const unsigned long long tag = (a.seq & 0xFFFFFFFFull) << 32;  // (the host never hands out a sequence number whose low word is 0)
    const size_t keys_per_slot = (size_t)world * a.max_q * a.max_k * 2;  // u64 words
    const size_t flags_off = 2 * keys_per_slot;
    const int root = a.gather_root_plus1 - 1;  // -1: all to all
    const bool gather = root >= 0;
    if (a.trace != nullptr && threadIdx.x == 0) a.trace[qi * 4 + 0] = globaltimer_ns();
    // ---- push: my K keys of this query (and the flag) into every receiving inbox, mine included
    const int ntarget = gather ? 1 : world;
    for (int i = threadIdx.x; i < ntarget * K; i += blockDim.x) {
        const int t = i / K, j = i - t * K;
        const int p = gather ? root : t;
        const unsigned long long key = local[j];
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(a.peers[p]) + (size_t)slot * keys_per_slot +
                                  (((size_t)a.rank * a.max_q + qi) * a.max_k + j) * 2;
        st_volatile_u64(dst, tag | (key >> 32));
        st_volatile_u64(dst + 1, tag | (key & 0xFFFFFFFFull));
    }
    if ((int)threadIdx.x < ntarget) {
        const int p = gather ? root : (int)threadIdx.x;
        unsigned long long *w = reinterpret_cast<unsigned long long *>(a.peers[p]) + flags_off +
                                ((size_t)slot * world + a.rank) * a.max_q + qi;
        st_volatile_u64(w, tag | (local_flag ? 1ull : 0ull));
    }
    if (a.trace != nullptr && threadIdx.x == 0) a.trace[qi * 4 + 1] = globaltimer_ns();
    if (gather && a.rank != root) return;
    // ---- receive: poll my own inbox until every word of this query carries the tag
    if (threadIdx.x == 0) *s_flag = 0u;
    __syncthreads();
    unsigned long long *mine = reinterpret_cast<unsigned long long *>(a.peers[a.rank]);
    const int n = world * K;
    int np = 2;
    while (np < n) np <<= 1;
    const unsigned long long t0 = globaltimer_ns();
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
        unsigned long long v = 0ull;
        if (i < n) {
            const int r = i / K, j = i - r * K;
            unsigned long long *src = mine + (size_t)slot * keys_per_slot + (((size_t)r * a.max_q + qi) * a.max_k + j) * 2;
            unsigned long long w0, w1;
            unsigned int spins = 0;
            for (;;) {
                w0 = ld_volatile_u64(src);
                w1 = ld_volatile_u64(src + 1);
                if ((w0 & 0xFFFFFFFF00000000ull) == tag && (w1 & 0xFFFFFFFF00000000ull) == tag) break;
                if ((++spins & 1023u) == 0u && globaltimer_ns() - t0 > a.timeout_ns) {  // a peer died: do not hang the GPU
                    atomicOr(s_flag, 0xFFFFFFFFu);
                    w0 = w1 = 0ull;
                    break;
                }
            }
            v = (w0 << 32) | (w1 & 0xFFFFFFFFull);
            if (gather) {  // consumed
                st_volatile_u64(src, 0ull);
                st_volatile_u64(src + 1, 0ull);
            }
        }
        buf[i] = v;
    }
    if ((int)threadIdx.x < world) {
        unsigned long long *w = mine + flags_off + ((size_t)slot * world + threadIdx.x) * a.max_q + qi;
        unsigned long long word;
        unsigned int spins = 0;
        for (;;) {
            word = ld_volatile_u64(w);
            if ((word & 0xFFFFFFFF00000000ull) == tag) break;
            if ((++spins & 1023u) == 0u && globaltimer_ns() - t0 > a.timeout_ns) {
                atomicOr(s_flag, 0xFFFFFFFFu);
                word = 0ull;
                break;
            }
        }
        if (word & 1ull) atomicOr(s_flag, 1u);
        if (gather) st_volatile_u64(w, 0ull);
    }
    __syncthreads();
    if (a.trace != nullptr && threadIdx.x == 0) a.trace[qi * 4 + 2] = globaltimer_ns();
    // ---- fold world x K keys (world * K <= kMergeBuf, checked by the host)
    if (n <= 256 && np <= (int)blockDim.x) {
        // few keys: every key counts the keys above it (keys are unique: one per slot) -- one barrier instead of the
        // ~15 of a bitonic sort
        const int i = threadIdx.x;
        const unsigned long long my = i < n ? buf[i] : 0ull;
        int rk = 0;
        if (my != 0ull)
            for (int j = 0; j < n; ++j) rk += buf[j] > my ? 1 : 0;
        const int nz = __syncthreads_count(my != 0ull);
        if (my != 0ull && rk < K) out_keys[(size_t)qi * K + rk] = my;
        if (i < K && i >= nz) out_keys[(size_t)qi * K + i] = 0ull;
    } else {
        block_sort_desc(buf, np);
        for (int i = threadIdx.x; i < K; i += blockDim.x) out_keys[(size_t)qi * K + i] = buf[i];
    }
    if (out_flag != nullptr && threadIdx.x == 0) out_flag[qi] = *s_flag;
    if (a.trace != nullptr && threadIdx.x == 0) a.trace[qi * 4 + 3] = globaltimer_ns();
}

}  // namespace gf
