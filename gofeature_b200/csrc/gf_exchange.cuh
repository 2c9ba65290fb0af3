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

// The whole CTA (blockDim.x >= world): push the K winners `local` of query qi (shared or global memory) and the
// certificate flag into every rank's inbox over NVLink peer memory, wait until the `world` words of this query
// carry this step's sequence number, fold the world x K keys.  buf: kMergeBuf keys of shared memory; s_flag: one
// shared word.  Inbox (per rank, two slots alternating with the sequence parity so a fast peer can run one step
// ahead):  keys [2][world][max_q][max_k] u64      words [2][world][max_q] u64 = seq << 1 | uncertified-flag
__device__ __forceinline__ void exchange_fold(const ExchangeArgs &a, int qi, const unsigned long long *local,
                                              unsigned int local_flag, unsigned long long *buf, unsigned int *s_flag,
                                              unsigned long long *out_keys, unsigned int *out_flag)
{
    const int K = a.K, world = a.world;
    const int slot = (int)(a.seq & 1ull);
    const size_t keys_per_slot = (size_t)world * a.max_q * a.max_k;
    const size_t words_off = 2 * keys_per_slot;  // in u64 units
    if (a.gather_root_plus1 > 0) {
        // ---- gather to one rank: K stores + one word per query instead of world x (K + 1), one waiter
        const int root = a.gather_root_plus1 - 1;
        unsigned long long *rbase = reinterpret_cast<unsigned long long *>(a.peers[root]);
        for (int j = threadIdx.x; j < K; j += blockDim.x)
            rbase[(size_t)slot * keys_per_slot + ((size_t)a.rank * a.max_q + qi) * a.max_k + j] = local[j];
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            volatile unsigned long long *w = rbase + words_off + ((size_t)slot * world + a.rank) * a.max_q + qi;
            *w = (a.seq << 1) | (local_flag ? 1ull : 0ull);
        }
        if (a.rank != root) return;
        if (threadIdx.x == 0) *s_flag = 0u;
        __syncthreads();
        if ((int)threadIdx.x < world) {
            volatile unsigned long long *w = rbase + words_off + ((size_t)slot * world + threadIdx.x) * a.max_q + qi;
            const unsigned long long t0 = globaltimer_ns();
            unsigned long long word;
            unsigned int spins = 0;
            while (((word = *w) >> 1) != a.seq) {
                if ((++spins & 1023u) == 0u && globaltimer_ns() - t0 > a.timeout_ns) {
                    atomicOr(s_flag, 0xFFFFFFFFu);
                    break;
                }
            }
            if (word & 1ull) atomicOr(s_flag, 1u);
            *w = 0ull;  // consumed: the next step may carry the same sequence number
        }
        __threadfence_system();
        __syncthreads();
        const unsigned long long *mine = rbase + (size_t)slot * keys_per_slot;
        const int n = world * K;
        int np = 2;
        while (np < n) np <<= 1;
        for (int i = threadIdx.x; i < np; i += blockDim.x) {
            unsigned long long v = 0ull;
            if (i < n) {
                const int r = i / K, j = i - r * K;
                v = __ldcv(mine + ((size_t)r * a.max_q + qi) * a.max_k + j);
            }
            buf[i] = v;
        }
        __syncthreads();
        block_sort_desc(buf, np);
        for (int i = threadIdx.x; i < K; i += blockDim.x) out_keys[(size_t)qi * K + i] = buf[i];
        if (out_flag != nullptr && threadIdx.x == 0) out_flag[qi] = *s_flag;
        return;
    }
    // ---- push my K keys of this query into every inbox (mine included)
    for (int i = threadIdx.x; i < world * K; i += blockDim.x) {
        const int p = i / K, j = i - p * K;
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(a.peers[p]) + (size_t)slot * keys_per_slot +
                                  ((size_t)a.rank * a.max_q + qi) * a.max_k + j;
        *dst = local[j];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const unsigned long long word = (a.seq << 1) | (local_flag ? 1ull : 0ull);
        volatile unsigned long long *w = reinterpret_cast<unsigned long long *>(a.peers[threadIdx.x]) + words_off +
                                         ((size_t)slot * world + a.rank) * a.max_q + qi;
        *w = word;
    }
    // ---- wait for everybody's keys of this query
    if (threadIdx.x == 0) *s_flag = 0u;
    __syncthreads();
    if ((int)threadIdx.x < world) {
        volatile unsigned long long *w = reinterpret_cast<unsigned long long *>(a.peers[a.rank]) + words_off +
                                         ((size_t)slot * world + threadIdx.x) * a.max_q + qi;
        const unsigned long long t0 = globaltimer_ns();
        unsigned long long word;
        unsigned int spins = 0;
        while (((word = *w) >> 1) != a.seq) {
            if ((++spins & 1023u) == 0u && globaltimer_ns() - t0 > a.timeout_ns) {  // a peer died: do not hang the GPU
                atomicOr(s_flag, 0xFFFFFFFFu);
                break;
            }
        }
        if (word & 1ull) atomicOr(s_flag, 1u);
    }
    __threadfence_system();
    __syncthreads();
    // ---- fold world x K keys (world * K <= kMergeBuf, checked by the host)
    const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(a.peers[a.rank]) +
                                     (size_t)slot * keys_per_slot;
    const int n = world * K;
    int np = 2;
    while (np < n) np <<= 1;
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
        unsigned long long v = 0ull;
        if (i < n) {
            const int r = i / K, j = i - r * K;
            v = __ldcv(mine + ((size_t)r * a.max_q + qi) * a.max_k + j);
        }
        buf[i] = v;
    }
    __syncthreads();
    block_sort_desc(buf, np);
    for (int i = threadIdx.x; i < K; i += blockDim.x) out_keys[(size_t)qi * K + i] = buf[i];
    if (out_flag != nullptr && threadIdx.x == 0) out_flag[qi] = *s_flag;
}

}  // namespace gf
