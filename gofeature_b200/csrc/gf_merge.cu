// gf_merge.cu -- candidate-list merge and the synthetic row generator, sm_100a.
//
// merge_kernel folds `lists` per-CTA (or per-rank) candidate lists of K keys into the
// top-K per query: the device-side counterpart of MaxNFeatureResult (util.go:95-112)
// as used at set.go:154-157, on (score,slot) keys instead of Go structs.
#include <cuda_bf16.h>
#include "gf_device.cuh"
#include "gf_exchange.cuh"
#include "gf_kernels.h"

namespace gf {

namespace {

constexpr int kMergeThreads = 256;
__device__ __forceinline__ int pow2_at_least(int v)
{
    int p = 2;
    while (p < v) p <<= 1;
    return p;
}

// grid (q, groups); d_in [groups][lists][q][K]; d_out [groups][q][K]
__global__ void __launch_bounds__(kMergeThreads) merge_kernel(const unsigned long long *__restrict__ d_in, int lists,
                                                             int q, int K, unsigned long long *__restrict__ d_out,
                                                             const unsigned int *__restrict__ d_q_count,
                                                             const int *__restrict__ d_out_map)
{
    pdl_wait();  // a link of the search chain (gf_device.cuh)
    if (threadIdx.x == 0) pdl_launch_dependents();
    if (d_q_count != nullptr && (unsigned)blockIdx.x >= *d_q_count) return;  // fix-up pass with fewer live queries
    __shared__ unsigned long long buf[kMergeBuf];
    __shared__ unsigned long long red[kMergeThreads / 32];
    __shared__ unsigned long long s_tau;
    __shared__ int s_n;

    const int qi = blockIdx.x, group = blockIdx.y;
    const unsigned long long *base = d_in + ((size_t)group * lists * q + qi) * (size_t)K;
    const size_t list_stride = (size_t)q * K;

    // tau = best K-th key of any single list: that list alone holds K keys >= tau
    unsigned long long m = 0ull;
    for (int l = threadIdx.x; l < lists; l += blockDim.x) {
        unsigned long long v = base[(size_t)l * list_stride + (K - 1)];
        m = v > m ? v : m;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, m, off);
        m = o > m ? o : m;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0ull;
        for (int w = 0; w < kMergeThreads / 32; ++w) t = red[w] > t ? red[w] : t;
        s_tau = t;
    }
    __syncthreads();

    const long long total = (long long)lists * K;
    const int chunk = kMergeBuf - K;  // with <= K keys kept, one chunk can never overflow buf
    for (long long c0 = 0; c0 < total; c0 += chunk) {
        const unsigned long long tau = s_tau;
        long long c1 = c0 + chunk < total ? c0 + chunk : total;
        for (long long i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
            long long l = i / K;
            int j = (int)(i - l * K);
            unsigned long long key = base[(size_t)l * list_stride + j];
            if (key != 0ull && key >= tau) {
                int pos = atomicAdd(&s_n, 1);
                buf[pos] = key;
            }
        }
        __syncthreads();
        int n = s_n;
        if (c1 < total && n > K) {
            int np = pow2_at_least(n);
            for (int i = n + threadIdx.x; i < np; i += blockDim.x) buf[i] = 0ull;
            __syncthreads();
            block_sort_desc(buf, np);
            if (threadIdx.x == 0) {
                s_n = K;
                if (buf[K - 1] > s_tau) s_tau = buf[K - 1];
            }
            __syncthreads();
        }
    }
    int n = s_n;
    int np = pow2_at_least(n > K ? n : K);
    for (int i = n + threadIdx.x; i < np; i += blockDim.x) buf[i] = 0ull;
    __syncthreads();
    block_sort_desc(buf, np);
    // fix-up pass: query qi of this launch is row d_out_map[qi] of the caller's result
    unsigned long long *out = d_out + ((size_t)group * q + (d_out_map != nullptr ? d_out_map[qi] : qi)) * (size_t)K;
    for (int i = threadIdx.x; i < K; i += blockDim.x) out[i] = buf[i];
}

// ------------------------------------------------------------------ synthetic rows
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ float uniform_pm1(unsigned long long seed, unsigned long long counter)
{
    unsigned long long h = splitmix64(seed ^ splitmix64(counter));
    unsigned int m = (unsigned int)(h >> 40);
    return __fadd_rn(__fmul_rn((float)m, 1.0f / 8388608.0f), -1.0f);
}

// the source ordinal of row `ord` (itself, or an earlier row it duplicates) and its value at element e
__device__ __forceinline__ unsigned long long synth_source(const SynthSpec &sp, unsigned long long ord)
{
    if (sp.dup_ppm != 0u && ord > 0ull &&
        splitmix64(sp.seed ^ 0xD0B1E5ull ^ splitmix64(ord)) % 1000000ull < (unsigned long long)sp.dup_ppm)
        return splitmix64(sp.seed ^ 0x5EED0FD0ull ^ splitmix64(ord)) % ord;
    return ord;
}
__device__ __forceinline__ float synth_value(const SynthSpec &sp, unsigned long long src, unsigned long long centre,
                                             int dims, int e)
{
    const float nv = uniform_pm1(sp.seed, src * (unsigned long long)dims + (unsigned long long)e);
    if (sp.centres == 0ull) return nv;
    const float cv = uniform_pm1(sp.seed ^ 0xC3A7E25ull, centre * (unsigned long long)dims + (unsigned long long)e);
    return __fmaf_rn(sp.spread, nv, cv);
}

// one warp per row; element e handled by lane (e%128)/4, partial sum e%4 (canonical order)
__global__ void __launch_bounds__(256) synth_fill_kernel(float *__restrict__ rows, long long nrows, int dims,
                                                         const SynthSpec sp, long long first_row, int normalize)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < nrows; r += nwarps) {
        const unsigned long long src = synth_source(sp, (unsigned long long)(first_row + r));
        const unsigned long long centre = sp.centres ? splitmix64(sp.seed ^ 0xC1A55E5ull ^ splitmix64(src)) % sp.centres : 0ull;
        float inv_ok = 1.0f, nrm = 1.0f;
        if (normalize) {
            float p[4] = {0.f, 0.f, 0.f, 0.f};
            for (int j = 0; j < dims; j += 128) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    int e = j + 4 * lane + c;
                    if (e < dims) {
                        float v = synth_value(sp, src, centre, dims, e);
                        p[c] = __fmaf_rn(v, v, p[c]);
                    }
                }
            }
            float ss = canonical_reduce(p[0], p[1], p[2], p[3]);
            nrm = __fsqrt_rn(ss);
            inv_ok = (nrm == 0.0f) ? 0.0f : 1.0f;
        }
        float *out = rows + (size_t)r * dims;
        for (int j = 0; j < dims; j += 128) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int e = j + 4 * lane + c;
                if (e < dims) {
                    float v = synth_value(sp, src, centre, dims, e);
                    if (normalize) v = (inv_ok == 0.0f) ? 0.0f : __fdiv_rn(v, nrm);
                    out[e] = v;
                }
            }
        }
    }
}

}  // namespace

// max |row|^2 over a contiguous range of rows (incremental upkeep of the certificate's norm bound)
namespace {
__global__ void rownorm_range_kernel(const float *__restrict__ base, long long rows, int dims, unsigned int *out_bits)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float best = 0.f;
    for (long long r = warp; r < rows; r += nwarps) {
        const float *x = base + (size_t)r * dims;
        float acc = 0.f;
        for (int e = lane; e < dims; e += 32) {
            float v = x[e];
            acc += v * v;
        }
        for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (acc == acc) best = fmaxf(best, acc);
    }
    if (lane == 0 && best > 0.f) atomicMax(out_bits, __float_as_uint(best * 1.0001f));
}
}  // namespace

cudaError_t rownorm_range_launch(const float *d_rows, int64_t rows, int dims, unsigned int *d_out_bits,
                                 cudaStream_t stream)
{
    if (rows <= 0) return cudaSuccess;
    long long blocks = (rows + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    rownorm_range_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_rows, rows, dims, d_out_bits);
    return cudaGetLastError();
}

// fp32 rows -> bf16 shadow rows (round to nearest even) + the two bounds the bf16 certificate uses.
// One warp per row, 8 bytes of bf16 per lane per step (dims % 4 == 0 on the shadow path).
namespace {
__global__ void shadow_rows_kernel(const float *__restrict__ base, long long rows, int dims,
                                   __nv_bfloat16 *__restrict__ out, unsigned int *small)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float best = 0.f, best_res = 0.f;
    for (long long r = warp; r < rows; r += nwarps) {
        const float4 *x = reinterpret_cast<const float4 *>(base + (size_t)r * dims);
        uint2 *o = reinterpret_cast<uint2 *>(out + (size_t)r * dims);
        float acc = 0.f, res = 0.f;
        for (int e = lane; e < dims / 4; e += 32) {
            const float4 v = x[e];
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 w;
            w.x = *reinterpret_cast<const unsigned int *>(&lo);
            w.y = *reinterpret_cast<const unsigned int *>(&hi);
            o[e] = w;
            acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            const float d0 = __low2float(lo) - v.x, d1 = __high2float(lo) - v.y, d2 = __low2float(hi) - v.z,
                        d3 = __high2float(hi) - v.w;
            res += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
        for (int off = 16; off >= 1; off >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, off);
            res += __shfl_xor_sync(0xffffffffu, res, off);
        }
        if (acc == acc) best = fmaxf(best, acc);
        // a residual that is not finite (row element beyond the bf16 range, Inf, NaN) must fail every certificate
        best_res = (res == res) ? fmaxf(best_res, res) : __int_as_float(0x7f800000);
    }
    if (lane == 0) {
        if (best > 0.f) atomicMax(small + 0, __float_as_uint(best * 1.0001f));
        if (best_res > 0.f) atomicMax(small + 2, __float_as_uint(best_res * 1.0001f));
    }
}
}  // namespace

namespace {
// int8 shadow: one warp per row, the row stays in registers between the max pass and the rounding pass
__global__ void shadow_rows_i8_kernel(const float *__restrict__ base, long long rows, int dims,
                                      signed char *__restrict__ out, float *__restrict__ scales, unsigned int *small)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float best = 0.f, best_res = 0.f;
    for (long long r = warp; r < rows; r += nwarps) {
        const float4 *x = reinterpret_cast<const float4 *>(base + (size_t)r * dims);
        float acc = 0.f, amax = 0.f;
        bool bad = false;
        for (int e = lane; e < dims / 4; e += 32) {
            const float4 v = x[e];
            acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
        }
        for (int off = 16; off >= 1; off >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, off);
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
        }
        const float sc = amax / 127.f, inv = sc > 0.f ? 1.f / sc : 0.f;
        if (!(acc == acc) || !(amax < __int_as_float(0x7f800000)) || !(inv < __int_as_float(0x7f800000))) bad = true;
        float res = 0.f;
        unsigned int *o = reinterpret_cast<unsigned int *>(out + (size_t)r * dims);
        for (int e = lane; e < dims / 4; e += 32) {
            const float4 v = x[e];  // second read: L1/L2 hit
            int q0 = __float2int_rn(v.x * inv), q1 = __float2int_rn(v.y * inv), q2 = __float2int_rn(v.z * inv),
                q3 = __float2int_rn(v.w * inv);
            q0 = max(-127, min(127, q0));
            q1 = max(-127, min(127, q1));
            q2 = max(-127, min(127, q2));
            q3 = max(-127, min(127, q3));
            if (bad) q0 = q1 = q2 = q3 = 0;
            o[e] = (unsigned)(q0 & 255) | ((unsigned)(q1 & 255) << 8) | ((unsigned)(q2 & 255) << 16) |
                   ((unsigned)(q3 & 255) << 24);
            const float d0 = (float)q0 * sc - v.x, d1 = (float)q1 * sc - v.y, d2 = (float)q2 * sc - v.z,
                        d3 = (float)q3 * sc - v.w;
            res += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
        for (int off = 16; off >= 1; off >>= 1) res += __shfl_xor_sync(0xffffffffu, res, off);
        if (lane == 0) scales[r] = bad ? 0.f : sc;
        if (acc == acc) best = fmaxf(best, acc);
        // Inf / NaN / overflowing rows cannot be bounded: every certificate of this set must fail
        best_res = (!bad && res == res) ? fmaxf(best_res, res) : __int_as_float(0x7f800000);
    }
    if (lane == 0) {
        if (best > 0.f) atomicMax(small + 0, __float_as_uint(best * 1.0001f));
        if (best_res > 0.f) atomicMax(small + 2, __float_as_uint(best_res * 1.0001f));
    }
}
}  // namespace

cudaError_t shadow_rows_i8_launch(const float *d_rows, int64_t rows, int dims, void *d_shadow_rows, float *d_scales,
                                  unsigned int *d_small, cudaStream_t stream)
{
    if (rows <= 0) return cudaSuccess;
    if (dims % 4 != 0) return cudaErrorInvalidValue;
    long long blocks = (rows + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    shadow_rows_i8_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_rows, rows, dims, (signed char *)d_shadow_rows, d_scales,
                                                                d_small);
    return cudaGetLastError();
}

cudaError_t shadow_rows_launch(const float *d_rows, int64_t rows, int dims, void *d_shadow_rows, unsigned int *d_small,
                               cudaStream_t stream)
{
    if (rows <= 0) return cudaSuccess;
    if (dims % 4 != 0) return cudaErrorInvalidValue;
    long long blocks = (rows + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    shadow_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_rows, rows, dims, (__nv_bfloat16 *)d_shadow_rows, d_small);
    return cudaGetLastError();
}

// Mutation helpers that take their (small) argument lists straight from the pinned upload ring: a Delete of
// 10 000 ids or an Add that refills 10 000 scattered tombstones used to be 10 000-20 000 tiny memset / memcpy
// calls; each is now one kernel per ring chunk.
namespace {
// row i of `src` (pinned host memory or device) -> row slots[i] of the block at `base`
__global__ void scatter_rows_kernel(const float *__restrict__ src, const int *__restrict__ slots, int n, int dims,
                                    float *__restrict__ base)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp; j < n; j += nwarps) {
        const float *s = src + (size_t)j * dims;
        float *d = base + (size_t)slots[j] * dims;
        if ((dims & 3) == 0) {
            const float4 *s4 = reinterpret_cast<const float4 *>(s);
            float4 *d4 = reinterpret_cast<float4 *>(d);
            for (int e = lane; e < dims / 4; e += 32) d4[e] = s4[e];
        } else {
            for (int e = lane; e < dims; e += 32) d[e] = s[e];
        }
    }
}
// zero the float32 row (and its bf16 shadow row) at element offset off[i] of the slab(s)
__global__ void zero_rows_kernel(float *__restrict__ slab, unsigned char *__restrict__ shadow, int shadow_eb,
                                 const unsigned long long *__restrict__ off, int n, int dims)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp; j < n; j += nwarps) {
        float *d = slab + off[j];
        for (int e = lane; e < dims; e += 32) d[e] = 0.f;
        if (shadow != nullptr) {  // dims * shadow_eb bytes at the same element offset (dims % 4 == 0 on shadow paths)
            unsigned int *h = reinterpret_cast<unsigned int *>(shadow + off[j] * shadow_eb);
            for (int e = lane; e < dims * shadow_eb / 4; e += 32) h[e] = 0u;
        }
    }
}
}  // namespace

cudaError_t scatter_rows_launch(const float *src, const int *slots, int n, int dims, float *d_base, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    int blocks = (n + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    scatter_rows_kernel<<<blocks, 256, 0, stream>>>(src, slots, n, dims, d_base);
    return cudaGetLastError();
}

cudaError_t zero_rows_launch(float *d_slab, void *d_shadow, int shadow_eb, const unsigned long long *elem_off, int n,
                             int dims, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    if (d_shadow != nullptr && dims % 4 != 0) d_shadow = nullptr;  // no shadow path for such a set
    int blocks = (n + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    zero_rows_kernel<<<blocks, 256, 0, stream>>>(d_slab, (unsigned char *)d_shadow, shadow_eb, elem_off, n, dims);
    return cudaGetLastError();
}

// dst row j = src row idx[j]: packs the live rows of a block (tombstone compaction)
namespace {
__global__ void gather_rows_kernel(const float *__restrict__ src, const int *__restrict__ idx, int n, int dims,
                                   float *__restrict__ dst)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp; j < n; j += nwarps) {
        const float *s = src + (size_t)idx[j] * dims;
        float *d = dst + (size_t)j * dims;
        for (int e = lane; e < dims; e += 32) d[e] = s[e];
    }
}
}  // namespace

cudaError_t gather_rows_launch(const float *d_src, const int *d_idx, int n, int dims, float *d_dst,
                               cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    int blocks = (n + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    gather_rows_kernel<<<blocks, 256, 0, stream>>>(d_src, d_idx, n, dims, d_dst);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ peer exchange + merge (multi-GPU)
// One kernel instead of ncclAllGather + merge: CTA i (query i) STORES this rank's K winners of query i
// straight into every peer's inbox over NVLink (peer memory mapped with CUDA IPC), publishes a sequence
// word per (rank, query), waits until the `world` words of query i have arrived in its own inbox, and
// folds the world x K keys into the top-K.  The payload is 8 K bytes per query per peer, so the cost is
// one NVLink store round (a few microseconds) instead of a collective's launch + protocol latency.
// Inbox (per rank, two slots alternating with the sequence parity so a fast peer can run one step ahead):
//   keys  [2][world][max_q][max_k] u64      words [2][world][max_q] u64 = seq << 1 | uncertified-flag
namespace {
__global__ void __launch_bounds__(256) exchange_merge_kernel(const ExchangeArgs a)
{
    __shared__ unsigned long long buf[kMergeBuf];
    __shared__ unsigned int s_flag;
    pdl_wait();
    if (threadIdx.x == 0) pdl_launch_dependents();
    const int qi = blockIdx.x;
    exchange_fold(a, qi, a.local_keys + (size_t)qi * a.K, a.local_flags != nullptr ? a.local_flags[qi] : 0u, buf, &s_flag,
                  a.out_keys, a.out_flags);
}
}  // namespace

cudaError_t exchange_merge_launch(void *const *d_peers, int rank, int world, int q, int K, int max_q, int max_k,
                                  unsigned long long seq, const unsigned long long *d_local_keys,
                                  const unsigned int *d_local_flags, unsigned long long *d_out_keys,
                                  unsigned int *d_out_flags, cudaStream_t stream)
{
    if (q < 1 || K < 1 || q > max_q || K > max_k || world * K > kMergeBuf || world > 256) return cudaErrorInvalidValue;
    ExchangeArgs a{d_peers, rank, world, q, K, max_q, max_k, seq, d_local_keys, d_local_flags, d_out_keys, d_out_flags,
                   5000000000ull};
    return launch_chain(exchange_merge_kernel, dim3(q), dim3(256), 0, stream, a);
}

// the same kernel with caller-built arguments (gather to one rank: gf_context_create_multi's in-process sets)
cudaError_t exchange_args_launch(const ExchangeArgs &a, cudaStream_t stream)
{
    if (a.q < 1 || a.K < 1 || a.q > a.max_q || a.K > a.max_k || a.world * a.K > kMergeBuf || a.world > 256)
        return cudaErrorInvalidValue;
    return launch_chain(exchange_merge_kernel, dim3(a.q), dim3(256), 0, stream, a);
}

cudaError_t merge_launch(const unsigned long long *d_in, int groups, int lists, int q, int K,
                         unsigned long long *d_out, cudaStream_t stream, const unsigned int *d_q_count,
                         const int *d_out_map)
{
    if (d_out_map != nullptr && groups != 1) return cudaErrorInvalidValue;
    if (K < 1 || K > kMaxK || q < 1 || lists < 1 || groups < 1) return cudaErrorInvalidValue;
    dim3 grid(q, groups);
    return launch_chain(merge_kernel, grid, dim3(kMergeThreads), 0, stream, d_in, lists, q, K, d_out, d_q_count,
                        d_out_map);
}

cudaError_t synth_fill_launch(float *d_rows, int64_t nrows, int dims, const SynthSpec &spec, int64_t first_row,
                              int normalize, cudaStream_t stream)
{
    if (nrows <= 0) return cudaSuccess;
    long long warps = nrows;
    long long blocks = (warps + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    synth_fill_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_rows, nrows, dims, spec, first_row, normalize);
    return cudaGetLastError();
}

}  // namespace gf
