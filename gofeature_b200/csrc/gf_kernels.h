// gf_kernels.h -- host-callable launchers of the sm_100a kernels (internal, C++).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gf {

// One scanned block of a set: `rows` consecutive feature rows starting at `base`
// (row j at base + j*dims floats, block.go:99,111) = slots [slot_base, slot_base+rows).
struct Segment {
    const float *base;
    uint32_t rows;
    uint32_t slot_base;
    uint32_t tile_start;  // index of this segment's first tile in the set-wide tile order
    uint32_t pad;
};
static_assert(sizeof(Segment) == 24, "Segment layout");

constexpr int kScanMaxQ = 8;        // queries one scan pass holds in registers
constexpr int kMaxK = 1024;         // largest top-K one pass selects; larger limits run in bands
constexpr int kScanConsumerWarps = 7;
constexpr int kScanThreads = 256;   // 7 consumer warps + 1 producer warp (register file is granted per 4 warps)
constexpr int kScanTileBytes = 28672;  // 7 warps x 4 KB: 14 rows at dims = 512

struct ScanPlan {
    int dims;
    int tile_rows;     // rows per tile
    int stages;        // pipeline depth
    int list_cap;      // keys per candidate list (power of two >= 2K)
    int qt;            // template query count (1,2,4,8)
    int grid_x;        // CTAs per group
    int groups;        // 1 (set-wide top-K) or nseg (per-block top-K)
    size_t smem_bytes;
    bool tma;          // bulk-copy fast path (dims in {128,256,512,1024}, 16-byte aligned blocks)
    bool warp_lists;   // bulk-copy path, sorted lists (K <= 128): every query's list is owned by one consumer warp, the others post to its mailbox (gf_device.cuh)
};

// tile_rows for a given dims on the bulk-copy path (0 = not eligible)
int scan_tile_rows(int dims);
// largest query count one scan pass can take for this dims (register / shared-memory budget)
int scan_max_q(int dims);

// Fill in a plan.  per_segment: one candidate list per (segment, query) instead of per query.
cudaError_t scan_make_plan(int device_sms, int dims, int q, int K, uint32_t total_tiles, int nseg, bool per_segment,
                           bool aligned16, ScanPlan *plan);

// Fused dot-product + streaming top-K over the segments.
//   d_queries : q queries; element l of query i at d_queries[i*q_stride + l*l_stride]
//   d_upper   : optional q exclusive upper-bound keys (band passes for limit > kMaxK), may be null
//   d_q_map / d_q_count : optional device-side indirection (fix-up of uncertified tensor-path queries):
//               query i is row d_q_map[i] of d_queries, only *d_q_count (<= q) queries are live
//   d_partial : [groups][grid_x][q][K] keys out
cudaError_t scan_launch(const ScanPlan &plan, const Segment *d_segs, int nseg, uint32_t total_tiles,
                        const float *d_queries, int q, int64_t q_stride, int64_t l_stride, int K,
                        const unsigned long long *d_upper, unsigned long long *d_partial, cudaStream_t stream,
                        const int *d_q_map = nullptr, const unsigned int *d_q_count = nullptr);

// Merge `lists` lists of K keys per (group, query): d_in [groups][lists][q][K] -> d_out [groups][q][K]
// (descending, zero padded).
// d_out_map (optional, groups == 1): query i of this launch is written to row d_out_map[i] of d_out.
cudaError_t merge_launch(const unsigned long long *d_in, int groups, int lists, int q, int K,
                         unsigned long long *d_out, cudaStream_t stream, const unsigned int *d_q_count = nullptr,
                         const int *d_out_map = nullptr);

// Peer exchange + merge over NVLink peer memory (gf_merge.cu): d_peers = world inbox base pointers, inbox
// layout keys [2][world][max_q][max_k][2] u64 (a key travels as two self-validating words, gf_exchange.cuh) then flag
// words [2][world][max_q] u64; world * K <= 2048.
constexpr int kExchangeMaxKeys = 2048;
inline size_t exchange_inbox_bytes(int world, int max_q, int max_k)
{
    return ((size_t)4 * world * max_q * max_k + (size_t)2 * world * max_q) * 8;
}
// arguments of the cross-rank step (exchange_fold, gf_exchange.cuh)
struct ExchangeArgs {
    void *const *peers;  // world inbox base pointers (device array); peers[rank] is this rank's own
    int rank, world, q, K, max_q, max_k;
    unsigned long long seq;
    const unsigned long long *local_keys;  // [q][K] (exchange_merge_kernel; the fused tail pushes from shared memory)
    const unsigned int *local_flags;       // optional [q]
    unsigned long long *out_keys;          // [q][K]
    unsigned int *out_flags;               // optional [q]: OR of every rank's flag; 0xFFFFFFFF = a peer never arrived
    unsigned long long timeout_ns;
    // 0: every rank pushes to every rank and folds (one process per GPU: all ranks need the answer).
    // r + 1: gather to rank r -- the others push their keys into r's inbox only and leave; r alone waits, folds,
    // writes out_keys / out_flags and clears the words it consumed, so the SAME sequence number serves the next
    // step (the caller does not start one before r has finished: gf_context_create_multi's in-process sets,
    // whose steps are recorded CUDA graphs with these arguments baked in).
    int gather_root_plus1;
    // optional [max_q][4] device words: globaltimer (ns) when query qi's CTA entered the exchange, had pushed and
    // published, had seen every rank's word, had folded -- the breakdown bench.py reports as `exchange_trace`
    unsigned long long *trace;
};
cudaError_t exchange_merge_launch(void *const *d_peers, int rank, int world, int q, int K, int max_q, int max_k,
                                  unsigned long long seq, const unsigned long long *d_local_keys,
                                  const unsigned int *d_local_flags, unsigned long long *d_out_keys,
                                  unsigned int *d_out_flags, cudaStream_t stream);

cudaError_t exchange_args_launch(const ExchangeArgs &a, cudaStream_t stream);

// ---- tcgen05 contraction path (gf_gemm.cu) ----
constexpr int kGemmMaxQ = 256;       // queries per pass (N of the MMA)
constexpr int kGemmCandCap = 4096;   // candidate keys kept per query
constexpr int kGemmMaxSampleTiles = 1024;
constexpr float kTf32ErrCoeff = 2.2e-3f;  // |tf32 score - exact| <= coeff * |row| * |query| (two 10-bit truncations)
constexpr float kAccErrCoeff = 2.5e-4f;   // fp32 accumulation inside the tensor core (the slack kTf32ErrCoeff holds too)
constexpr float kI8ErrCoeff = 2e-6f;      // int8 pass: exact int32 accumulation, then two float roundings of the scales
// kGemmGroupTop: ONE pass, no threshold -- per (32-row group, query) the best two approximate keys go to d_gtop;
// grouptop_tail_kernel finds the cut (a 4096-bin histogram of score images), re-scores and certifies
enum { kGemmCollect = 0, kGemmTileMax = 1, kGemmDump = 2, kGemmGroupTop = 4 };
constexpr int kGtBins = 4096;      // histogram bins = top 12 bits of the score image (sign, exponent, 3 mantissa bits)
constexpr int kGtCap = 1024;       // candidates re-scored per query at most
constexpr int kGtMaxQ = 32;        // padded batch sizes that take the one-pass path
constexpr int kGtMaxK = 32;        // ... up to this K (the tail ranks its candidates with an O(n^2 / threads) count)

struct GemmLaunch {
    const Segment *d_segs;
    const uint32_t *d_gtile_start;  // nseg + 1 prefix of 128-row tiles
    int nseg;
    int dims;
    const float *slab_base;   // first row of the cache slab (tensor map A covers the whole slab)
    uint64_t slab_rows;
    const float *d_queries_padded;  // [nq][dims]
    const void *shadow_base;        // shadow of the slab (same row geometry); null = kind::tf32 on the fp32 rows
    int shadow_eb;                  // bytes per shadow element: 2 = bf16 (kind::f16), 1 = int8 + row scales (kind::i8)
    const void *d_queries_shadow;   // [nq][dims] in the shadow's format
    const float *d_row_scales;      // int8: one scale per slab row, block b row j at b * (block_size / 512) + j
    const float *d_q_scale;         // int8: [nq]
    int64_t block_size;             // bytes of one slab block (int8: locates a row's scale)
    int nq, q;
    int mode;
    uint32_t tile_first, tile_stride, tile_count;
    const float *d_thr;
    unsigned long long *d_cand;
    unsigned int *d_cnt;
    unsigned int cap;
    float *d_tilemax;
    float *d_dump;
    unsigned long long dump_ld;
    int sms;
    uint2 *d_gtop;            // kGemmGroupTop: [nq][4 quarters][tile_count] (best key, second key) of every 32-row group
    // overlapped chains: wait for *d_wait_ctr >= wait_expect (the preparation's CTAs) instead of the previous grid
    const unsigned int *d_wait_ctr;
    unsigned int wait_expect;
};
int gemm_pad_queries(int q);
cudaError_t gemm_launch(const GemmLaunch &g, cudaStream_t stream);
// wait_ctr (optional): the kernel does not wait for its predecessor in the stream (an overlapped one-pass chain,
// Set::topk_tensor) but for *wait_ctr to reach wait_expect -- the tails that last used the scratch set it writes
cudaError_t prep_queries_launch(const float *src, int64_t q_stride, int64_t l_stride, int q, int nq, int dims,
                                float *dst, float *qnorm2, float *copy, void *dst_shadow, int shadow_eb, float *qres2,
                                float *q_scale, cudaStream_t stream, const unsigned int *wait_ctr = nullptr,
                                unsigned int wait_expect = 0, unsigned int *done_ctr = nullptr);
cudaError_t rownorm_max_launch(const Segment *d_segs, int nseg, int dims, unsigned int *d_out_bits, int sms,
                               cudaStream_t stream);
cudaError_t threshold_launch(const float *d_tilemax, int ntiles, int nq, int q, int m, float *d_thr,
                             unsigned int *d_cnt, cudaStream_t stream);
cudaError_t rescore_launch(const Segment *d_segs, const uint32_t *d_seg_slot_base, int nseg, int dims, int q,
                           const float *d_queries_padded, int kp, const unsigned long long *d_sel,
                           const unsigned int *d_cnt, unsigned long long *d_exact, cudaStream_t stream);
constexpr int kFixupMaxQ = 8;  // uncertified queries one device-side fix-up scan pass can take
// What the LAST finalize launch of a call also does, in the last CTA to finish: flags[0..q_total) (1 =
// uncertified) -> map[0..n) of their indices, *count = min(n, maxn), those flags cleared; what does not fit
// in one fix-up pass stays flagged and is added to *uncertified (the host API re-runs those).
struct FixupPlan {
    unsigned int *flags_all;  // null: this launch does not compact
    int q_total;
    int maxn;
    int *map;
    unsigned int *count;
    unsigned int *uncertified;
    unsigned int *ticket;  // zero between calls
    unsigned int *flagged = nullptr;  // always counted (atomicAdd per flagged query): gf_set_device_counters
};
// d_flags[qi] = 1 when query qi could not be certified (it must be re-run by the exact scan)
cudaError_t finalize_launch(const unsigned long long *d_exact, const unsigned long long *d_sel,
                            const unsigned int *d_cnt, const float *d_thr, unsigned int cap, int kp, int K, int q,
                            long long rows_total, const float *d_qnorm2, const unsigned int *d_maxnorm2_bits,
                            float err_coeff, const float *d_qres2, const unsigned int *d_maxres2_bits,
                            unsigned long long *d_out, unsigned int *d_flags, int all, const FixupPlan &fix,
                            cudaStream_t stream);

// The large-batch (two-pass) chain's tail as ONE kernel, one CTA per query, for all queries of up to
// kTailBatchPasses candidate passes at once (round 1 / 2a: select -> rescore -> finalize, three launches per pass of
// 256 queries, ~0.11 ms per pass against 0.13 ms for the pass itself at 1 024 queries x top-100):
//   1  best kp_sel candidates by approximate score, sorted (radix select + bitonic sort in shared memory)
//   2  exact canonical score of the best K + 32 of them; L = the K-th best exact score so far
//   3  exact score of every further candidate whose approximate score reaches L - delta.  Everything left out has
//      approx < L - delta, hence exact < L <= the final K-th score: the certificate holds by construction, and the
//      depth is ~1 delta below the K-th score instead of the 2.25 delta a depth chosen from approximate scores alone
//      needs (uniform rows, top-100 of 1 M: ~230 re-scored rows per query instead of ~470)
//   4  best K exact keys -> d_out, certificate against the COLLECT threshold and the first candidate not re-scored
struct CollectTail {
    const unsigned long long *d_cand;  // [q][cap] unsorted approximate keys
    const unsigned int *d_cnt;         // [q] keys appended (may exceed cap: overflow -> flagged)
    const float *d_thr;                // [q] COLLECT thresholds: approx <= thr for every row never collected
    unsigned int cap;
    int kp_sel;                        // depth of the approximate selection, <= 1024
    int K;
    int q;
    long long rows_total;
    const Segment *d_segs;
    const uint32_t *d_seg_slot_base;
    int nseg;
    int dims;
    uint32_t rows_per_block;           // first guess of a slot's segment: slot / rows_per_block
    const float *d_queries;            // [q][dims] float32, row-major
    const float *d_qnorm2;
    const unsigned int *d_maxnorm2_bits;
    float err_coeff;
    const float *d_qres2;
    const unsigned int *d_maxres2_bits;
    unsigned long long *d_out;         // [q][K]
    unsigned int *d_flags;             // [q]
};
constexpr int kTailBatchPasses = 4;
cudaError_t collect_tail_launch(const CollectTail &t, const FixupPlan &fix, cudaStream_t stream);

// The one-pass path's last kernel, one thread-block cluster per query: cut = the histogram bin that holds the kp-th
// best group; every group at or above the cut has its best row re-scored exactly (canonical float32 order); the
// cluster's leader ranks the exact keys, re-scores whole groups whose SECOND key could still matter, emits the best K
// and certifies them against everything that was not re-scored -- and, for a sharded set, ends with the cross-rank
// exchange (has_xchg).
struct GtTail {
    const uint2 *d_gtop;
    uint32_t groups;              // 4 * tile_count
    const Segment *d_segs;
    const uint32_t *d_gtile_start;
    int nseg, dims;
    const float *d_queries_padded;
    const float *d_q_scale;       // int8 shadow: the keys hold acc * row_scale, still without the query's scale; else null
    int kp, K, q;
    long long rows_total;
    const float *d_qnorm2;
    const unsigned int *d_maxnorm2_bits;
    float err_coeff;
    const float *d_qres2;
    const unsigned int *d_maxres2_bits;
    unsigned long long *d_out;    // [q][K]
    unsigned int *d_flags;        // [q]
    unsigned int *d_flagged;      // set-wide counter of queries the certificate refused (gf_set_device_counters)
    uint32_t tiles_per_block;     // 128-row tiles of a full block: first guess of a tile's segment
    // sharded sets: the leader ends with the cross-rank exchange + fold (d_out / d_flags then take the LOCAL winners)
    int has_xchg;
    ExchangeArgs xchg;
    // overlapped chains: stay behind the previous step's tail (*wait_ctr >= wait_expect), count this step's finished
    // clusters in *done_ctr
    const unsigned int *wait_ctr;
    unsigned int wait_expect;
    unsigned int *done_ctr;
};
cudaError_t grouptop_tail_launch(const GtTail &t, cudaStream_t stream);

cudaError_t rownorm_range_launch(const float *d_rows, int64_t rows, int dims, unsigned int *d_out_bits,
                                 cudaStream_t stream);
// rows -> bf16 shadow rows (round to nearest), and the certificate's bounds of the set:
// d_small[0] = max |row|^2 bits, d_small[2] = max |bf16(row) - row|^2 bits (atomicMax, upper-bounded)
cudaError_t shadow_rows_launch(const float *d_rows, int64_t rows, int dims, void *d_shadow_rows, unsigned int *d_small,
                               cudaStream_t stream);
// int8 variant: shadow row = round(row / scale), scale = max|row| / 127 (one float per row, d_scales[i]);
// the same two bounds in d_small, the residual measured against scale * int8
cudaError_t shadow_rows_i8_launch(const float *d_rows, int64_t rows, int dims, void *d_shadow_rows, float *d_scales,
                                  unsigned int *d_small, cudaStream_t stream);

// src / slots / elem_off may live in pinned host memory (the upload ring): the kernels read them over PCIe
cudaError_t scatter_rows_launch(const float *src, const int *slots, int n, int dims, float *d_base, cudaStream_t stream);
cudaError_t zero_rows_launch(float *d_slab, void *d_shadow, int shadow_eb, const unsigned long long *elem_off, int n,
                             int dims, cudaStream_t stream);
cudaError_t gather_rows_launch(const float *d_src, const int *d_idx, int n, int dims, float *d_dst,
                               cudaStream_t stream);

// What a synthetic row is made of (bit-identical to oracle/gf_oracle.c gf_oracle_synth_rows / _synth_rows_ex):
//   centres == 0  row(o)[e] = U(-1,1) of a counter hash (search_test.go:65), the uniform recipe
//   centres  > 0  row(o) = centre(c(o)) + spread * U(-1,1)^dims, c(o) a hash of the ordinal: `centres` clusters whose
//                 members are scattered over the set (what a face-feature index holds: many near neighbours)
//   dup_ppm       that many rows per million are exact copies of an EARLIER ordinal's row
// then per-row L2 normalisation (normalize != 0).
struct SynthSpec {
    uint64_t seed;
    uint64_t centres = 0;
    float spread = 0.f;
    uint32_t dup_ppm = 0;
    SynthSpec(uint64_t s = 0) : seed(s) {}
};
cudaError_t synth_fill_launch(float *d_rows, int64_t nrows, int dims, const SynthSpec &spec, int64_t first_row,
                              int normalize, cudaStream_t stream);

}  // namespace gf
