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
//   d_partial : [groups][grid_x][q][K] keys out
cudaError_t scan_launch(const ScanPlan &plan, const Segment *d_segs, int nseg, uint32_t total_tiles,
                        const float *d_queries, int q, int64_t q_stride, int64_t l_stride, int K,
                        const unsigned long long *d_upper, unsigned long long *d_partial, cudaStream_t stream);

// Merge `lists` lists of K keys per (group, query): d_in [groups][lists][q][K] -> d_out [groups][q][K]
// (descending, zero padded).
cudaError_t merge_launch(const unsigned long long *d_in, int groups, int lists, int q, int K,
                         unsigned long long *d_out, cudaStream_t stream);

// Synthetic rows, bit-identical to oracle/gf_oracle.c gf_oracle_synth_rows.
cudaError_t synth_fill_launch(float *d_rows, int64_t nrows, int dims, uint64_t seed, int64_t first_row, int normalize,
                              cudaStream_t stream);

}  // namespace gf
