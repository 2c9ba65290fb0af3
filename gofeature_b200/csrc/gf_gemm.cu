// gf_gemm.cu -- tcgen05 / TMEM contraction for large query batches (Q >= 9), sm_100a.
//
// score[row][q] = <feature row, query q> for 128 feature rows x NQ queries per tile:
//   A operand = feature rows   (M = 128 rows, K-major: the row-major block exactly as it sits in
//               HBM, block.go:99), TMA 2-D box 32 floats x 128 rows, SWIZZLE_128B
//   B operand = queries        (N = NQ, K-major), TMA box 32 floats x NQ rows from the padded batch
//   D         = fp32 accumulators in TMEM, [128 lanes = rows][NQ columns = queries], double buffered
// Rows on M / queries on N makes the tensor time scale with the real batch (N/2 cycles per K = 8
// step) instead of a batch padded to 128, so the kernel stays HBM-bound up to ~256 queries per pass.
//
// kind::tf32 reads the fp32 bits and keeps 10 mantissa bits (kind::f16 / kind::i8 read a bf16 / int8 shadow of
// the rows): these scores only SELECT candidates.  Every candidate is re-scored in the canonical float32 order
// and a certificate proves that no row outside the candidate set can be in the top-K; the Q x N score
// matrix never leaves the SM except in DUMP mode (small sets / diagnostics).  Epilogue modes:
//   COLLECT  (0) compare against a per-query threshold estimated by a TILEMAX (1) sample, append the survivors
//                (select / rescore / finalize kernels follow: the two-pass chain)
//   GROUPTOP (4) no threshold: the best two keys of every 32-row group per query; grouptop_tail_kernel (one
//                thread-block cluster per query) cuts, re-scores, certifies -- the one-pass chain of small batches
//   DUMP (2) every score, for small sets and diagnostics
//
// Warp roles (192 threads): warp 0 TMA producer (one thread), warp 1 TMEM allocator + MMA issuer
// (one thread), warps 2..5 epilogue (tcgen05.ld, one feature row per thread).
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>

#include "gf_device.cuh"
#include "gf_exchange.cuh"
#include "gf_kernels.h"

namespace gf {

namespace {

#ifndef GF_GEMM_EPI16_MIN
#define GF_GEMM_EPI16_MIN 128
#endif
// warp 0 producer, warp 1 MMA, then the epilogue warps: 4 (one per TMEM lane quarter) or, from 64 queries up,
// 8 (two per quarter, each taking half of the query columns) -- one warp per scheduler cannot hide its own
// tcgen05.ld / compare latency, and at 256 queries the epilogue, not the MMA, bounded the tile time
__host__ __device__ constexpr int gemm_epi_warps(int nq) { return nq >= GF_GEMM_EPI16_MIN ? 16 : (nq >= 64 ? 8 : 4); }
__host__ __device__ constexpr int gemm_threads(int nq) { return 64 + 32 * gemm_epi_warps(nq); }
constexpr int kTileRows = 128;
constexpr int kABytes = kTileRows * 128;  // one K block = one 128-byte swizzle row per feature row
// elements per K block: 32 floats (kind::tf32 on the fp32 slab), 64 bf16 (kind::f16 on the bf16 shadow) or
// 128 int8 (kind::i8 on the int8 shadow, int32 accumulators)
template <int EB>
struct KBlock {
    static constexpr int kElems = 128 / EB;
};

struct GemmArgs {
    const Segment *segs;
    const uint32_t *gtile_start;  // nseg + 1, in 128-row tiles
    int nseg;
    int dims;
    const float *slab_base;  // row 0 of tensor map A (the shadow slab has the same row geometry)
    int mode;
    uint32_t tile_first, tile_stride, tile_count;
    int q;
    const float *thr;
    unsigned long long *cand;
    unsigned int *cnt;
    unsigned int cap;
    float *tilemax;
    float *dump;
    unsigned long long dump_ld;
    // int8 shadow (EB = 1): score = acc_int32 * row_scale * q_scale
    const float *row_scales;      // [block][scales_per_block]: scale of row j of slab block b at b * scales_per_block + j
    const float *q_scale;         // [nq]
    uint32_t scales_per_block;
    long long block_floats;       // block_size / 4: slab offset (in floats) -> block index
    // GROUPTOP mode (4): best two keys of every 32-row group
    uint2 *gtop;                  // [nq][4 quarters][tile_count]
    uint32_t groups;              // 4 * tile_count
    uint32_t tiles_per_block;     // 128-row tiles of a full block: first guess of a tile's segment
    // overlapped one-pass chains: instead of waiting for the whole previous kernel (griddepcontrol.wait -- grids
    // retire in stream order, so that would also wait for the previous STEP's tail), wait until the query
    // preparation's CTAs have counted themselves in
    const unsigned int *wait_ctr;
    unsigned int wait_expect;
};

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
// thread-block cluster helpers (the one-pass chain's tail): `leader_addr` maps a local shared address to the same
// offset in CTA 0 of the cluster
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t leader_addr(const void *p)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(smem_u32(p)));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T over 32 bytes of K: one K = 8 step of kind::tf32 (EB = 4)
// or one K = 16 step of kind::f16 with bf16 operands (EB = 2)
template <int EB>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate)
{
    if (EB == 1) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else if (EB == 4) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// K-major, SWIZZLE_128B shared-memory operand descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4, LBO = 1 (unused for swizzled K-major), SBO = 1024 B between 8-row groups,
// version = 1 (Blackwell), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t smem_desc_sw128(const void *p)
{
    uint64_t d = (uint64_t)((smem_u32(p) >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1), A and B formats in bits
// 7-9 / 10-12 (kind::tf32: TF32 = 2; kind::f16: BF16 = 1), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
// kind::i8: D = S32 (2), A = B = signed 8-bit (1)
template <int EB>
__host__ __device__ constexpr uint32_t instr_desc(int M, int N)
{
    return ((EB == 1 ? 2u : 1u) << 4) | ((EB == 4 ? 2u : 1u) << 7) | ((EB == 4 ? 2u : 1u) << 10) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[N]);
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// The segment a worker is in, held in registers: a worker's tiles ascend, and with the contiguous tile ranges of
// the full passes a segment lasts for ~50 tiles, so the dependent metadata loads (tile prefix -> segment -> row
// scale) leave the per-tile critical path.
struct SegCursor {
    uint32_t segi = 0, t_begin = 0, t_end = 0;  // tiles [t_begin, t_end) belong to segment segi; t_end == 0: not loaded
    uint32_t rows = 0, slot_base = 0;
    long long slab_row0 = 0;                    // row of the slab the segment starts at (TMA coordinate)
    size_t scale_base = 0;                      // int8: index of the segment's first row scale
};
__device__ __forceinline__ void seek_tile(const GemmArgs &a, SegCursor &c, uint32_t t)
{
    if (t < c.t_end) return;
    const uint32_t last = (uint32_t)a.nseg - 1u;
    uint32_t s = c.t_end == 0u ? 0u : c.segi + 1u;  // the next segment, usually
    if (s > last) s = last;
    if (!(a.gtile_start[s] <= t && (s == last || t < a.gtile_start[s + 1]))) {
        // a jump (a worker's first tile, or a strided sample): blocks fill in order, so tile / tiles-per-block is
        // the segment unless earlier blocks are partial -- verify, else binary search
        s = t / a.tiles_per_block;
        if (s > last) s = last;
        if (!(a.gtile_start[s] <= t && (s == last || t < a.gtile_start[s + 1]))) {
            uint32_t lo = 0, hi = last;
            while (lo < hi) {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if (a.gtile_start[mid] <= t) lo = mid; else hi = mid - 1;
            }
            s = lo;
        }
    }
    c.segi = s;
    c.t_begin = a.gtile_start[s];
    c.t_end = s == last ? 0xFFFFFFFFu : a.gtile_start[s + 1];
    const Segment sg = a.segs[s];
    c.rows = sg.rows;
    c.slot_base = sg.slot_base;
    c.slab_row0 = (long long)((sg.base - a.slab_base) / a.dims);
    c.scale_base = a.row_scales != nullptr ? (size_t)((sg.base - a.slab_base) / a.block_floats) * a.scales_per_block : 0;
}
__device__ __forceinline__ void locate_tile(const GemmArgs &a, uint32_t t, SegCursor &c, uint32_t &row0, uint32_t &valid)
{
    seek_tile(a, c, t);
    row0 = (t - c.t_begin) * kTileRows;
    valid = c.rows - row0 < (uint32_t)kTileRows ? c.rows - row0 : (uint32_t)kTileRows;
}

// TM = feature-row tiles that share one load of the query operand (only TM = 1 is instantiated: TM = 2 measured slower).
// COLLECT appends are staged in shared memory and flushed 128 at a time: a global atomicAdd per survivor
// inside the score loop stalled the whole epilogue warp for an L2 round trip (4 per warp per tile at 256
// queries x ~500 candidates), longer than the MMAs of the next tile take
#ifndef GF_GEMM_TMEM_BUFS
#define GF_GEMM_TMEM_BUFS 2
#endif
#ifndef GF_GEMM_STAGES_SMALL
#define GF_GEMM_STAGES_SMALL 6
#endif
constexpr int kAppendStage = 256;                  // entries in all: 16 per epilogue warp (up to 16 warps)
constexpr int kAppendPerWarp = 16;
constexpr int kAppendStageBytes = kAppendStage * 12 + 64;
// GROUPTOP: each epilogue warp stages its (best, second) keys of kGtFlush consecutive tiles per query column in its own
// piece of shared memory and writes them out 128 bytes per column at a time (rows padded to 17 entries: bank spread)
#ifndef GF_GT_FLUSH
#define GF_GT_FLUSH 64
#endif
constexpr int kGtFlush = GF_GT_FLUSH;  // a multiple of 16 (128 bytes of keys per column), 32 entries per warp store
constexpr int kGtStageRow = kGtFlush + 1;
__host__ __device__ constexpr int gt_stage_bytes(int nq) { return nq <= kGtMaxQ ? 4 * nq * kGtStageRow * 8 : 0; }

// RESB: the query operand stays RESIDENT in shared memory for the whole launch (dims * element bytes = 512 per query:
// 512-d int8, all four K blocks, loaded once) instead of travelling with every stage.  At 256 queries the 32 KB of
// queries per stage were two thirds of the L2 -> SM traffic (ncu: 1.54 GB of TMA loads for 0.52 GB of DRAM reads per
// pass) and held the tensor pipe at 51 %; resident, a stage is 16 KB of feature rows and nothing else.
constexpr int kResRowBytes = 512;
template <int NQ, int TM, bool RESB = false>
struct GemmCfg {
    static constexpr int kResBytes = RESB ? NQ * kResRowBytes : 0;
    static constexpr int kBBytes = RESB ? 0 : NQ * 128;
    static constexpr int kStageBytes = TM * kABytes + kBBytes;
    static constexpr int kStagesFit =
        (int)((232448 - 1024 - 320 - NQ * 12 - 1024 - kAppendStageBytes - gt_stage_bytes(NQ) - kResBytes) / kStageBytes);
    // small batches are pure HBM streaming with short tiles (int8: 4 K blocks of 18 KB per tile): a deep ring
    // keeps more than two tiles of loads in flight across the per-tile work of the producer
    static constexpr int kStagesCap = GF_GEMM_STAGES_SMALL;
    static constexpr int kStages = kStagesFit > (NQ <= 32 ? kStagesCap : 6) ? (NQ <= 32 ? kStagesCap : 6) : kStagesFit;
    // accumulator sets in TMEM (small batches leave most of the 512 columns free, but more than two sets measured slower)
    static constexpr int kBufFit = 512 / (TM * NQ);
    static constexpr int kBufMax = GF_GEMM_TMEM_BUFS;  // 2 measured best (8: COLLECT 0.088 -> 0.096 ms at 16 queries)
    static constexpr int kBuf = kBufFit >= kBufMax ? kBufMax : (kBufFit >= 4 && kBufMax >= 4 ? 4 : (kBufFit >= 2 ? 2 : 1));
    static_assert((2 * 11 + 2 * 8) * 8 + 8 <= 312, "barrier region");
    static constexpr int kTmemColsRaw = kBuf * TM * NQ;
    static constexpr int kTmemCols = kTmemColsRaw < 32 ? 32 : kTmemColsRaw;
    static constexpr int kColsPerEpiWarp = NQ / (gemm_epi_warps(NQ) / 4);
    static constexpr int kChunk = kColsPerEpiWarp >= 32 ? 32 : 16;  // columns per tcgen05.ld
    static_assert(kColsPerEpiWarp % kChunk == 0, "epilogue column split");
    static constexpr size_t kSmem =
        1024 /*align slack*/ + (size_t)kResBytes + (size_t)kStages * kStageBytes + 320 /*barriers*/ + NQ * 12 +
        kAppendStageBytes + gt_stage_bytes(NQ);
    static_assert((2 * kStages + 2 * kBuf) * 8 + 8 <= 304, "room for the resident operand's barrier at +304");
    static_assert(kStages >= 2, "pipeline too shallow");
    static_assert((kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols <= 512, "TMEM columns");
};

template <int NQ, int TM, int EB, bool RESB = false>
__global__ void __launch_bounds__(gemm_threads(NQ), 1)
    gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs a)
{
    static_assert(!RESB || TM == 1, "resident query operand: one tile per step");
    using Cfg = GemmCfg<NQ, TM, RESB>;
    extern __shared__ uint8_t smem_raw[];
    // (aligned by an OFFSET from the shared-memory symbol: rounding the generic address through an integer made every
    // access below a generic LD.E / ST.E / ATOM.E instead of LDS / STS / ATOMS)
    uint8_t *smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *res_b = smem_al;                        // RESB: [K block][NQ rows][128 B], each K block a TMA box
    uint8_t *smem = smem_al + Cfg::kResBytes;        // the stage ring
    uint8_t *tail = smem + (size_t)Cfg::kStages * Cfg::kStageBytes;
    uint64_t *bfull = reinterpret_cast<uint64_t *>(tail + 304);
    uint64_t *full = reinterpret_cast<uint64_t *>(tail);
    uint64_t *empty = full + Cfg::kStages;
    uint64_t *tfull = empty + Cfg::kStages;
    uint64_t *tempty = tfull + Cfg::kBuf;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + Cfg::kBuf);
    float *thr_s = reinterpret_cast<float *>(tail + 320);
    unsigned int *tmax_s = reinterpret_cast<unsigned int *>(thr_s + NQ);
    float *qs_s = reinterpret_cast<float *>(tmax_s + NQ);  // int8 path: per-query dequantisation scale
    unsigned long long *stage_key = reinterpret_cast<unsigned long long *>(qs_s + NQ);  // 8-byte aligned: NQ % 2 == 0
    int *stage_q = reinterpret_cast<int *>(stage_key + kAppendStage);
    int *stage_n = stage_q + kAppendStage;
    uint2 *gt_stage = reinterpret_cast<uint2 *>(stage_n + 16);  // [4 warps][NQ][kGtStageRow]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kBK = KBlock<EB>::kElems;
    const int kblocks = a.dims / kBK;
    // a group = the tiles one worker (CTA) takes per step: TM consecutive tiles
    constexpr int kEpiWarps = gemm_epi_warps(NQ), kEpiThreads = 32 * kEpiWarps;
    constexpr int kGroupTiles = TM;
    const uint32_t ngroups = (a.tile_count + kGroupTiles - 1) / kGroupTiles;
    const uint32_t worker = blockIdx.x;
    const uint32_t nworkers = gridDim.x;
    // which groups this worker takes: strided over the workers (the threshold sample), or -- the full passes, COLLECT
    // and GROUPTOP: every tile, stride 1 -- ONE contiguous range of tiles per worker: a worker then stays inside a
    // segment for ~50 tiles instead of changing segment with every tile
    const bool chunked = a.tile_stride == 1u && (a.mode == 0 || a.mode == 4);
    const uint32_t g_begin = chunked ? (uint32_t)((unsigned long long)worker * ngroups / nworkers) : worker;
    const uint32_t g_step = chunked ? 1u : nworkers;
    const uint32_t g_count = chunked ? (uint32_t)((unsigned long long)(worker + 1) * ngroups / nworkers) - g_begin
                                     : (ngroups > worker ? (ngroups - worker + nworkers - 1) / nworkers : 0u);

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < Cfg::kBuf; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], kEpiWarps);
        }
        if (RESB) mbar_init(bfull, 1);
        mbar_fence_init();
    }
    // The feature rows of the first tile do not depend on anything this chain computes: their TMA loads go out NOW --
    // before the wait for the query preparation, the threshold loads and the barrier below -- and only the query halves
    // of those stages follow once the queries exist.  (A launch streams for ~80 us; ~4 us of it were start-up.)
    int pre_stages = 0;
    if (TM == 1 && threadIdx.x == 0 && g_count > 0 && (a.mode == 0 || a.mode == 4)) {
        SegCursor c0;
        uint32_t row0, valid;
        locate_tile(a, a.tile_first + g_begin * a.tile_stride, c0, row0, valid);
        const int slab_row0 = (int)c0.slab_row0 + (int)row0;
        pre_stages = kblocks < Cfg::kStages ? kblocks : Cfg::kStages;
        for (int kb = 0; kb < pre_stages; ++kb) {
            mbar_arrive_expect_tx(&full[kb], Cfg::kStageBytes);
            tma_load_2d(smem + (size_t)kb * Cfg::kStageBytes, &tmA, kb * kBK, slab_row0, &full[kb]);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)Cfg::kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // everything above is CTA-local set-up and overlaps the previous kernel; below, its results are read
    if (a.wait_ctr == nullptr) {
        pdl_wait();
    } else {
        if (threadIdx.x == 0) {
            while ((int)(*(volatile const unsigned int *)a.wait_ctr - a.wait_expect) < 0) {
            }
            __threadfence();
        }
        __syncthreads();
    }
    if (a.mode != 0 && a.mode != 3 && a.mode != 4 && threadIdx.x == 0) pdl_launch_dependents();  // short launches: let the successor start up now
    for (int i = threadIdx.x; i < NQ; i += blockDim.x) {
        float th = (a.thr != nullptr && i < a.q) ? a.thr[i] : __int_as_float(0x7f800000);
        if (EB == 1) {
            // the epilogue compares acc * row_scale against thr / q_scale: one multiply per score instead of two
            const float qs = a.q_scale[i];
            qs_s[i] = qs;
            th = qs > 0.f ? th / qs : __int_as_float(0x7f800000);
        }
        thr_s[i] = th;
        tmax_s[i] = 0u;
    }
    if (threadIdx.x < 16) stage_n[threadIdx.x] = 0;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // COLLECT over int8 accumulators: the smallest threshold of every 8 query columns (the epilogue's group test);
    // the tile-maximum words are not used in this mode
    float *thrmin_s = reinterpret_cast<float *>(tmax_s);
    if (EB == 1 && a.mode == 0) {
        for (int g = threadIdx.x; g < NQ / 8; g += blockDim.x) {
            float m = __int_as_float(0x7f800000);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float t = thr_s[8 * g + j];
                m = (t < m || t != t) ? t : m;  // (a NaN threshold sticks: the group test then always passes on)
            }
            thrmin_s[g] = m;
        }
        __syncthreads();
    }

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            SegCursor segc;
            if (RESB && g_count > 0) {  // the whole query operand, once
                mbar_arrive_expect_tx(bfull, (uint32_t)kblocks * NQ * 128u);
                for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(res_b + (size_t)kb * NQ * 128, &tmB, kb * kBK, 0, bfull);
            }
            for (uint32_t itp = 0; itp < g_count; ++itp) {
                const uint32_t g = g_begin + itp * g_step;
                int slab_row[TM];
#pragma unroll
                for (int j = 0; j < TM; ++j) {
                    uint32_t i = g * TM + j;
                    if (i >= a.tile_count) i = a.tile_count - 1;  // odd tail: shadow the last tile
                    const uint32_t t = a.tile_first + i * a.tile_stride;
                    uint32_t row0, valid;
                    locate_tile(a, t, segc, row0, valid);
                    slab_row[j] = (int)segc.slab_row0 + (int)row0;
                }
                for (int kb = 0; kb < kblocks; ++kb) {
                    uint8_t *sa = smem + (size_t)s * Cfg::kStageBytes;
                    if (itp == 0 && kb < pre_stages) {  // (its feature rows are on their way since the kernel started)
                        if (!RESB) tma_load_2d(sa + TM * kABytes, &tmB, kb * kBK, 0, &full[s]);
                    } else {
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_arrive_expect_tx(&full[s], Cfg::kStageBytes);
#pragma unroll
                        for (int j = 0; j < TM; ++j) tma_load_2d(sa + j * kABytes, &tmA, kb * kBK, slab_row[j], &full[s]);
                        if (!RESB) tma_load_2d(sa + TM * kABytes, &tmB, kb * kBK, 0, &full[s]);
                    }
                    if (++s == Cfg::kStages) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc<EB>(kTileRows, NQ);
            int s = 0;
            uint32_t ph = 0, it = 0;
            if (RESB && g_count > 0) {
                mbar_wait(bfull, 0);
                tc_fence_after();
            }
            for (; it < g_count; ++it) {
                const uint32_t buf = it % Cfg::kBuf;
                mbar_wait(&tempty[buf], ((it / Cfg::kBuf) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_addr = tmem_base + buf * (TM * NQ);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint8_t *sa = smem + (size_t)s * Cfg::kStageBytes;
                    const uint64_t bdesc = smem_desc_sw128(RESB ? res_b + (size_t)kb * NQ * 128 : sa + TM * kABytes);
#pragma unroll
                    for (int j = 0; j < TM; ++j) {
                        const uint64_t adesc = smem_desc_sw128(sa + j * kABytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k)  // 32 bytes (>> 4 = 2) further along the swizzled row
                            tc_mma<EB>(d_addr + j * NQ, adesc + 2 * k, bdesc + 2 * k, idesc,
                                             (uint32_t)((kb | k) != 0));
                    }
                    tc_commit(&empty[s]);  // frees the stage when these MMAs have read it
                    if (++s == Cfg::kStages) {
                        s = 0;
                        ph ^= 1;
                    }
                }
                tc_commit(&tfull[buf]);  // accumulators of the group complete
            }
        }
    } else {
        // ===================== epilogue: one feature row per thread =====================
        const int quarter = warp & 3;  // TMEM lanes this warp may read
        const uint32_t row = quarter * 32 + lane;
        // with 8 epilogue warps, warps 2..5 take the lower half of the query columns, warps 6..9 the upper half
        constexpr int kColsPerWarp = NQ / (kEpiWarps / 4);
        const int col_first = ((warp - 2) >> 2) * kColsPerWarp;
        const int wslot = warp - 2;  // this warp's private append stage
        uint32_t it = 0;
        SegCursor segc;
        // int8: a row's dequantisation scale comes from a 4-bytes-per-row array that streams from DRAM; loaded when
        // its tile starts, that ~1 us latency sat on the epilogue's critical path once per tile (measured: the pass
        // went from HBM-bound to epilogue-bound as soon as the per-tile work grew).  It is loaded ONE TILE AHEAD.
        constexpr bool kPrefRs = EB == 1 && TM == 1;
        // (the cursor moves on to the next tile's segment here; this tile's row0 / valid / slot are taken before)
        auto load_rs = [&](uint32_t it_) -> float {
            if (it_ >= g_count) return 1.f;
            const uint32_t i_ = g_begin + it_ * g_step;
            if (i_ >= a.tile_count) return 1.f;
            uint32_t row0_, valid_;
            locate_tile(a, a.tile_first + i_ * a.tile_stride, segc, row0_, valid_);
            if (row >= valid_) return 1.f;
            return a.row_scales[segc.scale_base + row0_ + row];
        };
        float rs_pref = 1.f;
        if (kPrefRs) rs_pref = load_rs(0);
        for (; it < g_count; ++it) {
            const uint32_t g = g_begin + it * g_step;
            const uint32_t buf = it % Cfg::kBuf;
            // the wait for the accumulators comes AFTER this tile's segment / slot / row-scale loads have been
            // issued (the epilogue is the critical path at large batches: their latency hides behind the wait)
            bool waited = false, released = false;
            // (three flags instead of `switch (mode)`: the compiler turned the chain of comparisons into a jump table,
            // an indirect branch per chunk of columns in the hottest loop)
            const bool m_collect = a.mode == 0, m_tilemax = a.mode == 1, m_grouptop = a.mode == 4;
#pragma unroll 1
            for (int j = 0; j < TM; ++j) {
                const uint32_t i = g * TM + j;
                if (i >= a.tile_count) break;  // shadow tile of an odd tail: nothing to report
                const uint32_t t = a.tile_first + i * a.tile_stride;
                uint32_t row0, valid;
                locate_tile(a, t, segc, row0, valid);
                const bool live = row < valid;
                const uint32_t slot = segc.slot_base + row0 + row;
                const uint32_t acc_col = buf * (TM * NQ) + j * NQ;
                float rs = 1.f;  // int8 path: this row's dequantisation scale (a segment is one slab block)
                if (kPrefRs) {
                    rs = rs_pref;
                    rs_pref = load_rs(it + 1);  // in flight across the wait below and this tile's work
                } else if (EB == 1 && live) {
                    rs = a.row_scales[segc.scale_base + row0 + row];
                }
                auto val = [&](uint32_t bits) -> float {  // accumulator -> score (int8: still without the query's scale)
                    return EB == 1 ? __int2float_rn((int)bits) * rs : __uint_as_float(bits);
                };
                if (!waited) {
                    mbar_wait(&tfull[buf], (it / Cfg::kBuf) & 1);
                    tc_fence_after();
                    waited = true;
                }
                // COLLECT: fixed per-query threshold, the rare survivor is staged and appended later.
                // A survivor is staged RAW -- (float(acc) * row scale, slot, query) -- and becomes a key when
                // the warp flushes its stage, one lane per entry.
                auto stage_raw = [&](float sc_raw, int qi) {
                    const int sp = atomicAdd(&stage_n[wslot], 1);
                    if (sp < kAppendPerWarp) {
                        stage_key[wslot * kAppendPerWarp + sp] =
                            ((unsigned long long)slot << 32) | (unsigned long long)__float_as_uint(sc_raw);
                        stage_q[wslot * kAppendPerWarp + sp] = qi;
                    } else {  // stage full (a burst of survivors): append directly
                        const float sc = EB == 1 ? sc_raw * qs_s[qi] : sc_raw;
                        const unsigned pos = atomicAdd(&a.cnt[qi], 1u);
                        if (pos < a.cap) a.cand[(size_t)qi * a.cap + pos] = make_key(sc, slot);
                    }
                };
                if constexpr (EB == 1 && TM == 1) {
                    if (m_collect) {
                        // int8 COLLECT.  score > threshold  <=>  sign of fma(-float(acc), row scale, threshold').
                        // One such test per score (I2F + FFMA + SHF, and a 32-way select chain per survivor) made
                        // the epilogue the bound of the 128 / 256-query passes and held the accumulator set for
                        // ~2 500 cycles per tile: the MMAs of tile i+2 waited for it, the ring filled and drained in
                        // bursts, and a 5-stage ring cannot cover the HBM latency of bursts (256 queries: 2.5 us per
                        // tile for 1.3 us of MMAs).  Survivors are ~8e-4 of the scores, so the test runs on the MAXIMUM
                        // of 8 accumulators against the smallest of their 8 thresholds first: the integer maximum is 4
                        // instructions per 8 scores (VIMNMX3), the float test 3 per group.  float_ru(max) >=
                        // float_rn(acc), min threshold <= threshold, row scale >= 0 and the single rounding of the fma
                        // make the group test a superset of the per-score test (monotone); `!(t > 0)` also passes -0
                        // and NaN on.  A group that passes (every fourth) re-tests its 8 scores exactly as before.
                        // The accumulator set is released as soon as the LAST chunk of columns is in registers.
                        // (All 64 columns of a warp in registers at once released it earlier still, but at the 96
                        // registers an 18-warp CTA may have the row-scale prefetch was spilled the moment it was
                        // issued -- a store waiting for the load it was meant to hide.)
                        static_assert(Cfg::kChunk % 8 == 0 && Cfg::kChunk <= 32, "groups of 8 columns");
#pragma unroll 1
                        for (int c0 = col_first; c0 < col_first + kColsPerWarp; c0 += Cfg::kChunk) {
                            uint32_t r[Cfg::kChunk];
                            tmem_ld<Cfg::kChunk>(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc_col + c0, r);
                            tmem_ld_wait();
                            if (c0 + Cfg::kChunk >= col_first + kColsPerWarp) {
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(&tempty[buf]);
                                released = true;
                            }
                            // group tests, branch-free: bit g of `trig` = group g may hold a survivor of this row
                            unsigned trig = 0;
#pragma unroll
                            for (int g = 0; g < Cfg::kChunk / 8; ++g) {
                                const int b = 8 * g;
                                int m = __vimax3_s32((int)r[b], (int)r[b + 1], (int)r[b + 2]);
                                const int m2 = __vimax3_s32((int)r[b + 3], (int)r[b + 4], (int)r[b + 5]);
                                m = __vimax3_s32(m, (int)r[b + 6], (int)r[b + 7]);
                                m = max(m, m2);
                                const float tp = __fmaf_rn(-__int2float_ru(m), rs, thrmin_s[(c0 >> 3) + g]);
                                trig |= (tp > 0.f) ? 0u : (1u << g);
                            }
                            if (!live) trig = 0;
                            // the rare group that passed: its 8 accumulators move to a common set of registers (ONE
                            // copy of the code below for all groups -- unrolled per group, with the survivor path
                            // inline, the kernel was 6 760 instructions and spent two thirds of its time in
                            // instruction-cache misses)
                            while (trig) {
                                const int g = __ffs(trig) - 1;
                                trig &= trig - 1;
                                uint32_t v[8];
                                if (g == 0 || Cfg::kChunk == 8) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) v[j] = r[j];
                                } else if (g == 1 || Cfg::kChunk == 16) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) v[j] = r[(8 + j) % Cfg::kChunk];
                                } else if (g == 2) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) v[j] = r[(16 + j) % Cfg::kChunk];
                                } else {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) v[j] = r[(24 + j) % Cfg::kChunk];
                                }
                                const int cg = c0 + 8 * g;
                                const float4 tha = *reinterpret_cast<const float4 *>(thr_s + cg);
                                const float4 thb = *reinterpret_cast<const float4 *>(thr_s + cg + 4);
                                const float th[8] = {tha.x, tha.y, tha.z, tha.w, thb.x, thb.y, thb.z, thb.w};
                                unsigned pass = 0;  // column j of the group -> bit 7 - j
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    pass = __funnelshift_l(
                                        __float_as_uint(__fmaf_rn(-__int2float_rn((int)v[j]), rs, th[j])), pass, 1);
                                while (pass) {
                                    const int bit = __ffs(pass) - 1;
                                    pass &= pass - 1;
                                    const int j = 7 - bit;
                                    uint32_t acc = v[0];
#pragma unroll
                                    for (int jj = 1; jj < 8; ++jj)
                                        if (jj == j) acc = v[jj];
                                    stage_raw(__int2float_rn((int)acc) * rs, cg + j);
                                }
                            }
                        }
                        continue;  // (TM == 1: the tile is done)
                    }
                }
#pragma unroll 1
                for (int c0 = col_first; c0 < col_first + kColsPerWarp; c0 += Cfg::kChunk) {
                    uint32_t r[Cfg::kChunk];
                    tmem_ld<Cfg::kChunk>(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc_col + c0, r);
                    tmem_ld_wait();
                    if (m_collect) {
                        if constexpr (EB == 1) {
                            // (int8 COLLECT runs before this loop: all columns in registers, accumulator released)
                        } else {
                            // float accumulators: score > threshold  <=>  sign of (threshold - score); the sign bits
                            // are shifted into `pass` one per score (column jj ends up in bit kChunk-1-jj): 2
                            // instructions per score (FADD, SHF)
                            unsigned pass = 0;
                            const float4 *thr4 = reinterpret_cast<const float4 *>(thr_s + c0);  // 16-byte aligned
                            auto below = [&](uint32_t bits, float th) -> uint32_t {
                                return __float_as_uint(__fadd_rn(th, -__uint_as_float(bits)));
                            };
#pragma unroll
                            for (int jj = 0; jj < Cfg::kChunk; jj += 4) {
                                const float4 th = thr4[jj >> 2];
                                pass = __funnelshift_l(below(r[jj + 0], th.x), pass, 1);
                                pass = __funnelshift_l(below(r[jj + 1], th.y), pass, 1);
                                pass = __funnelshift_l(below(r[jj + 2], th.z), pass, 1);
                                pass = __funnelshift_l(below(r[jj + 3], th.w), pass, 1);
                            }
                            if (!live) pass = 0;
                            while (pass) {
                                const int bit = __ffs(pass) - 1;
                                pass &= pass - 1;
                                const int b = Cfg::kChunk - 1 - bit;
                                float sc = 0.f;
#pragma unroll
                                for (int jj = 0; jj < Cfg::kChunk; ++jj)
                                    if (jj == b) sc = val(r[jj]);
                                stage_raw(sc, c0 + b);
                            }
                        }
                    } else if (m_tilemax) {
                        // TILEMAX: per query the best score of this tile (threshold estimation sample)
                        float v[Cfg::kChunk];
#pragma unroll
                        for (int jj = 0; jj < Cfg::kChunk; ++jj)
                            v[jj] = live ? (EB == 1 ? val(r[jj]) * qs_s[c0 + jj] : val(r[jj])) : __int_as_float(0xff800000);
                        // transposed max-reduce: lane jj ends with the max of column c0 + jj over 32 rows
                        int off = 16;
#pragma unroll
                        for (int n = Cfg::kChunk; n > 1; n >>= 1) {
                            const bool up = (lane & off) != 0;
#pragma unroll
                            for (int k = 0; k < n / 2; ++k) {
                                const float send = up ? v[k] : v[k + n / 2];
                                const float keep = up ? v[k + n / 2] : v[k];
                                v[k] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, off));
                            }
                            off >>= 1;
                        }
                        float m = v[0];
                        for (; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
                        const int col = c0 + (lane >> (Cfg::kChunk == 32 ? 0 : 1));
                        if (Cfg::kChunk == 32 || (lane & 1) == 0) atomicMax(&tmax_s[col], score_image(m));
                    } else if (m_grouptop) {
                        if constexpr (NQ <= kGtMaxQ && TM == 1) {
                            // the scores are in registers (one chunk holds all NQ columns): release the accumulator
                            // NOW -- an arrive (release) issued after the global store / histogram atomic below
                            // would wait for their acknowledgement, an L2 round trip per tile
                            static_assert(Cfg::kChunk == kColsPerWarp, "GROUPTOP: one chunk per warp");
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tempty[buf]);
                            released = true;
                            // GROUPTOP: per query the best TWO keys of this warp's 32 rows.  key = score image with
                            // its low 5 bits replaced by the lane (the row inside the group): max() carries the
                            // argmax along, and a hidden row's image is <= (second key | 31).  Transposed
                            // top-2 reduce: at each step a lane hands half of its columns to its partner.
                            uint32_t k1[Cfg::kChunk], k2[Cfg::kChunk];
#pragma unroll
                            for (int jj = 0; jj < Cfg::kChunk; ++jj) {
                                // branch-free orderable image (-0.0 sorts just below +0.0, NaN on top: approximate
                                // keys only select candidates, a NaN never passes the certificate)
                                const uint32_t u = __float_as_uint(val(r[jj]));
                                const uint32_t im = u ^ ((uint32_t)((int)u >> 31) | 0x80000000u);
                                k1[jj] = live ? ((im & ~31u) | (uint32_t)lane) : 0u;
                                k2[jj] = 0u;
                            }
                            int off = 16;
#pragma unroll
                            for (int n = Cfg::kChunk; n > 1; n >>= 1) {
                                const bool up = (lane & off) != 0;
#pragma unroll
                                for (int k = 0; k < n / 2; ++k) {
                                    const uint32_t s1 = up ? k1[k] : k1[k + n / 2], s2 = up ? k2[k] : k2[k + n / 2];
                                    const uint32_t m1 = up ? k1[k + n / 2] : k1[k], m2 = up ? k2[k + n / 2] : k2[k];
                                    const uint32_t r1 = __shfl_xor_sync(0xffffffffu, s1, off);
                                    // (the first step exchanges single keys: every second key is still 0)
                                    const uint32_t r2 = n == Cfg::kChunk ? 0u : __shfl_xor_sync(0xffffffffu, s2, off);
                                    k1[k] = max(m1, r1);
                                    k2[k] = max(min(m1, r1), max(m2, r2));
                                }
                                off >>= 1;
                            }
                            uint32_t b1 = k1[0], b2 = k2[0];
                            for (; off >= 1; off >>= 1) {  // 16 columns: lanes l and l ^ 1 hold halves of the same column
                                const uint32_t r1 = __shfl_xor_sync(0xffffffffu, b1, off);
                                const uint32_t r2 = __shfl_xor_sync(0xffffffffu, b2, off);
                                const uint32_t hi = max(b1, r1);
                                b2 = max(min(b1, r1), max(b2, r2));
                                b1 = hi;
                            }
                            // d_gtop is [query][quarter][tile]: a warp's keys of consecutive tiles are consecutive in
                            // memory.  They wait in the warp's own stage and leave 128 bytes per column at a time --
                            // one scattered 8-byte store per column per tile cost the pass ~0.5 us per query column.
                            const int col = c0 + (lane >> (Cfg::kChunk == 32 ? 0 : 1));
                            uint2 *stw = gt_stage + (size_t)quarter * NQ * kGtStageRow;
                            const uint32_t fslot = it % kGtFlush;
                            if (Cfg::kChunk == 32 || (lane & 1) == 0) stw[col * kGtStageRow + fslot] = make_uint2(b1, b2);
                            if (fslot == kGtFlush - 1 || it + 1 >= g_count) {
                                __syncwarp();
                                const uint32_t i0 = i - fslot;  // (a worker's tiles are consecutive in this mode)
#pragma unroll
                                for (uint32_t s0 = 0; s0 < (uint32_t)kGtFlush; s0 += 32u) {
                                    const uint32_t sl = s0 + (uint32_t)lane;
                                    if (sl <= fslot) {
                                        uint2 *dst = a.gtop + (size_t)quarter * a.tile_count + i0 + sl;
#pragma unroll 4
                                        for (int c = 0; c < a.q; ++c) dst[(size_t)c * a.groups] = stw[c * kGtStageRow + sl];
                                    }
                                }
                                __syncwarp();
                            }
                        }
                    } else {
                        // DUMP: every score, coalesced over rows (small sets and diagnostics)
                        if (live) {
#pragma unroll
                            for (int jj = 0; jj < Cfg::kChunk; ++jj)
                                if (c0 + jj < a.q)
                                    a.dump[(size_t)(c0 + jj) * a.dump_ld + (size_t)i * kTileRows + row] =
                                        EB == 1 ? val(r[jj]) * qs_s[c0 + jj] : val(r[jj]);
                        }
                    }
                }
                if (m_tilemax) {
                    named_bar_sync(2, kEpiThreads);
                    for (int c = threadIdx.x - 64; c < NQ; c += kEpiThreads) {
                        a.tilemax[(size_t)i * NQ + c] = image_score(tmax_s[c]);
                        tmax_s[c] = 0u;
                    }
                    named_bar_sync(2, kEpiThreads);
                }
            }
            if (!waited) {  // (only a shadow tile of an odd tail gets here without having waited)
                mbar_wait(&tfull[buf], (it / Cfg::kBuf) & 1);
                tc_fence_after();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0 && !released) {
                mbar_arrive(&tempty[buf]);
            }
            if (a.mode == 0) {
                // the accumulator is released; each warp flushes ITS staged survivors when its stage is half
                // full or this was the CTA's last group -- no barrier between the epilogue warps
                __syncwarp();
                const int staged = stage_n[wslot];
                if (staged >= kAppendPerWarp / 2 || it + 1 >= g_count) {
                    const int n = staged < kAppendPerWarp ? staged : kAppendPerWarp;
                    if (lane < n) {
                        const int qi = stage_q[wslot * kAppendPerWarp + lane];
                        const unsigned long long raw = stage_key[wslot * kAppendPerWarp + lane];
                        float sc = __uint_as_float((uint32_t)raw);
                        if (EB == 1) sc *= qs_s[qi];
                        const unsigned pos = atomicAdd(&a.cnt[qi], 1u);
                        if (pos < a.cap) a.cand[(size_t)qi * a.cap + pos] = make_key(sc, (uint32_t)(raw >> 32));
                    }
                    __syncwarp();
                    if (lane == 0) stage_n[wslot] = 0;
                    __syncwarp();
                }
            }
        }
        if ((a.mode == 0 || a.mode == 4) && warp == 2 && lane == 0) pdl_launch_dependents();  // this CTA's tiles are done
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols)
                     : "memory");
}

// ------------------------------------------------------------------ helper kernels
// queries [q][dims] (any stride) -> padded row-major batch [nq][dims] (rows >= q zeroed) + |q|^2
// With a shadow: dst_sh [nq][dims] = the batch in the shadow's format -- bf16 round-to-nearest (sh_eb = 2) or
// int8 with one scale per query, scale = max|q| / 127 (sh_eb = 1, q_scale out) -- and qres2 = |q - dequantised|^2.
__global__ void prep_queries_kernel(const float *__restrict__ src, long long q_stride, long long l_stride, int q,
                                    int nq, int dims, float *__restrict__ dst, float *__restrict__ qnorm2,
                                    float *__restrict__ copy, void *__restrict__ dst_sh, int sh_eb,
                                    float *__restrict__ qres2, float *__restrict__ q_scale,
                                    const unsigned int *__restrict__ wait_ctr, unsigned int wait_expect,
                                    unsigned int *__restrict__ done_ctr)
{
    // Plain chains: the previous call's kernels may still be reading dst / dst_sh -> wait for them.  Overlapped
    // one-pass chains (wait_ctr != null) write the OTHER of two scratch sets: nothing of the previous step reads it,
    // so this kernel does not wait for the previous step's tail at all -- that tail then runs next to this step's
    // candidate pass.  The step before the previous one used this set: its tail's completion counter is checked.
    if (wait_ctr == nullptr) pdl_wait();
    pdl_launch_dependents();
    if (wait_ctr != nullptr) {
        if (threadIdx.x == 0)
            while ((int)(*(volatile const unsigned int *)wait_ctr - wait_expect) < 0) {
            }
        __syncthreads();
    }
    const int qi = blockIdx.x;
    __shared__ float red[32], red2[32];
    __shared__ float s_scale;
    float acc = 0.f, res = 0.f, amax = 0.f;
    // up to kHold elements per thread stay in registers between the two passes (dims <= 1024 with 128 threads):
    // re-reading the padded copy from global memory was a second dependent round trip in a kernel that is all latency
    constexpr int kHold = 8;
    const bool hold = dims <= kHold * (int)blockDim.x;
    float hv[kHold];
#pragma unroll
    for (int u = 0; u < kHold; ++u) hv[u] = 0.f;
    if (hold) {
#pragma unroll
        for (int u = 0; u < kHold; ++u) {
            const int e = threadIdx.x + u * blockDim.x;
            if (e < dims && qi < q) hv[u] = src[(long long)qi * q_stride + (long long)e * l_stride];
        }
    }
#pragma unroll
    for (int u = 0; u < kHold; ++u) {
        const int e = threadIdx.x + u * blockDim.x;
        if (!hold || e >= dims) continue;
        const float v = hv[u];
        dst[(size_t)qi * dims + e] = v;
        if (copy != nullptr && qi < q) copy[(size_t)qi * dims + e] = v;
        acc += v * v;
        amax = fmaxf(amax, fabsf(v));  // (NaN is skipped by fmaxf: the norm below still turns NaN)
        if (dst_sh != nullptr && sh_eb == 2) {
            const __nv_bfloat16 b = __float2bfloat16_rn(v);
            reinterpret_cast<__nv_bfloat16 *>(dst_sh)[(size_t)qi * dims + e] = b;
            const float d = __bfloat162float(b) - v;
            res += d * d;
        }
    }
    for (int e = threadIdx.x; e < dims && !hold; e += blockDim.x) {
        float v = qi < q ? src[(long long)qi * q_stride + (long long)e * l_stride] : 0.f;
        dst[(size_t)qi * dims + e] = v;
        if (copy != nullptr && qi < q) copy[(size_t)qi * dims + e] = v;
        acc += v * v;
        amax = fmaxf(amax, fabsf(v));
        if (dst_sh != nullptr && sh_eb == 2) {
            const __nv_bfloat16 b = __float2bfloat16_rn(v);
            reinterpret_cast<__nv_bfloat16 *>(dst_sh)[(size_t)qi * dims + e] = b;
            const float d = __bfloat162float(b) - v;
            res += d * d;
        }
    }
    if (dst_sh != nullptr && sh_eb == 1) {
        for (int off = 16; off >= 1; off >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
        __syncthreads();
        if (threadIdx.x == 0) {
            float m = 0.f;
            for (int w = 0; w < (blockDim.x + 31) / 32; ++w) m = fmaxf(m, red[w]);
            s_scale = m / 127.f;
        }
        __syncthreads();
        const float sc = s_scale, inv = sc > 0.f ? 1.f / sc : 0.f;
        auto quant = [&](int e, float v) {
            int qv = __float2int_rn(v * inv);
            qv = qv > 127 ? 127 : (qv < -127 ? -127 : qv);
            reinterpret_cast<signed char *>(dst_sh)[(size_t)qi * dims + e] = (signed char)qv;
            const float d = (float)qv * sc - v;
            res += d * d;
        };
#pragma unroll
        for (int u = 0; u < kHold; ++u) {
            const int e = threadIdx.x + u * blockDim.x;
            if (hold && e < dims) quant(e, hv[u]);
        }
        for (int e = threadIdx.x; e < dims && !hold; e += blockDim.x) quant(e, dst[(size_t)qi * dims + e]);
        __syncthreads();
    }
    for (int off = 16; off >= 1; off >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, off);
        res += __shfl_xor_sync(0xffffffffu, res, off);
    }
    if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5] = acc;
        red2[threadIdx.x >> 5] = res;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f, t2 = 0.f;
        for (int w = 0; w < (blockDim.x + 31) / 32; ++w) {
            t += red[w];
            t2 += red2[w];
        }
        qnorm2[qi] = t * 1.0001f;  // upper bounds are what the certificate needs
        if (qres2 != nullptr) qres2[qi] = (t2 == t2 ? t2 : __int_as_float(0x7f800000)) * 1.0001f;
        if (q_scale != nullptr) q_scale[qi] = (dst_sh != nullptr && sh_eb == 1) ? s_scale : 1.f;
    }
    if (done_ctr != nullptr) {  // the candidate pass of an overlapped chain waits for nq of these
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(done_ctr, 1u);
    }
}

// max over rows of |row|^2 (upper-bounded): one warp per row, atomicMax on the float bits (>= 0)
__global__ void rownorm_max_kernel(const Segment *__restrict__ segs, int nseg, int dims, unsigned int *out_bits)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float best = 0.f;
    for (int s = 0; s < nseg; ++s) {
        const Segment sg = segs[s];
        for (long long r = warp; r < sg.rows; r += nwarps) {
            const float *x = sg.base + (size_t)r * dims;
            float acc = 0.f;
            for (int e = lane; e < dims; e += 32) {
                float v = x[e];
                acc += v * v;
            }
            for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (acc == acc) best = fmaxf(best, acc);
        }
    }
    if (lane == 0 && best > 0.f) atomicMax(out_bits, __float_as_uint(best * 1.0001f));
}

// per query: threshold = m-th largest of `ntiles` tile maxima; also resets the candidate counter.
// One warp per query: MSB-first radix select on the orderable score images (32 count steps).
template <int NV>
__global__ void threshold_kernel(const float *__restrict__ tilemax, int ntiles, int nq, int q, int m,
                                 float *__restrict__ thr, unsigned int *__restrict__ cnt)
{
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= nq) return;
    unsigned int v[NV];  // NV = values per lane: 5 for one wave of 148 sample tiles
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int t = lane + 32 * i;
        v[i] = t < ntiles ? score_image(tilemax[(size_t)t * nq + qi]) : 0u;
    }
    // largest x such that count(v >= x) >= m
    unsigned int prefix = 0u;
    for (int bit = 31; bit >= 0; --bit) {
        const unsigned int cand = prefix | (1u << bit);
        int c = 0;
#pragma unroll
        for (int i = 0; i < NV; ++i) c += (v[i] >= cand) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= m) prefix = cand;
    }
    if (lane == 0) {
        thr[qi] = qi < q ? image_score(prefix) : __int_as_float(0x7f800000);
        cnt[qi] = 0u;
    }
}

// <row, query> in the canonical float32 order (oracle/gf_oracle.c: four strided partial sums per lane, then
// canonical_reduce); every lane of the warp returns the score
__device__ __forceinline__ float exact_dot(const float *__restrict__ x, const float *__restrict__ qv, int dims, int lane)
{
    float p[4] = {0.f, 0.f, 0.f, 0.f};
    if ((dims & 511) == 0) {
        // the tensor path's rows are 16-byte aligned: four 128-bit loads per operand in flight at once
        for (int j = 0; j < dims; j += 512) {
            float4 xv[4], qq[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                xv[u] = __ldg(reinterpret_cast<const float4 *>(x + j + 128 * u) + lane);
                qq[u] = __ldg(reinterpret_cast<const float4 *>(qv + j + 128 * u) + lane);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {  // same element order as the scalar loop below
                p[0] = __fmaf_rn(xv[u].x, qq[u].x, p[0]);
                p[1] = __fmaf_rn(xv[u].y, qq[u].y, p[1]);
                p[2] = __fmaf_rn(xv[u].z, qq[u].z, p[2]);
                p[3] = __fmaf_rn(xv[u].w, qq[u].w, p[3]);
            }
        }
    } else {
        for (int j = 0; j < dims; j += 128) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int e = j + 4 * lane + c;
                if (e < dims) p[c] = __fmaf_rn(x[e], qv[e], p[c]);
            }
        }
    }
    return canonical_reduce(p[0], p[1], p[2], p[3]);
}

// candidates (approx keys) -> for the best `kp` of each query: exact canonical score of the row.
// d_sel [q][kp] approx keys (descending, zero padded) in; d_exact [q][kp] exact keys out.
__global__ void __launch_bounds__(256)
    rescore_kernel(const Segment *__restrict__ segs, const uint32_t *__restrict__ seg_slot_base, int nseg, int dims,
                   const unsigned int *__restrict__ d_cnt, const float *__restrict__ queries, int kp,
                   const unsigned long long *__restrict__ d_sel, unsigned long long *__restrict__ d_exact)
{
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int qi = blockIdx.y;
    if (w >= kp) return;
    if (d_cnt != nullptr && (unsigned)w >= d_cnt[qi]) return;  // unsorted candidate buffer: only cnt entries are valid
    const unsigned long long key = d_sel[(size_t)qi * kp + w];
    if (key == 0ull) {
        if (lane == 0) d_exact[(size_t)qi * kp + w] = 0ull;
        return;
    }
    const uint32_t slot = key_slot(key);
    // segment whose slot range holds this slot: slot_base values ascend with the segment index
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (seg_slot_base[mid] <= slot) lo = mid; else hi = mid - 1;
    }
    const float sc = exact_dot(segs[lo].base + (size_t)(slot - seg_slot_base[lo]) * dims, queries + (size_t)qi * dims, dims, lane);
    if (lane == 0) d_exact[(size_t)qi * kp + w] = make_key(sc, slot);
}

// Best K of up to kGemmCandCap keys held 16 per thread (256 threads), written sorted (descending,
// zero padded to a power of two) to `out` in shared memory; returns how many keys were kept (>= min(K, n);
// more only when scores tie at the cut).  A radix select on the 32-bit score image finds the cut, then
// only the survivors are bitonic-sorted -- a full sort of ~1000 candidates per query was the slowest
// small kernel of the tensor path.
__device__ int block_topk_256(const unsigned long long (&mine)[kGemmCandCap / 256], int K, unsigned long long *out,
                              int *s_cnt)
{
    constexpr int R = kGemmCandCap / 256;
    // the cut = K-th largest score image, found 8 bits at a time (MSB first) with a 256-bin histogram:
    // 4 passes of 3 barriers (the bit-by-bit version took 96 barriers and was the slowest small kernel)
    __shared__ int hist[256];
    __shared__ unsigned int s_prefix;
    __shared__ int s_krem;
    unsigned int prefix = 0u, mask = 0u;
    int krem = K;
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[threadIdx.x] = 0;  // blockDim.x == 256
        __syncthreads();
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const unsigned int img = (unsigned int)(mine[i] >> 32);
            if (mine[i] != 0ull && (img & mask) == prefix) atomicAdd(&hist[(img >> shift) & 255u], 1);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // lane l owns bins 255 - 8l .. 248 - 8l (descending); find the bin holding the krem-th key
            const int lane = threadIdx.x;
            int h[8], local = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                h[j] = hist[255 - 8 * lane - j];
                local += h[j];
            }
            int incl = local;
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += t;
            }
            const int excl = incl - local;  // keys in bins above this lane's
            const bool here = excl < krem && incl >= krem;
            const unsigned ball = __ballot_sync(0xffffffffu, here);
            if (ball == 0u) {
                if (lane == 0) {  // fewer than krem keys left: keep them all
                    s_prefix = prefix;
                    s_krem = -1;
                }
            } else if (lane == __ffs(ball) - 1) {
                int run = excl, bin = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (run < krem && run + h[j] >= krem) {
                        bin = 255 - 8 * lane - j;
                        break;
                    }
                    run += h[j];
                }
                s_prefix = prefix | ((unsigned int)bin << shift);
                s_krem = krem - run;
            }
        }
        __syncthreads();
        if (s_krem < 0) break;  // (uniform) every remaining key with this prefix is kept
        prefix = s_prefix;
        krem = s_krem;
        mask |= 255u << shift;
    }
    // a partial prefix (early exit) has zero low bits: `>= prefix` keeps everything at or above it
    if (threadIdx.x == 0) *s_cnt = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < R; ++i)
        if (mine[i] != 0ull && (unsigned int)(mine[i] >> 32) >= prefix) out[atomicAdd(s_cnt, 1)] = mine[i];
    __syncthreads();
    const int m = *s_cnt;
    int np = 2;
    while (np < m || np < K) np <<= 1;
    for (int i = m + threadIdx.x; i < np; i += blockDim.x) out[i] = 0ull;
    __syncthreads();
    for (int k = 2; k <= np; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long x = out[i], y = out[ixj];
                    bool desc = ((i & k) == 0);
                    if (desc ? (x < y) : (x > y)) {
                        out[i] = y;
                        out[ixj] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
    return m;
}

// |approx - exact| bound of the candidate pass (the same bound finalize_kernel certifies with)
__device__ __forceinline__ float cert_delta(int qi, const float *qnorm2, const unsigned int *maxnorm2_bits, float err_coeff,
                                            const float *qres2, const unsigned int *maxres2_bits)
{
    const float mn = sqrtf(__uint_as_float(__ldcg(maxnorm2_bits))), qn = sqrtf(__ldcg(qnorm2 + qi));
    float delta = err_coeff * mn * qn;
    if (qres2 != nullptr) {
        const float qr = sqrtf(__ldcg(qres2 + qi));
        delta += sqrtf(__uint_as_float(__ldcg(maxres2_bits))) * (qn + qr) * 1.0001f + mn * qr + 1e-30f * (1.f + qn);
    }
    return delta;
}

// the last CTA of a call's last finalize / tail launch lists the flagged queries for the fix-up scan
__device__ __forceinline__ void fixup_list(const FixupPlan &fix)
{
    if (fix.flags_all == nullptr) return;
    __shared__ unsigned int s_ticket;
    if (threadIdx.x == 0) {
        __threadfence();
        s_ticket = atomicAdd(fix.ticket, 1u);
    }
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    if (threadIdx.x < 32) {
        __threadfence();
        const int lane = threadIdx.x;
        int n = 0;
        for (int base = 0; base < fix.q_total; base += 32) {
            const int i = base + lane;
            const bool f = i < fix.q_total && __ldcg(fix.flags_all + i) != 0u;
            const unsigned ball = __ballot_sync(0xffffffffu, f);
            const int pos = n + __popc(ball & ((1u << lane) - 1u));
            if (f && pos < fix.maxn) {
                fix.map[pos] = i;
                fix.flags_all[i] = 0u;
            }
            n += __popc(ball);
        }
        if (lane == 0) {
            *fix.count = (unsigned)(n < fix.maxn ? n : fix.maxn);
            if (n > fix.maxn) atomicAdd(fix.uncertified, (unsigned)(n - fix.maxn));
            *fix.ticket = 0u;
        }
    }
}

// per query: sort the re-scored keys, emit the best K, and certify: the K-th exact score must beat
// everything a row outside the re-scored set could reach (its approx ceiling + the TF32 error bound).
//   all == 0: d_exact [q][kp] holds the best kp candidates by approx score, d_sel their approximate keys
//   all == 1: d_exact [q][kp = cap] holds ALL min(cnt, cap) collected candidates
__global__ void __launch_bounds__(256)
    finalize_kernel(const unsigned long long *__restrict__ d_exact, const unsigned long long *__restrict__ d_sel,
                    const unsigned int *__restrict__ cnt, const float *__restrict__ thr, unsigned int cap, int kp,
                    int all, int K, long long rows_total, const float *__restrict__ qnorm2,
                    const unsigned int *__restrict__ maxnorm2_bits, float err_coeff,
                    const float *__restrict__ qres2, const unsigned int *__restrict__ maxres2_bits,
                    unsigned long long *__restrict__ d_out, unsigned int *__restrict__ d_flags, const FixupPlan fix)
{
    __shared__ unsigned long long buf[kGemmCandCap];
    __shared__ int s_cnt;
    pdl_wait();
    pdl_launch_dependents();
    const int qi = blockIdx.x;
    const unsigned int n = cnt[qi];
    const int selected = all ? (int)(n < cap ? n : cap) : (int)((unsigned)kp < n ? (unsigned)kp : n);
    const int valid = selected;
    unsigned long long mine[kGemmCandCap / 256];
#pragma unroll
    for (int i = 0; i < kGemmCandCap / 256; ++i) {
        const int idx = threadIdx.x + 256 * i;
        mine[i] = idx < valid ? d_exact[(size_t)qi * kp + idx] : 0ull;
    }
    block_topk_256(mine, K, buf, &s_cnt);
    for (int i = threadIdx.x; i < K; i += blockDim.x) d_out[(size_t)qi * K + i] = buf[i];
    if (threadIdx.x == 0) {
        const long long keff = rows_total < K ? rows_total : K;
        bool ok = true;
        if (n > cap) ok = false;               // candidate buffer overflowed: the selection saw a subset
        if ((long long)n < keff) ok = false;   // threshold too high: not even K candidates
        if (ok && (long long)n < rows_total) { // some rows were never re-scored
            float ceil_approx = thr[qi];       // rows that were never collected have approx <= thr
            if (!all && valid < selected)      // selected (sorted by approx) but not re-scored: the first of them bounds the rest
                ceil_approx = fmaxf(ceil_approx, key_score(d_sel[(size_t)qi * kp + valid]));
            else if (!all && n > (unsigned)kp) // collected but not among the best kp by approx score
                ceil_approx = fmaxf(ceil_approx, key_score(d_sel[(size_t)qi * kp + kp - 1]));
            // tf32: |approx - exact| <= coeff |row| |q| (operand truncation + accumulation).
            // shadow pass (qres2 != null; r^, q^ = the bf16 or scaled-int8 images): approx - exact =
            // <r^ - r, q^> + <r, q^ - q>, bounded by the MEASURED residual norms max|r^ - r| (kept current by the
            // shadow_rows kernels) and |q^ - q|; coeff |row| |q| covers the accumulation (fp32 in the tensor core
            // for bf16; exact int32 + two float roundings for int8).
            const float mn = sqrtf(__uint_as_float(*maxnorm2_bits)), qn = sqrtf(qnorm2[qi]);
            float delta = err_coeff * mn * qn;
            if (qres2 != nullptr) {  // |q^| <= |q| + |q^ - q|
                const float qr = sqrtf(qres2[qi]);
                delta += sqrtf(__uint_as_float(*maxres2_bits)) * (qn + qr) * 1.0001f + mn * qr + 1e-30f * (1.f + qn);
            }
            const unsigned long long kth = buf[keff - 1];
            if (kth == 0ull || !(key_score(kth) > ceil_approx + delta)) ok = false;
        }
        d_flags[qi] = ok ? 0u : 1u;  // flagged queries are re-run by the exact scan
        if (!ok && fix.flagged != nullptr) atomicAdd(fix.flagged, 1u);
    }
    fixup_list(fix);
}

// ------------------------------------------------------------------ two-pass chain: the fused tail (gf_kernels.h)
__global__ void __launch_bounds__(256) collect_tail_kernel(const CollectTail a, const FixupPlan fix)
{
    // [0, 1024): the selection, sorted by approximate score; [1024, 2048): row pointers, then exact keys;
    // [2048, 4096): the final top-K (block_topk_256 may use all 4096 entries for the selection before that)
    __shared__ unsigned long long buf[kGemmCandCap];
    __shared__ int s_cnt, s_need;
    __shared__ unsigned long long s_kth;
    pdl_wait();
    pdl_launch_dependents();
    const int qi = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int n = a.d_cnt[qi];
    const unsigned int ncl = n < a.cap ? n : a.cap;
    unsigned long long mine[kGemmCandCap / 256];
#pragma unroll
    for (int i = 0; i < kGemmCandCap / 256; ++i) {
        const unsigned idx = threadIdx.x + 256 * i;
        mine[i] = idx < ncl ? a.d_cand[(size_t)qi * a.cap + idx] : 0ull;
    }
    const int m = block_topk_256(mine, a.kp_sel, buf, &s_cnt);
    const int have = m < a.kp_sel ? m : a.kp_sel;
    unsigned long long *ex = buf + 1024;
    const float *qv = a.d_queries + (size_t)qi * a.dims;
    const int K = a.K;

    // rows [lo, hi) of the selection: locate (all threads, the dependent loads of the searches overlap), then a warp
    // per row computes the canonical float32 score
    auto rescore_range = [&](int lo, int hi) {
        for (int i = lo + (int)threadIdx.x; i < hi; i += 256) {
            const uint32_t slot = key_slot(buf[i]);
            int s = (int)(slot / a.rows_per_block);
            if (s > a.nseg - 1) s = a.nseg - 1;
            if (!(a.d_seg_slot_base[s] <= slot && (s + 1 == a.nseg || slot < a.d_seg_slot_base[s + 1]))) {
                int l = 0, h = a.nseg - 1;
                while (l < h) {
                    const int mid = (l + h + 1) >> 1;
                    if (a.d_seg_slot_base[mid] <= slot) l = mid; else h = mid - 1;
                }
                s = l;
            }
            ex[i] = (unsigned long long)(uintptr_t)(a.d_segs[s].base + (size_t)(slot - a.d_seg_slot_base[s]) * a.dims);
        }
        __syncthreads();
        for (int i = lo + warp; i < hi; i += 8) {
            const float *x = reinterpret_cast<const float *>((uintptr_t)ex[i]);
            const uint32_t slot = key_slot(buf[i]);
            __syncwarp();
            const float sc = exact_dot(x, qv, a.dims, lane);
            if (lane == 0) ex[i] = make_key(sc, slot);
        }
        __syncthreads();
    };

    const float delta = cert_delta(qi, a.d_qnorm2, a.d_maxnorm2_bits, a.err_coeff, a.d_qres2, a.d_maxres2_bits);
    const int k1 = have < K + 32 ? have : K + 32;
    rescore_range(0, k1);
    int need = have;
    if (k1 >= K && k1 < have) {
        // L = the K-th best exact key of the first k1 (keys are unique: rank by counting)
        for (int i = threadIdx.x; i < k1; i += 256) {
            const unsigned long long me = ex[i];
            int rank = 0;
            for (int j = 0; j < k1; ++j) rank += ex[j] > me ? 1 : 0;
            if (rank == K - 1) s_kth = me;
        }
        if (threadIdx.x == 0) s_need = 0;
        __syncthreads();
        // (a hair more than delta: finalize's own float test, K-th > ceiling + delta, must hold after rounding)
        const float cut = key_score(s_kth) - delta * 1.001f - 1e-30f;
        int c = 0;
        for (int i = k1 + (int)threadIdx.x; i < have; i += 256) c += (key_score(buf[i]) >= cut) ? 1 : 0;  // sorted: a prefix
        if (c) atomicAdd(&s_need, c);
        __syncthreads();
        need = cut == cut ? k1 + s_need : have;  // (NaN scores: no shortcut)
        rescore_range(k1, need);
    }

    // best K of the `need` exact keys
#pragma unroll
    for (int i = 0; i < kGemmCandCap / 256; ++i) {
        const int idx = threadIdx.x + 256 * i;
        mine[i] = idx < need ? ex[idx] : 0ull;
    }
    unsigned long long *fin = buf + 2048;
    block_topk_256(mine, K, fin, &s_cnt);
    for (int i = threadIdx.x; i < K; i += blockDim.x) a.d_out[(size_t)qi * K + i] = fin[i];
    if (threadIdx.x == 0) {
        const long long keff = a.rows_total < K ? a.rows_total : K;
        bool ok = true;
        if (n > a.cap) ok = false;              // candidate buffer overflowed: the selection saw a subset
        if ((long long)n < keff) ok = false;    // threshold too high: not even K candidates
        if (ok && (long long)n < a.rows_total) {
            float ceil_approx = a.d_thr[qi];    // rows that were never collected have approx <= thr
            if (need < have)                    // selected (sorted by approx) but not re-scored: the first bounds the rest
                ceil_approx = fmaxf(ceil_approx, key_score(buf[need]));
            else if (n > (unsigned)a.kp_sel)    // collected but not among the best kp_sel by approx score
                ceil_approx = fmaxf(ceil_approx, key_score(buf[a.kp_sel - 1]));
            const unsigned long long kth = fin[keff - 1];
            if (kth == 0ull || !(key_score(kth) > ceil_approx + delta)) ok = false;
        }
        a.d_flags[qi] = ok ? 0u : 1u;
        if (!ok && fix.flagged != nullptr) atomicAdd(fix.flagged, 1u);
    }
    fixup_list(fix);
}

// ------------------------------------------------------------------ one-pass path: the tail
// group index (quarter * tiles + tile of the set-wide tile order) + row inside the group -> the row and its slot
__device__ __forceinline__ bool gt_locate(const GtTail &a, uint32_t group, uint32_t lane_row, const float **x, uint32_t *slot)
{
    const uint32_t ntiles = a.groups >> 2;  // d_gtop is [query][quarter][tile]
    const uint32_t tile = group % ntiles, quarter = group / ntiles;
    // blocks are filled in order, so nearly every segment holds tiles_per_block tiles: guess, verify with two
    // independent loads, and only search when the guess is off (a binary search is 6+ DEPENDENT loads)
    int lo = (int)(tile / a.tiles_per_block);
    if (lo > a.nseg - 1) lo = a.nseg - 1;
    const uint32_t s0 = a.d_gtile_start[lo], s1 = a.d_gtile_start[lo + 1];
    if (!(s0 <= tile && tile < s1)) {
        int hi = a.nseg - 1;
        lo = 0;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (a.d_gtile_start[mid] <= tile) lo = mid; else hi = mid - 1;
        }
    }
    const uint32_t row = (tile - a.d_gtile_start[lo]) * (uint32_t)kTileRows + quarter * 32u + lane_row;
    if (row >= a.d_segs[lo].rows) return false;
    *x = a.d_segs[lo].base + (size_t)row * a.dims;
    *slot = a.d_segs[lo].slot_base + row;
    return true;
}

// One thread-block CLUSTER per query (kGtCluster CTAs of 512 threads): the candidates' exact keys meet in the
// leader CTA's shared memory over DSMEM, so the tail has no global atomics, no ticket and no second trip to L2.
//   A  every CTA: cut = the histogram bin holding the kp-th best group of its query (suffix scan of 4096 bins)
//   B  its share of the groups (32-entry chunks dealt round-robin to the CTAs, so that a run of good rows spreads
//      over the cluster): best key at or above the cut -> candidate, else -> running max (ceiling 1)
//   C  a warp per candidate: exact canonical score of the group's best row -> the leader's region of this CTA
//   D  leader: rank the exact keys, emit the best K; a group whose SECOND key could still reach the K-th score has
//      all its rows re-scored (the third-best row of a group is only known to be below the second key);
//      certificate: K-th exact score > every approximate ceiling + delta.
constexpr int kGtCluster = 8;
constexpr int kGtThreads = 512;
constexpr int kGtRegion = kGtCap / kGtCluster;  // exact keys one CTA may hand to the leader
constexpr int kGtLocal = 256;                   // candidates one CTA may find
__device__ __forceinline__ void st_cluster_u64(uint32_t addr, unsigned long long v)
{
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__global__ void __launch_bounds__(kGtThreads) grouptop_tail_kernel(const GtTail a)
{
    __shared__ unsigned long long r_keys[kGtCap];  // leader: region r = [r * kGtRegion, ...) written by CTA r
    __shared__ unsigned long long r_c2[kGtCap];    // (second key << 32) | (group << 5 | lane of the best row)
    // phase A: the histogram; phase D (leader, long after): the exact keys and second keys, dense
    __shared__ __align__(16) unsigned char s_union[kGtCap * 16];
    static_assert(kGtBins * 4 <= kGtCap * 16, "histogram fits the dense arrays");
    unsigned int *lh = reinterpret_cast<unsigned int *>(s_union);
    unsigned long long *d_keys = reinterpret_cast<unsigned long long *>(s_union);
    uint2 *d_c2 = reinterpret_cast<uint2 *>(s_union + kGtCap * 8);
    __shared__ uint32_t r_cnt[kGtCluster], r_ncmax[kGtCluster], r_over[kGtCluster];
    __shared__ uint32_t s_e[kGtLocal], s_k2[kGtLocal];
    __shared__ int s_wsum[kGtThreads / 32];
    __shared__ int s_cut, s_n, s_nexp;
    __shared__ unsigned int s_ncmax, s_myflag, s_xflag;
    __shared__ unsigned long long s_kth;
    __shared__ unsigned long long s_out[kGtMaxK];  // the answer, written out in one piece (it may live in pinned host memory)
    __shared__ uint32_t s_exp[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qi = blockIdx.y;
    const uint32_t rank = cluster_ctarank();
    if (tid == 0) {
        s_cut = 0;
        s_n = 0;
        s_ncmax = 0u;
    }
    pdl_wait();
    pdl_launch_dependents();
    if (a.wait_ctr != nullptr) {
        // overlapped chains: the previous step's tail may still be running next to the pass that just finished; it
        // writes the caller's result rows and pushes into the same exchange inbox -- stay behind it
        if (tid == 0)
            while ((int)(*(volatile const unsigned int *)a.wait_ctr - a.wait_expect) < 0) {
            }
        __syncthreads();
    }
    // this CTA's 32-entry chunks: chunk c = (k * kGtCluster + rank), k = iteration * 16 + warp; the first kGtPre
    // per thread are loaded NOW, next to the histogram, so that the two L2 round trips overlap
    constexpr int kGtPre = 8;
    constexpr int kWarps = kGtThreads / 32;
    const uint2 *gt = a.d_gtop + (size_t)qi * a.groups;
    auto entry_of = [&](int iter) -> uint32_t {
        return ((uint32_t)(iter * kWarps + warp) * kGtCluster + rank) * 32u + (uint32_t)lane;
    };
    uint2 pre[kGtPre];
#pragma unroll
    for (int u = 0; u < kGtPre; ++u) {
        const uint32_t e = entry_of(u);
        pre[u] = e < a.groups ? __ldcg(gt + e) : make_uint2(0u, 0u);
    }
    // ---- A: the cut.  Histogram (4096 bins = top 12 bits of the key) of every thread's BEST key -- 4096 values per
    // query, each the best of 1/4096 of the groups, so the kp-th largest of them is the kp-th best group's key give
    // or take a few per cent -- built in shared memory, summed into the leader's over DSMEM, cut found by the
    // leader and handed back.  (The pass used to count every group in a global histogram: one more scattered
    // write per group and query column in its epilogue.)
    {
        unsigned int tbest = 0u;
#pragma unroll
        for (int u = 0; u < kGtPre; ++u) tbest = max(tbest, pre[u].x);
        // (sets beyond 1 M rows: the rest of this thread's share, kGtPre independent loads at a time; read again in B)
        for (int it0 = kGtPre; entry_of(it0) - (uint32_t)lane < a.groups; it0 += kGtPre) {
            uint2 more[kGtPre];
#pragma unroll
            for (int u = 0; u < kGtPre; ++u) {
                const uint32_t e = entry_of(it0 + u);
                more[u] = e < a.groups ? __ldcg(gt + e) : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < kGtPre; ++u) tbest = max(tbest, more[u].x);
        }
        for (int b = tid; b < kGtBins; b += kGtThreads) lh[b] = 0u;
        cluster_sync_all();  // every CTA's histogram is zero before anybody adds to the leader's
        if (tbest != 0u) atomicAdd(&lh[tbest >> 20], 1u);
        __syncthreads();
        if (rank != 0) {
            for (int b = tid; b < kGtBins; b += kGtThreads) {
                const unsigned int c = lh[b];
                if (c != 0u)
                    asm volatile("red.shared::cluster.add.u32 [%0], %1;" ::"r"(leader_addr(&lh[b])), "r"(c) : "memory");
            }
        }
        cluster_sync_all();
        if (rank == 0) {
            // thread t owns bins 4095 - 8 t .. 4088 - 8 t (descending): suffix scan
            const int top = kGtBins - 1 - 8 * tid;
            int c[8], local = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                c[j] = (int)lh[top - j];
                local += c[j];
            }
            int incl = local;
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += t;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            int before = 0;
            for (int w = 0; w < warp; ++w) before += s_wsum[w];
            incl += before;
            const int excl = incl - local;
            if (excl < a.kp && incl >= a.kp) {
                int run = excl, bin = top;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (run < a.kp && run + c[j] >= a.kp) bin = top - j;
                    run += c[j];
                }
                for (uint32_t r = 0; r < (uint32_t)kGtCluster; ++r) {  // hand the cut to every CTA of the cluster
                    uint32_t ra;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(&s_cut)), "r"(r));
                    st_cluster_u32(ra, (uint32_t)bin);
                }
            }
        }
        cluster_sync_all();
    }
    const unsigned int cutimg = (unsigned int)s_cut << 20;
    // ---- B
    {
        unsigned int ncmax = 0u;
        auto classify = [&](uint32_t e, uint2 v) {
            if (v.x >= cutimg && v.x != 0u && cutimg != 0u) {
                const int pos = atomicAdd(&s_n, 1);
                if (pos < kGtLocal) {
                    s_e[pos] = (e << 5) | (v.x & 31u);
                    s_k2[pos] = v.y;
                }
            } else {
                ncmax = max(ncmax, v.x);
            }
        };
#pragma unroll
        for (int u = 0; u < kGtPre; ++u) {
            const uint32_t e = entry_of(u);
            if (e < a.groups) classify(e, pre[u]);
        }
        for (int it0 = kGtPre; entry_of(it0) - (uint32_t)lane < a.groups; it0 += kGtPre) {  // (warp-uniform bound)
            uint2 more[kGtPre];
#pragma unroll
            for (int u = 0; u < kGtPre; ++u) {
                const uint32_t e = entry_of(it0 + u);
                more[u] = e < a.groups ? __ldcg(gt + e) : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < kGtPre; ++u) {
                const uint32_t e = entry_of(it0 + u);
                if (e < a.groups) classify(e, more[u]);
            }
        }
        ncmax = __reduce_max_sync(0xffffffffu, ncmax);
        if (lane == 0 && ncmax != 0u) atomicMax(&s_ncmax, ncmax);
        __syncthreads();
    }
    // ---- C: exact scores straight into the leader's shared memory
    const float *qv = a.d_queries_padded + (size_t)qi * a.dims;
    {
        const int found_n = s_n;
        const int n = found_n < kGtRegion ? found_n : kGtRegion;
        const uint32_t keys0 = leader_addr(r_keys + rank * kGtRegion), c20 = leader_addr(r_c2 + rank * kGtRegion);
        for (int idx = warp; idx < n; idx += kWarps) {
            const uint32_t ee = s_e[idx];
            const float *x;
            uint32_t slot;
            const bool found = gt_locate(a, ee >> 5, ee & 31u, &x, &slot);  // (always: a dead row has no key)
            const float sc = found ? exact_dot(x, qv, a.dims, lane) : 0.f;
            if (lane == 0) {
                st_cluster_u64(keys0 + 8u * idx, found ? make_key(sc, slot) : 0ull);
                st_cluster_u64(c20 + 8u * idx, ((unsigned long long)(found ? s_k2[idx] : 0u) << 32) | ee);
            }
        }
        if (tid == 0) {
            st_cluster_u32(leader_addr(r_cnt + rank), (uint32_t)n);
            st_cluster_u32(leader_addr(r_ncmax + rank), s_ncmax);
            st_cluster_u32(leader_addr(r_over + rank), (found_n > kGtRegion || cutimg == 0u) ? 1u : 0u);
        }
    }
    cluster_sync_all();  // release / acquire at cluster scope: the leader sees every CTA's keys
    if (rank != 0) return;
    // ---- D: the leader
    int n = 0;
    bool over = false;
    unsigned int ncmax_all = 0u;
    int base_of[kGtCluster];
#pragma unroll
    for (int r = 0; r < kGtCluster; ++r) {
        base_of[r] = n;
        n += (int)r_cnt[r];
        over = over || r_over[r] != 0u;
        ncmax_all = max(ncmax_all, r_ncmax[r]);
    }
    for (int sl = tid; sl < kGtCap; sl += kGtThreads) {
        const int r = sl / kGtRegion, idx = sl % kGtRegion;
        if (idx < (int)r_cnt[r]) {
            int b = 0;
#pragma unroll
            for (int rr = 0; rr < kGtCluster; ++rr)
                if (rr == r) b = base_of[rr];
            d_keys[b + idx] = r_keys[sl];
            const unsigned long long c = r_c2[sl];
            d_c2[b + idx] = make_uint2((uint32_t)(c >> 32), (uint32_t)c);
        }
    }
    const long long keff = a.rows_total < a.K ? a.rows_total : a.K;
    const float qs = a.d_q_scale != nullptr ? __ldcg(a.d_q_scale + qi) : 1.f;
    const float delta = cert_delta(qi, a.d_qnorm2, a.d_maxnorm2_bits, a.err_coeff, a.d_qres2, a.d_maxres2_bits);
    // approximate score a row with key image <= im can reach (the low 5 bits of a key are not score bits)
    auto reach = [&](unsigned int im) -> float {
        const float v = image_score(im | 31u) * qs;
        return v + 4e-6f * fabsf(v);
    };
    bool ok = !over;
    unsigned long long *out = a.d_out + (size_t)qi * a.K;
    for (int round = 0; round < 2; ++round) {
        if (tid == 0) {
            s_kth = 0ull;
            s_nexp = 0;
        }
        __syncthreads();
        // rank of every key among the n exact keys (distinct: one per slot); the best K go straight out
        for (int i = tid; i < n; i += kGtThreads) {
            const unsigned long long my = d_keys[i];
            int rk = 0;
            for (int j = 0; j < n; ++j) rk += d_keys[j] > my ? 1 : 0;
            if (rk < a.K) s_out[rk] = my;
            if (rk == (int)keff - 1) s_kth = my;
        }
        for (int i = n + tid; i < a.K; i += kGtThreads) s_out[i] = 0ull;
        __syncthreads();
        const unsigned long long kth = s_kth;
        if (kth == 0ull) {  // not even K candidates
            ok = false;
            break;
        }
        const float tk = key_score(kth);
        // groups whose hidden rows (bounded by the second key) could still reach the K-th score
        for (int i = tid; i < n; i += kGtThreads) {
            const uint2 c = d_c2[i];
            if (c.x != 0u && !(tk > reach(c.x) + delta)) {
                const int pos = atomicAdd(&s_nexp, 1);
                if (pos < 8) s_exp[pos] = c.y;
                d_c2[i].x = 0u;  // (re-scored below, or the query is flagged)
            }
        }
        __syncthreads();
        const int nexp = s_nexp;
        if (nexp == 0) break;
        if (round == 1 || nexp > 8 || n + 31 * nexp > kGtCap) {
            ok = false;
            break;
        }
        // re-score every other row of those groups: a warp per row
        if (tid == 0) s_n = n;
        __syncthreads();
        for (int w = warp; w < nexp * 32; w += kWarps) {
            const uint32_t ee = s_exp[w >> 5];
            const uint32_t lr = (uint32_t)(w & 31);
            if (lr == (ee & 31u)) continue;  // the group's best row is already there
            const float *x;
            uint32_t slot;
            if (!gt_locate(a, ee >> 5, lr, &x, &slot)) continue;  // beyond the segment's rows
            const float sc = exact_dot(x, qv, a.dims, lane);
            if (lane == 0) {
                const int pos = atomicAdd(&s_n, 1);
                d_keys[pos] = make_key(sc, slot);
                d_c2[pos] = make_uint2(0u, 0u);
            }
        }
        __syncthreads();
        n = s_n;
        __syncthreads();
    }
    if (tid < a.K) out[tid] = s_out[tid];  // (every way out of the loop is behind a barrier that follows the ranking)
    if (tid == 0) {
        // rows of groups below the cut: their best key is <= the largest non-candidate key
        if (ok && ncmax_all != 0u && !(key_score(s_kth) > reach(ncmax_all) + delta)) ok = false;
        a.d_flags[qi] = ok ? 0u : 1u;
        if (!ok && a.d_flagged != nullptr) atomicAdd(a.d_flagged, 1u);
        s_myflag = ok ? 0u : 1u;
    }
    if (a.has_xchg) {
        // sharded set: push these K winners and the flag to every rank, wait for theirs, fold (the histogram /
        // dense-key storage is free again: it holds the world x K keys)
        static_assert(kGtCap * 16 >= kMergeBuf * 8, "the fold fits the tail's shared memory");
        __syncthreads();
        exchange_fold(a.xchg, qi, s_out, s_myflag, reinterpret_cast<unsigned long long *>(s_union), &s_xflag,
                      a.xchg.out_keys, a.xchg.out_flags);
    }
    if (a.done_ctr != nullptr) {  // this query's answer is out: one more finished cluster of this scratch set
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(a.done_ctr, 1u);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

cudaError_t make_map(CUtensorMap *map, const void *base, uint64_t rows, int dims, uint32_t box_rows, int eb)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return cudaErrorNotSupported;
    cuuint64_t gdim[2] = {(cuuint64_t)dims, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)dims * (cuuint64_t)eb};
    cuuint32_t box[2] = {(cuuint32_t)(128 / eb), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, eb == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                                             : CU_TENSOR_MAP_DATA_TYPE_UINT8,
                    2, (void *)base,
                    gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// resident query operand (RESB): 512 bytes per query = all K blocks of a 512-d int8 / 256-d bf16 / 128-d float32 query
template <int NQ, int EB>
cudaError_t launch_gemm_resb(const CUtensorMap &tmA, const CUtensorMap &tmB, const GemmArgs &a, int sms, cudaStream_t stream)
{
    using Cfg = GemmCfg<NQ, 1, true>;
    auto kern = gemm_kernel<NQ, 1, EB, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    const int grid = sms < (int)a.tile_count ? sms : (int)a.tile_count;
    if (grid < 1) return cudaSuccess;
    return launch_chain(kern, dim3(grid), dim3(gemm_threads(NQ)), Cfg::kSmem, stream, tmA, tmB, a);
}

template <int NQ, int TM, int EB>
cudaError_t launch_gemm(const CUtensorMap &tmA, const CUtensorMap &tmB, const GemmArgs &a, int sms,
                        cudaStream_t stream)
{
    using Cfg = GemmCfg<NQ, TM>;
    auto kern = gemm_kernel<NQ, TM, EB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
    if (e != cudaSuccess) return e;
    // one carve-out for every kernel of a search chain: an SM only changes it when it is empty, and the one-pass
    // chain's tail shares SMs with the next step's pass (grouptop_tail_launch)
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    const uint32_t ngroups = (a.tile_count + TM - 1) / TM;
    const int grid = sms < (int)ngroups ? sms : (int)ngroups;
    if (grid < 1) return cudaSuccess;
    return launch_chain(kern, dim3(grid), dim3(gemm_threads(NQ)), Cfg::kSmem, stream, tmA, tmB, a);
}

}  // namespace

int gemm_pad_queries(int q)
{
    if (q <= 16) return 16;
    if (q <= 32) return 32;
    if (q <= 64) return 64;
    if (q <= 128) return 128;
    return 256;
}

template <int EB>
static cudaError_t gemm_dispatch(const GemmLaunch &g, const CUtensorMap &tmA, const CUtensorMap &tmB, const GemmArgs &a,
                                 cudaStream_t stream)
{
    if constexpr (EB == 1) {
        // large int8 batches on 512-d rows: the query operand stays resident in shared memory
        if (g.dims * EB == kResRowBytes) {
            if (g.nq == 256) return launch_gemm_resb<256, EB>(tmA, tmB, a, g.sms, stream);
            if (g.nq == 128) return launch_gemm_resb<128, EB>(tmA, tmB, a, g.sms, stream);
        }
    }
    // (tried and dropped: two row tiles per query load, TM = 2 -- 0.58 vs 0.40 ms at 256 queries; a cta_group::2
    // pair with M = 256 -- 0.246 vs 0.210 ms; sample + thresholds + COLLECT behind a grid barrier in one launch)
    switch (g.nq) {
        case 16: return launch_gemm<16, 1, EB>(tmA, tmB, a, g.sms, stream);
        case 32: return launch_gemm<32, 1, EB>(tmA, tmB, a, g.sms, stream);
        case 64: return launch_gemm<64, 1, EB>(tmA, tmB, a, g.sms, stream);
        case 128: return launch_gemm<128, 1, EB>(tmA, tmB, a, g.sms, stream);
        case 256: return launch_gemm<256, 1, EB>(tmA, tmB, a, g.sms, stream);
    }
    return cudaErrorInvalidValue;
}

cudaError_t gemm_launch(const GemmLaunch &g, cudaStream_t stream)
{
    const bool shadow = g.shadow_base != nullptr;
    const int eb = shadow ? g.shadow_eb : 4;  // bytes per element the pass streams: 1 int8, 2 bf16, 4 float32
    if (eb != 1 && eb != 2 && eb != 4) return cudaErrorInvalidValue;
    const bool bf16 = eb == 2;
    if (g.dims % (128 / eb) != 0 || g.q < 1 || g.q > g.nq) return cudaErrorInvalidValue;
    if (shadow && g.d_queries_shadow == nullptr) return cudaErrorInvalidValue;
    if (eb == 1 && (g.d_row_scales == nullptr || g.d_q_scale == nullptr || g.block_size < 512)) return cudaErrorInvalidValue;
    CUtensorMap tmA, tmB;
    cudaError_t e = make_map(&tmA, shadow ? g.shadow_base : (const void *)g.slab_base, g.slab_rows, g.dims, kTileRows, eb);
    if (e != cudaSuccess) return e;
    e = make_map(&tmB, shadow ? g.d_queries_shadow : (const void *)g.d_queries_padded, (uint64_t)g.nq, g.dims,
                 (uint32_t)g.nq, eb);
    if (e != cudaSuccess) return e;
    GemmArgs a{};
    a.segs = g.d_segs;
    a.gtile_start = g.d_gtile_start;
    a.nseg = g.nseg;
    a.dims = g.dims;
    a.slab_base = g.slab_base;
    a.mode = g.mode;
    a.tile_first = g.tile_first;
    a.tile_stride = g.tile_stride;
    a.tile_count = g.tile_count;
    a.q = g.q;
    a.thr = g.d_thr;
    a.cand = g.d_cand;
    a.cnt = g.d_cnt;
    a.cap = g.cap;
    a.tilemax = g.d_tilemax;
    a.dump = g.d_dump;
    a.dump_ld = g.dump_ld;
    a.row_scales = g.d_row_scales;
    a.q_scale = g.d_q_scale;
    a.scales_per_block = (uint32_t)(g.block_size / 512);
    a.block_floats = g.block_size / 4;
    a.wait_ctr = g.d_wait_ctr;
    a.wait_expect = g.wait_expect;
    {
        const long long block_rows = g.block_size / ((long long)g.dims * 4);
        a.tiles_per_block = (uint32_t)((block_rows + kTileRows - 1) / kTileRows);
        if (a.tiles_per_block < 1u) a.tiles_per_block = 1u;
    }
    if (g.mode == kGemmGroupTop) {
        if (g.nq > kGtMaxQ || g.d_gtop == nullptr || g.tile_first != 0 || g.tile_stride != 1) return cudaErrorInvalidValue;
        a.gtop = g.d_gtop;
        a.groups = 4u * g.tile_count;
    }
    if (g.tile_count < 1) return cudaSuccess;
    if (eb == 1) return gemm_dispatch<1>(g, tmA, tmB, a, stream);
    return bf16 ? gemm_dispatch<2>(g, tmA, tmB, a, stream) : gemm_dispatch<4>(g, tmA, tmB, a, stream);
}

cudaError_t prep_queries_launch(const float *src, int64_t q_stride, int64_t l_stride, int q, int nq, int dims,
                                float *dst, float *qnorm2, float *copy, void *dst_shadow, int shadow_eb, float *qres2,
                                float *q_scale, cudaStream_t stream, const unsigned int *wait_ctr, unsigned int wait_expect,
                                unsigned int *done_ctr)
{
    static bool carve_set = false;
    if (!carve_set) {
        cudaFuncSetAttribute(prep_queries_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        carve_set = true;
    }
    return launch_chain(prep_queries_kernel, dim3(nq), dim3(128), 0, stream, src, (long long)q_stride,
                        (long long)l_stride, q, nq, dims, dst, qnorm2, copy, dst_shadow, shadow_eb, qres2, q_scale,
                        wait_ctr, wait_expect, done_ctr);
}

cudaError_t rownorm_max_launch(const Segment *d_segs, int nseg, int dims, unsigned int *d_out_bits, int sms,
                               cudaStream_t stream)
{
    rownorm_max_kernel<<<sms * 8, 256, 0, stream>>>(d_segs, nseg, dims, d_out_bits);
    return cudaGetLastError();
}

cudaError_t threshold_launch(const float *d_tilemax, int ntiles, int nq, int q, int m, float *d_thr,
                             unsigned int *d_cnt, cudaStream_t stream)
{
    if (ntiles > kGemmMaxSampleTiles || ntiles < 1) return cudaErrorInvalidValue;
    const int nv = (ntiles + 31) / 32;
    const int grid = (nq + 3) / 4;
    if (nv <= 5)
        return launch_chain(threshold_kernel<5>, dim3(grid), dim3(128), 0, stream, d_tilemax, ntiles, nq, q, m, d_thr, d_cnt);
    else if (nv <= 10)
        return launch_chain(threshold_kernel<10>, dim3(grid), dim3(128), 0, stream, d_tilemax, ntiles, nq, q, m, d_thr, d_cnt);
    else if (nv <= 19)
        return launch_chain(threshold_kernel<19>, dim3(grid), dim3(128), 0, stream, d_tilemax, ntiles, nq, q, m, d_thr, d_cnt);
    else
        return launch_chain(threshold_kernel<kGemmMaxSampleTiles / 32>, dim3(grid), dim3(128), 0, stream, d_tilemax, ntiles, nq, q, m, d_thr, d_cnt);
    return cudaGetLastError();
}

cudaError_t collect_tail_launch(const CollectTail &t, const FixupPlan &fix, cudaStream_t stream)
{
    if (t.cap > (unsigned)kGemmCandCap || t.kp_sel > 1024 || t.K > t.kp_sel || t.K + 32 > 1024 || t.q < 1 ||
        t.rows_per_block == 0)
        return cudaErrorInvalidValue;
    return launch_chain(collect_tail_kernel, dim3(t.q), dim3(256), 0, stream, t, fix);
}

cudaError_t grouptop_tail_launch(const GtTail &t, cudaStream_t stream)
{
    if (t.q < 1 || t.K < 1 || t.K > kGtMaxK || t.kp < t.K || t.kp > kGtCap || t.groups >= (1u << 27)) return cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kGtCluster, t.q);
    cfg.blockDim = dim3(kGtThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kGtCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    // the same shared-memory carve-out as the candidate pass: an SM can only change its carve-out when it is empty,
    // so with different preferences the next step's pass could not join this kernel's CTAs on their SMs
    static bool carve_set = false;
    if (!carve_set) {
        cudaFuncSetAttribute(grouptop_tail_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        carve_set = true;
    }
    return cudaLaunchKernelEx(&cfg, grouptop_tail_kernel, t);
}

cudaError_t rescore_launch(const Segment *d_segs, const uint32_t *d_seg_slot_base, int nseg, int dims, int q,
                           const float *d_queries_padded, int kp, const unsigned long long *d_sel,
                           const unsigned int *d_cnt, unsigned long long *d_exact, cudaStream_t stream)
{
    dim3 grid((kp * 32 + 255) / 256, q);
    return launch_chain(rescore_kernel, grid, dim3(256), 0, stream, d_segs, d_seg_slot_base, nseg, dims, d_cnt,
                        d_queries_padded, kp, d_sel, d_exact);
}

cudaError_t finalize_launch(const unsigned long long *d_exact, const unsigned long long *d_sel,
                            const unsigned int *d_cnt, const float *d_thr, unsigned int cap, int kp, int K, int q,
                            long long rows_total, const float *d_qnorm2, const unsigned int *d_maxnorm2_bits,
                            float err_coeff, const float *d_qres2, const unsigned int *d_maxres2_bits,
                            unsigned long long *d_out, unsigned int *d_flags, int all, const FixupPlan &fix,
                            cudaStream_t stream)
{
    if (kp > kGemmCandCap || K > kp) return cudaErrorInvalidValue;
    return launch_chain(finalize_kernel, dim3(q), dim3(256), 0, stream, d_exact, d_sel, d_cnt, d_thr, cap, kp, all, K,
                        rows_total, d_qnorm2, d_maxnorm2_bits, err_coeff, d_qres2, d_maxres2_bits, d_out, d_flags, fix);
}

}  // namespace gf
