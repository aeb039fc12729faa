// Single-frame row engine: the row operators of a KernelUpdateHead stage for ONE frame (or a few) as TWO launches.
//
// The online VPS caller hands the loop one frame per call (knet/video/kernel_iter_head.py:435-468): 100 kernel rows.  The
// warp-MMA chain of smallops.cu spends 14 dependent launches per stage on them (0.30 ms per frame at S = 3, launch latency);
// the tcgen05 chain of rowgemm_tc.cu needs hundreds of rows to fill its 128-row tiles.  Here a thread-block CLUSTER of
// 8 CTAs owns a 16-row tile and walks the whole operator sequence without leaving the chip:
//
//   * CTA r of the cluster computes output columns [32 r, 32 r + 32) of every Linear (mma.sync m16n8k16, the fp32 rows as
//     three bf16 planes so that every product is exact -- the numerics of smallops.cu), i.e. 1/8 of every weight matrix,
//     streamed once through a shared-memory ring of 32-row weight chunks (bulk copies, mbarrier completion) that starts
//     BEFORE the programmatic-dependency wait;
//   * the row transforms between the Linears (LayerNorm, the KernelUpdator gate, ReLU, residuals) run once per element on
//     the slice owner; LayerNorm statistics are merged across the 8 slices (Chan's formula on per-slice mean / M2) through
//     distributed shared memory;
//   * the owner hands its transformed slice, as bf16 planes, to the A-operand buffer of all 8 CTAs: the buffer is K-blocked (one
//     XOR-swizzled 3 KB block per 32-column slice), so a hand-off is ONE bulk copy shared -> distributed shared memory per
//     destination, reporting its bytes to a single-use mbarrier of the receiver -- no cluster-wide barrier, no per-thread remote
//     stores (those were 1.5-2.0 us per hand-off, the bulk copies 0.8-1.3 us); activations never touch global memory between two
//     Linears;
//   * the 8 heads of the attention map onto the 8 CTAs (head r = columns [32 r, 32 r + 32) of q / k / v / the attention
//     output), the FFN's hidden columns too: CTA r keeps its 256 hidden channels local (they are the K slice of its partial
//     second Linear) and the partial outputs are reduce-scattered to the column owners in a fixed order (deterministic).
//
// The attention needs the k / v rows of every tile of the frame, so the stage is split there: kernel A = x_feat, KernelUpdator,
// fc_norm, in-proj (writes q / k / v); kernel B = attention, out-proj + LN, FFN + LN, FC heads, fold (writes the mask conv's
// operand planes).  With pooling, its reduce and the mask conv: 5 launches per stage instead of 17.
//
// Reference math: knet/kernel_updator.py:56-94, knet/det/kernel_update_head.py:201-227 (+ the folded feat_transform).
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace vkn {

constexpr int FC_NT = 256;                      // 8 warps
constexpr int FC_CL = 8;                        // CTAs per cluster = column slices = attention heads
constexpr int FC_TM = 16;                       // rows per tile (one m16 MMA tile)
constexpr int FC_SW = 32;                       // columns per slice
constexpr int FC_K = 256;                       // K of every GEMM step (= C; the FFN's second Linear: a 256-wide K slice per CTA)
constexpr int FC_LD = FC_K + 8;                 // bf16 row stride of planes and weight chunks (528 B: conflict-free ldmatrix)
// A-operand buffer: K-BLOCKED, one 3072-byte block per 32-column slice = [plane 3][row 16][32 bf16], the four 16-byte pieces of a
// 64-byte row XOR-swizzled with (row >> 1) & 3 (conflict-free ldmatrix without padding).  A block is exactly what one CTA of the
// cluster hands to the others: ONE bulk copy (shared -> distributed shared memory) per destination.
constexpr int FC_BLK = 3 * FC_TM * FC_SW * 2;   // bytes of a slice block
constexpr int FC_ABUF = FC_CL * FC_BLK;         // bytes of an A-operand buffer (K = 256)
__device__ __forceinline__ uint32_t fc_a_off(int pl, int row, int k) {       // byte offset of element (plane, row, k), k even-aligned use
  return (uint32_t)((k >> 5) * FC_BLK + pl * (FC_TM * FC_SW * 2) + row * (FC_SW * 2) + ((((k & 31) >> 3) ^ ((row >> 1) & 3)) << 4) +
                    (k & 7) * 2);
}
constexpr int FC_CHUNK_B = 32 * FC_LD * 2;      // bytes of a weight chunk (32 rows x K)
constexpr int FC_MAXSLOT = 8;
constexpr int FC_MAXCHUNK = 24;
constexpr int FC_MAXVEC = 24;
constexpr int FC_MAXXCH = 12;                  // exchanges (one single-use mbarrier each)
constexpr int FC_SL = FC_TM * 33;               // floats of an fp32 slice buffer [16][33]

struct FcChunk {           // CTA r loads rows = clamp(rows_total - r * rows_per_rank, 0, 32) rows of K bf16 from base + r * rank_stride
  const __nv_bfloat16 *base;
  long long rank_stride;   // elements
  int ld;                  // elements between rows
  int rows_total, rows_per_rank;
  int pad_;
};
struct FcVec {             // 32 floats from p + r * rank_stride (element j valid iff r * rank_stride + j < total)
  const float *p;
  int rank_stride, total;
};
struct FcCommon {
  FcChunk chunk[FC_MAXCHUNK];
  FcVec vec[FC_MAXVEC];
  int nchunks, nslot, nvec;
  int N, B;                // kernels per frame, frames
  const __nv_bfloat16 *pack;   // optional (VknHeadW.fc_pack): this kernel's chunks re-laid [rank][chunk][32][FC_LD], zero padded
  int handoff_stores;      // 1: plane hand-offs as 16-byte st.async stores instead of one bulk copy per destination (VKN_FC_BULK=0)
  unsigned long long *dbg; // optional: 32 phase timestamps per CTA (vkn_debug_timestamps, tools/frame_chain_timeline.py); null in production
};
struct FcParamsA {
  FcCommon c;
  const float *xp0, *cnt, *pf;     // pooled sums [P][C], hard-mask pixel counts [P], proposal_feat [P][C]
  const float *x_feat_in;          // optional [P][C]: a pooled feature computed earlier (video head); xp0 / cnt are then unused
  float *x_feat_out;               // optional [P][C]
  float *obj0;                     // [P][C]: relu(fc_norm(fc_layer(.))) = the attention's identity rows
  float *qkv;                      // [P][3C]
};
struct FcParamsB {
  FcCommon c;
  const float *qkv, *obj0;
  float *obj_out;                  // [P][C]
  float *cls_out;                  // [P][ncls] or null
  float *a_ext;                    // [P][lda]: folded kernels, column C = folded bias
  __nv_bfloat16 *a_split;          // [3][B][Npad][C] bf16 planes of a_ext[:, :C]
  int ncls, lda, Npad;
  float scale;
};

// vector slots
enum { AV_FT_B = 0, AV_INP_B0, AV_INP_B1, AV_DYN_B0, AV_DYN_B1, AV_IG_B, AV_UG_B, AV_NIN_G, AV_NIN_B, AV_NOUT_G, AV_NOUT_B,
       AV_ININ_G, AV_ININ_B, AV_INOUT_G, AV_INOUT_B, AV_FC_B, AV_FCN_G, AV_FCN_B, AV_Q_B, AV_K_B, AV_V_B, AV_COUNT };
enum { BV_OUT_B = 0, BV_AN_G, BV_AN_B, BV_B1 /* 8 slots */, BV_B2 = BV_B1 + 8, BV_FN_G, BV_FN_B, BV_CLN_G, BV_CLN_B, BV_MLN_G,
       BV_MLN_B, BV_FCM_B, BV_FCC_B, BV_COUNT };
static_assert(AV_COUNT <= FC_MAXVEC && BV_COUNT <= FC_MAXVEC, "vector table too small");

// ---- shared-memory maps (bytes) ----------------------------------------------------------------------------------------
// both kernels: [ring | buf0 | buf1 | buf2 | slices | staging x2 | LN stats | vectors | mbarriers]
constexpr int FC_NSL_A = 5, FC_NSL_B = 3;
constexpr int FC_STAGE_B = FC_BLK;                                 // staging of one slice block (three of them: see fc_bcast_planes)
constexpr int FC_STATS_B = 4 * FC_CL * FC_TM * 8;                  // [4 LNs][8 ranks][16 rows] float2
constexpr int FC_VEC_B = FC_MAXVEC * 32 * 4;
constexpr size_t fc_smem_bytes(int nslot, int nslices) {
  return (size_t)nslot * FC_CHUNK_B + 3 * FC_ABUF + (size_t)nslices * FC_SL * 4 + 3 * FC_STAGE_B + FC_STATS_B + FC_VEC_B + 8 * (FC_MAXSLOT + FC_MAXXCH);
}
constexpr int FC_NSLOT_A = 7, FC_NSLOT_B = 8;
static_assert(fc_smem_bytes(FC_NSLOT_A, FC_NSL_A) <= 232448 && fc_smem_bytes(FC_NSLOT_B, FC_NSL_B) <= 232448, "over the 227 KB limit");

// ---- PTX ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fc_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void fc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fc_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void fc_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t fc_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// remote stores that report their bytes to an mbarrier of the DESTINATION CTA: the receiver waits for "all bytes of this
// exchange have landed" on its own barrier -- no cluster-wide barrier per hand-off
__device__ __forceinline__ void fc_st_remote_v4(uint32_t addr, const uint4 &v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fc_st_remote_f2(uint32_t addr, float a, float b, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(addr),
               "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fc_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fc_ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void fc_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// bounded mbarrier wait as ONE shared body (tc.cuh's inlined spin loop is unrolled by the compiler: ~40 instructions per site,
// ~60 sites per kernel)
__device__ __noinline__ void fc_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t it = 0; it < SPIN_LIMIT; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ float fc_lds(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 fc_lds2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void fc_sts(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

__device__ __forceinline__ unsigned long long fc_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define FC_TS(slot)                                                                                                       \
  do {                                                                                                                    \
    if (c.dbg != nullptr && threadIdx.x == 0)                                                                             \
      c.dbg[(((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 32 + (slot)] = fc_gtime();         \
  } while (0)

// ---- per-thread context -------------------------------------------------------------------------------------------------
struct FcCtx {
  int tid, warp, lane;
  int row, cp;             // transform mapping: thread -> (row = tid / 16, column pair cp = tid % 16) of the CTA's [16 x 32] slice
  uint32_t rank;
  uint32_t ring, buf[3], sl, stage[3], stats, vecs, bars, xbar;    // shared-memory addresses
  int nslot, nchunks, issued;
  bool stores;             // plane hand-offs as st.async stores (see fc_bcast_send)
  int row0, nvalid;        // first global row of the tile, valid rows in it
};

__device__ __forceinline__ void fc_setup(FcCtx &x, const FcCommon &c, uint8_t *smem, int nslices) {
  x.tid = threadIdx.x;
  x.warp = x.tid >> 5;
  x.lane = x.tid & 31;
  x.row = x.tid >> 4;
  x.cp = x.tid & 15;
  x.rank = fc_cluster_rank();
  x.nslot = c.nslot;
  x.nchunks = c.nchunks;
  x.issued = 0;
  x.stores = c.handoff_stores != 0;
  uint32_t a = smem_u32(smem);
  x.ring = a;                 a += (uint32_t)c.nslot * FC_CHUNK_B;
  for (int i = 0; i < 3; ++i) { x.buf[i] = a; a += FC_ABUF; }
  x.sl = a;                   a += (uint32_t)nslices * FC_SL * 4;
  x.stage[0] = a;             a += FC_STAGE_B;
  x.stage[1] = a;             a += FC_STAGE_B;
  x.stage[2] = a;             a += FC_STAGE_B;
  x.stats = a;                a += FC_STATS_B;
  x.vecs = a;                 a += FC_VEC_B;
  x.bars = a;                 a += 8 * FC_MAXSLOT;
  x.xbar = a;
  const int tile = blockIdx.y, b = blockIdx.z;
  x.row0 = b * c.N + tile * FC_TM;
  x.nvalid = min(FC_TM, c.N - tile * FC_TM);
}

// weight ring: chunk k lives in slot k % nslot; warp 0 issues (one 512-byte bulk copy per row and lane), everybody waits on the
// slot's mbarrier.  fc_ring_fill(consumed) may only be called after a __syncthreads that follows the last read of chunk consumed-1.
__device__ __forceinline__ void fc_ring_issue(const FcCommon *cp, int first, int upto, uint32_t rank, int nslot, int nchunks,
                                           uint32_t ring, uint32_t bars, int lane) {      // warp 0 only
  const FcCommon &c = *cp;
#pragma unroll 1
  for (int k = first; k < upto; ++k) {
    const FcChunk &ch = c.chunk[k];
    int rows = ch.rows_total - (int)rank * ch.rows_per_rank;
    rows = max(0, min(rows, min(ch.rows_per_rank, 32)));
    const uint32_t slot = (uint32_t)(k % nslot);
    const uint32_t bar = bars + 8u * slot;
    if (c.pack != nullptr) {        // host-prepared chunk image: ONE bulk copy
      if (lane == 0) {
        mbar_expect_tx(bar, rows > 0 ? (uint32_t)FC_CHUNK_B : 0u);
        if (rows > 0) fc_bulk_g2s(ring + slot * FC_CHUNK_B, c.pack + ((size_t)rank * nchunks + k) * (FC_CHUNK_B / 2), FC_CHUNK_B, bar);
      }
    } else {
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)rows * (FC_K * 2));
      __syncwarp();
      if (lane < rows)
        fc_bulk_g2s(ring + slot * FC_CHUNK_B + (uint32_t)lane * (FC_LD * 2),
                    ch.base + (long long)rank * ch.rank_stride + (long long)lane * ch.ld, FC_K * 2, bar);
    }
  }
}
__device__ __forceinline__ void fc_ring_fill(FcCtx &x, const FcCommon &c, int consumed) {
  const int upto = min(x.nchunks, consumed + x.nslot);
  if (x.issued < upto) {
    if (x.warp == 0) fc_ring_issue(&c, x.issued, upto, x.rank, x.nslot, x.nchunks, x.ring, x.bars, x.lane);
    x.issued = upto;
  }
}
__device__ __forceinline__ uint32_t fc_chunk(const FcCtx &x, int k) {
  const int slot = k % x.nslot;
  fc_mbar_wait(x.bars + 8u * (uint32_t)slot, (uint32_t)(k / x.nslot) & 1u);
  return x.ring + (uint32_t)slot * FC_CHUNK_B;
}

__device__ __forceinline__ void fc_load_vecs(const FcCtx &x, const FcCommon &c) {
  for (int v = x.warp; v < c.nvec; v += FC_NT / 32) {
    const FcVec &e = c.vec[v];
    const int off = (int)x.rank * e.rank_stride + x.lane;
    const float val = (e.p != nullptr && off < e.total) ? __ldg(e.p + off) : 0.f;
    fc_sts(x.vecs + (uint32_t)(v * 32 + x.lane) * 4u, val);
  }
}
__device__ __forceinline__ float2 fc_vec2(const FcCtx &x, int v, int col) { return fc_lds2(x.vecs + (uint32_t)(v * 32 + col) * 4u); }

// D[16 x 8] = sum over the three planes of A_pl[16 x 256] . W[8 rows][256]^T    (one independent accumulation chain per plane)
// (__noinline__: the two kernels are straight-line programs of ~10^4 instructions that every warp executes once -- shared
// bodies keep the hot loops in the instruction cache; ncu showed "no instruction" as the top stall with everything inlined)
struct FcFrag { float d[4]; };
struct FcFrag2 { float d0[4], d1[4]; };
__device__ __noinline__ FcFrag fc_mma_tile(uint32_t a_buf, uint32_t w_rows, int lane) {
  FcFrag out;
  float acc[3][4];
#pragma unroll
  for (int pl = 0; pl < 3; ++pl)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[pl][e] = 0.f;
  // lane's ldmatrix row address inside a slice block, for the two 16-wide k-steps of a 32-column block
  const int arow = lane & 15, aswz = (arow >> 1) & 3;
  const uint32_t a_row = a_buf + (uint32_t)(arow * (FC_SW * 2));
  const uint32_t a_h0 = a_row + (uint32_t)((((lane >> 4)) ^ aswz) << 4), a_h1 = a_row + (uint32_t)(((2 + (lane >> 4)) ^ aswz) << 4);
  const uint32_t b_addr = w_rows + (uint32_t)((lane & 7) * FC_LD + (lane >> 3) * 8) * 2u;
#pragma unroll
  for (int k0 = 0; k0 < FC_K; k0 += 32) {      // fully unrolled: all ldmatrix of a tile can be in flight (A/B: 213.8 -> 209.8 us per frame)
    uint32_t b[4];
    fc_ldsm_x4(b, b_addr + (uint32_t)k0 * 2u);
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        uint32_t a[4];
        fc_ldsm_x4(a, (h ? a_h1 : a_h0) + (uint32_t)((k0 >> 5) * FC_BLK + pl * (FC_TM * FC_SW * 2)));
        fc_mma(acc[pl], a, b[2 * h], b[2 * h + 1]);
      }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) out.d[e] = (acc[2][e] + acc[1][e]) + acc[0][e];      // small terms first
  return out;
}

// two n8 tiles that share the A fragments (the FFN phases: 16 tiles per phase -> every warp owns a pair; the A operand is
// what the shared-memory pipe is busy with)
__device__ __noinline__ FcFrag2 fc_mma_tile2(uint32_t a_buf, uint32_t w_rows0, uint32_t w_rows1, int lane) {
  FcFrag2 out;
  float acc[2][3][4];
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int pl = 0; pl < 3; ++pl)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[t][pl][e] = 0.f;
  // lane's ldmatrix row address inside a slice block, for the two 16-wide k-steps of a 32-column block
  const int arow = lane & 15, aswz = (arow >> 1) & 3;
  const uint32_t a_row = a_buf + (uint32_t)(arow * (FC_SW * 2));
  const uint32_t a_h0 = a_row + (uint32_t)((((lane >> 4)) ^ aswz) << 4), a_h1 = a_row + (uint32_t)(((2 + (lane >> 4)) ^ aswz) << 4);
  const uint32_t b_off = (uint32_t)((lane & 7) * FC_LD + (lane >> 3) * 8) * 2u;
#pragma unroll
  for (int k0 = 0; k0 < FC_K; k0 += 32) {      // fully unrolled: all ldmatrix of a tile can be in flight (A/B: 213.8 -> 209.8 us per frame)
    uint32_t b0[4], b1[4];
    fc_ldsm_x4(b0, w_rows0 + b_off + (uint32_t)k0 * 2u);
    fc_ldsm_x4(b1, w_rows1 + b_off + (uint32_t)k0 * 2u);
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        uint32_t a[4];
        fc_ldsm_x4(a, (h ? a_h1 : a_h0) + (uint32_t)((k0 >> 5) * FC_BLK + pl * (FC_TM * FC_SW * 2)));
        fc_mma(acc[0][pl], a, b0[2 * h], b0[2 * h + 1]);
        fc_mma(acc[1][pl], a, b1[2 * h], b1[2 * h + 1]);
      }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    out.d0[e] = (acc[0][2][e] + acc[0][1][e]) + acc[0][0][e];
    out.d1[e] = (acc[1][2][e] + acc[1][1][e]) + acc[1][0][e];
  }
  return out;
}
// GEMM phase over FOUR full chunks [k0, k0 + 4) with one A operand: warp w owns tile (w & 3) of chunks k0 + (w >> 2) and + 2
template <typename Epi>
__device__ __forceinline__ void fc_gemm2(const FcCtx &x, const FcCommon &, int k0, uint32_t a_buf, Epi epi) {
  const int kc0 = k0 + (x.warp >> 2), kc1 = kc0 + 2, j = x.warp & 3;
  const uint32_t w0 = fc_chunk(x, kc0), w1 = fc_chunk(x, kc1);
  const FcFrag2 f = fc_mma_tile2(a_buf, w0 + (uint32_t)(j * 8 * FC_LD) * 2u, w1 + (uint32_t)(j * 8 * FC_LD) * 2u, x.lane);
  epi(kc0, j, f.d0);
  epi(kc1, j, f.d1);
}

// GEMM phase over the ring chunks [k0, k0 + nk): n8-tile items round-robin over the 8 warps.
//   a_of(kc)          -> shared address of the A planes chunk kc multiplies
//   epi(kc, j, d)     -> consumes the fragment of tile j (rows g / g + 8, columns 8 j + 2 t4, +1) of chunk kc
template <typename AOf, typename Epi>
__device__ __forceinline__ void fc_gemm(const FcCtx &x, const FcCommon &c, int k0, int nk, AOf a_of, Epi epi) {
  for (int it = x.warp; it < nk * 4; it += FC_NT / 32) {
    const int kc = k0 + (it >> 2), j = it & 3;
    const FcChunk &ch = c.chunk[kc];
    int rows = ch.rows_total - (int)x.rank * ch.rows_per_rank;
    rows = min(rows, min(ch.rows_per_rank, 32));
    if (j * 8 >= rows) continue;                       // no live weight row in this tile (warp-uniform)
    const uint32_t w = fc_chunk(x, kc);
    const FcFrag f = fc_mma_tile(a_of(kc), w + (uint32_t)(j * 8 * FC_LD) * 2u, x.lane);
    epi(kc, j, f.d);
  }
}
// fragment -> fp32 slice buffer [16][33]
__device__ __forceinline__ void fc_frag_to_slice(uint32_t sl, int j, int lane, const float (&d)[4]) {
  const int g = lane >> 2, c = j * 8 + 2 * (lane & 3);
  fc_sts(sl + (uint32_t)(g * 33 + c) * 4u, d[0]);
  fc_sts(sl + (uint32_t)(g * 33 + c + 1) * 4u, d[1]);
  fc_sts(sl + (uint32_t)((g + 8) * 33 + c) * 4u, d[2]);
  fc_sts(sl + (uint32_t)((g + 8) * 33 + c + 1) * 4u, d[3]);
}
__device__ __forceinline__ float2 fc_slice2(const FcCtx &x, uint32_t sl) {
  float2 v;
  v.x = fc_lds(sl + (uint32_t)(x.row * 33 + 2 * x.cp) * 4u);
  v.y = fc_lds(sl + (uint32_t)(x.row * 33 + 2 * x.cp + 1) * 4u);
  return v;
}
__device__ __forceinline__ void fc_slice2_store(const FcCtx &x, uint32_t sl, float a, float b) {
  fc_sts(sl + (uint32_t)(x.row * 33 + 2 * x.cp) * 4u, a);
  fc_sts(sl + (uint32_t)(x.row * 33 + 2 * x.cp + 1) * 4u, b);
}

// staged slice block -> the same block of the A buffer of CTA `to` (one warp, one lane, per destination): a single bulk copy
// shared::cta -> shared::cluster that reports its 3072 bytes to the destination's barrier
__device__ __forceinline__ void fc_bulk_s2d(uint32_t dst_remote, uint32_t src, uint32_t bytes, uint32_t bar_remote) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_remote),
               "r"(src), "r"(bytes), "r"(bar_remote)
               : "memory");
}
// `stores`: the same block as 192 16-byte st.async stores (6 per lane) -- the transport of the first version, 13 us per frame
// slower; kept selectable (VKN_FC_BULK=0) because compute-sanitizer's memcheck, clean on these stores, reports every bulk copy
// shared::cta -> shared::cluster as "not located in remote CTA" although both transports use the same mapa addresses and
// produce bit-identical results (tests/test_gpu_parity.py::test_frame_chain_handoff_transports_agree)
__device__ __forceinline__ void fc_bcast_send(uint32_t st, uint32_t dst_block, uint32_t bar_local, uint32_t to, int lane, bool stores) {
  if (!stores) fence_proxy_async();   // the staging writes (generic proxy) are read by the copy engine (async proxy)
  __syncthreads();
  const uint32_t dst = fc_mapa(dst_block, to), bar = fc_mapa(bar_local, to);
  if (stores) {
#pragma unroll
    for (int i = 0; i < FC_BLK / 16 / 32; ++i) {
      const uint32_t o = (uint32_t)(lane + 32 * i) * 16u;
      fc_st_remote_v4(dst + o, lds_u4(st + o), bar);
    }
  } else if (lane == 0) {
    fc_bulk_s2d(dst, st, FC_BLK, bar);
  }
}
// the owner's transformed slice -> bf16 planes in the A buffer `dst` of ALL 8 CTAs (warp w serves CTA w)
// Staging buffers: a buffer may be rewritten only when every destination has received the block last sent from it.  With the
// assignment used by the kernels (A: 0 1 0 1; B: 0 1 0 1 2 0) a CTA has, between two uses of a buffer, waited for data that each
// peer sent only after receiving that earlier block.
__device__ __forceinline__ void fc_bcast_planes(const FcCtx &x, int stg, uint32_t dst, int xch, float v0, float v1) {
  uint32_t w3[3];
  split3_pair(v0, v1, w3[0], w3[1], w3[2]);
  const uint32_t st = x.stage[stg];
  const uint32_t o = (uint32_t)(x.row * (FC_SW * 2) + ((((x.cp >> 2)) ^ ((x.row >> 1) & 3)) << 4) + (x.cp & 3) * 4);
#pragma unroll
  for (int pl = 0; pl < 3; ++pl) sts_u32(st + (uint32_t)(pl * (FC_TM * FC_SW * 2)) + o, w3[pl]);
  fc_bcast_send(st, dst + (uint32_t)((int)x.rank * FC_BLK), x.xbar + 8u * (uint32_t)xch, (uint32_t)x.warp, x.lane, x.stores);
}
__device__ __forceinline__ void fc_xwait(const FcCtx &x, int xch) { fc_mbar_wait(x.xbar + 8u * (uint32_t)xch, 0u); }

// LayerNorm statistics of NLN row vectors whose 256 columns are spread over the 8 CTAs: per-slice (mean, M2) of the own 32
// columns -> every CTA's table (st.async, completion on the receiver's barrier `xch`) -> merged mean / rstd.
template <int NLN>
__device__ __forceinline__ void fc_stats_send(const FcCtx &x, int xch, const float2 (&v)[NLN]) {
  float m[NLN], q[NLN];
#pragma unroll
  for (int t = 0; t < NLN; ++t) m[t] = v[t].x + v[t].y;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1)
#pragma unroll
    for (int t = 0; t < NLN; ++t) m[t] += __shfl_xor_sync(0xffffffffu, m[t], o);
#pragma unroll
  for (int t = 0; t < NLN; ++t) {
    m[t] *= (1.0f / FC_SW);
    const float d0 = v[t].x - m[t], d1 = v[t].y - m[t];
    q[t] = d0 * d0 + d1 * d1;
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1)
#pragma unroll
    for (int t = 0; t < NLN; ++t) q[t] += __shfl_xor_sync(0xffffffffu, q[t], o);
  if (x.cp < FC_CL) {          // thread cp of the row's half-warp serves CTA cp
    const uint32_t bar = fc_mapa(x.xbar + 8u * (uint32_t)xch, (uint32_t)x.cp);
#pragma unroll
    for (int t = 0; t < NLN; ++t) {
      const uint32_t a = x.stats + (uint32_t)((t * FC_CL + (int)x.rank) * FC_TM + x.row) * 8u;
      fc_st_remote_f2(fc_mapa(a, (uint32_t)x.cp), m[t], q[t], bar);
    }
  }
  fc_xwait(x, xch);
}
__device__ __forceinline__ float2 fc_stats_merge_(uint32_t stats_row /* &stats[t][0][row] */) {
  float m[FC_CL], s = 0.f, q = 0.f;
#pragma unroll
  for (int i = 0; i < FC_CL; ++i) {
    const float2 e = fc_lds2(stats_row + (uint32_t)(i * FC_TM) * 8u);
    m[i] = e.x;
    s += e.x;
    q += e.y;
  }
  const float mean = s * (1.0f / FC_CL);
  float dd = 0.f;
#pragma unroll
  for (int i = 0; i < FC_CL; ++i) {
    const float d = m[i] - mean;
    dd = fmaf(d, d, dd);
  }
  q = fmaf((float)FC_SW, dd, q);
  return make_float2(mean, 1.0f / sqrtf(q * (1.0f / (FC_SW * FC_CL)) + 1e-5f));
}
__device__ __forceinline__ void fc_stats_merge(const FcCtx &x, int t, float &mean, float &rstd) {
  const float2 r = fc_stats_merge_(x.stats + (uint32_t)(t * FC_CL * FC_TM + x.row) * 8u);
  mean = r.x;
  rstd = r.y;
}
__device__ __forceinline__ float2 fc_ln_apply(const FcCtx &x, float2 v, float mean, float rstd, int vg, int vb) {
  const float2 g = fc_vec2(x, vg, 2 * x.cp), b = fc_vec2(x, vb, 2 * x.cp);
  return make_float2((v.x - mean) * rstd * g.x + b.x, (v.y - mean) * rstd * g.y + b.y);
}

// fp32 rows [16][256] from global -> three bf16 planes of an A buffer (local; warp w owns rows 2 w, 2 w + 1)
__device__ __forceinline__ void fc_rows_to_planes(const FcCtx &x, const float *src, int ld, uint32_t dst) {
  float2 t[2][4];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int r = 2 * x.warp + q;
#pragma unroll
    for (int p = 0; p < 4; ++p)
      t[q][p] = r < x.nvalid ? __ldcg(reinterpret_cast<const float2 *>(src + (size_t)(x.row0 + r) * ld + 64 * p + 2 * x.lane))
                             : make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int r = 2 * x.warp + q;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      uint32_t w3[3];
      split3_pair(t[q][p].x, t[q][p].y, w3[0], w3[1], w3[2]);
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) sts_u32(dst + fc_a_off(pl, r, 64 * p + 2 * x.lane), w3[pl]);
    }
  }
}

// xbytes[e] = bytes exchange e delivers to every CTA (0 = unused): its barrier is armed here, before the cluster barrier
// that precedes the first remote store, and used exactly once (phase 0) -- no parity bookkeeping
template <int NX>
__device__ __forceinline__ void fc_prologue(FcCtx &x, const FcCommon &c, const uint32_t (&xbytes)[NX]) {
  static_assert(NX <= FC_MAXXCH, "too many exchanges");
  if (x.tid == 0) {
    for (int i = 0; i < FC_MAXSLOT; ++i) mbar_init(x.bars + 8u * i, 1);
#pragma unroll
    for (int e = 0; e < NX; ++e) mbar_init(x.xbar + 8u * e, 1);
    fence_barrier_init();
#pragma unroll
    for (int e = 0; e < NX; ++e)
      if (xbytes[e] != 0) mbar_expect_tx(x.xbar + 8u * e, xbytes[e]);
  }
  __syncthreads();
  FC_TS(0);
  fc_ring_fill(x, c, 0);          // weights: nothing of the chain writes them -> before the dependency wait
  fc_load_vecs(x, c);
  FC_TS(1);
  fc_cluster_sync();              // every CTA of the cluster runs before the first remote store
  FC_TS(2);
  pdl_wait();
  FC_TS(3);
  pdl_trigger();                  // the next kernel's CTAs go to other SMs and prefetch their weights meanwhile
}

// ======================================================================================================================
// kernel A: x_feat = pooled . ft_w^T + cnt (x) ft_b;  KernelUpdator;  obj0 = relu(fc_norm(.));  q / k / v = in_proj(obj0)
// ring chunks: 0 ft_w | 1, 2 input_layer (in, out) | 3, 4 dynamic_layer (in, out) | 5 input_gate | 6 update_gate | 7 fc_layer |
//              8, 9, 10 in_proj (q, k, v)
// ======================================================================================================================
enum { XA_XF = 0, XA_GF, XA_ST3, XA_FEAT, XA_ST4, XA_OBJ, XA_COUNT };
constexpr uint32_t FC_XPL = FC_CL * FC_STAGE_B;                    // bytes a plane hand-off delivers to every CTA
constexpr uint32_t FC_XST = FC_CL * FC_TM * 8;                     // ... a LayerNorm statistics exchange, per LayerNorm

__global__ void __launch_bounds__(FC_NT, 1) vkn_frame_chain_a_kernel(const __grid_constant__ FcParamsA P) {
  extern __shared__ __align__(128) uint8_t fc_smem[];
  const FcCommon &c = P.c;
  FcCtx x;
  fc_setup(x, c, fc_smem, FC_NSL_A);
  const uint32_t S_XF = x.sl, S_II = x.sl + FC_SL * 4, S_IO = x.sl + 2 * FC_SL * 4, S_PI = x.sl + 3 * FC_SL * 4,
                 S_PO = x.sl + 4 * FC_SL * 4;
  const uint32_t S_IG = S_XF, S_UG = S_II, S_FC = S_PI;     // reuse once the first tenants are dead
  const bool have_xf = P.x_feat_in != nullptr;       // chunk 0 is then empty and nobody hands x_feat around
  {
    const uint32_t xb[XA_COUNT] = {have_xf ? 0u : FC_XPL, FC_XPL, 4 * FC_XST, FC_XPL, FC_XST, FC_XPL};
    fc_prologue(x, c, xb);
  }

  const bool live = x.row < x.nvalid;
  const size_t grow = (size_t)(x.row0 + x.row);
  const int gcol = FC_SW * (int)x.rank + 2 * x.cp;
  const float cnt = (live && !have_xf) ? __ldcg(P.cnt + grow) : 0.f;

  // ---- phase 1: x_feat (A = pooled sums) and input_layer (A = proposal_feat)
  if (have_xf) fc_rows_to_planes(x, P.x_feat_in, FC_K, x.buf[1]);
  else fc_rows_to_planes(x, P.xp0, FC_K, x.buf[0]);
  fc_rows_to_planes(x, P.pf, FC_K, x.buf[2]);
  __syncthreads();
  FC_TS(4);
  // x_feat first: its hand-off travels while input_layer (which does not need it) is multiplied
  if (!have_xf) {
    fc_gemm(x, c, 0, 1, [&](int) { return x.buf[0]; }, [&](int, int j, const float (&d)[4]) { fc_frag_to_slice(S_XF, j, x.lane, d); });
    __syncthreads();
    float2 v = fc_slice2(x, S_XF);
    const float2 b = fc_vec2(x, AV_FT_B, 2 * x.cp);
    v.x = fmaf(cnt, b.x, v.x);
    v.y = fmaf(cnt, b.y, v.y);
    if (P.x_feat_out != nullptr && live) *reinterpret_cast<float2 *>(P.x_feat_out + grow * FC_K + gcol) = v;
    fc_bcast_planes(x, 0, x.buf[1], XA_XF, v.x, v.y);
  }
  FC_TS(5);
  fc_gemm(x, c, 1, 2, [&](int) { return x.buf[2]; },
          [&](int kc, int j, const float (&d)[4]) { fc_frag_to_slice(kc == 1 ? S_II : S_IO, j, x.lane, d); });
  __syncthreads();
  fc_ring_fill(x, c, 3);
  if (!have_xf) fc_xwait(x, XA_XF);
  FC_TS(6);

  // ---- phase 2: dynamic_layer (A = x_feat) -> param_in, param_out;  gate_feats = input_in * param_in
  fc_gemm(x, c, 3, 2, [&](int) { return x.buf[1]; },
          [&](int kc, int j, const float (&d)[4]) { fc_frag_to_slice(kc == 3 ? S_PI : S_PO, j, x.lane, d); });
  __syncthreads();
  FC_TS(7);
  fc_ring_fill(x, c, 5);
  {
    const float2 a = fc_slice2(x, S_II), b = fc_slice2(x, S_PI);
    const float2 ba = fc_vec2(x, AV_INP_B0, 2 * x.cp), bb = fc_vec2(x, AV_DYN_B0, 2 * x.cp);
    fc_bcast_planes(x, 1, x.buf[0], XA_GF, (a.x + ba.x) * (b.x + bb.x), (a.y + ba.y) * (b.y + bb.y));      // kernel_updator.py:70
    fc_xwait(x, XA_GF);
  }
  FC_TS(8);

  // ---- phase 3: input_gate / update_gate (A = gate_feats), the gate (4 LayerNorms, 2 sigmoids)
  fc_gemm(x, c, 5, 2, [&](int) { return x.buf[0]; },
          [&](int kc, int j, const float (&d)[4]) { fc_frag_to_slice(kc == 5 ? S_IG : S_UG, j, x.lane, d); });
  __syncthreads();
  FC_TS(9);
  fc_ring_fill(x, c, 7);
  {
    float2 v[4];
    const float2 big = fc_vec2(x, AV_IG_B, 2 * x.cp), bug = fc_vec2(x, AV_UG_B, 2 * x.cp);
    const float2 bpo = fc_vec2(x, AV_DYN_B1, 2 * x.cp), bio = fc_vec2(x, AV_INP_B1, 2 * x.cp);
    v[0] = fc_slice2(x, S_UG);  v[0].x += bug.x;  v[0].y += bug.y;       // update gate pre-activation   (norm_in)
    v[1] = fc_slice2(x, S_PO);  v[1].x += bpo.x;  v[1].y += bpo.y;       // param_out                    (norm_out)
    v[2] = fc_slice2(x, S_IG);  v[2].x += big.x;  v[2].y += big.y;       // input gate pre-activation    (input_norm_in)
    v[3] = fc_slice2(x, S_IO);  v[3].x += bio.x;  v[3].y += bio.y;       // input_out                    (input_norm_out)
    fc_stats_send<4>(x, XA_ST3, v);
    FC_TS(10);
    float mean, rstd;
    fc_stats_merge(x, 0, mean, rstd);  v[0] = fc_ln_apply(x, v[0], mean, rstd, AV_NIN_G, AV_NIN_B);
    fc_stats_merge(x, 1, mean, rstd);  v[1] = fc_ln_apply(x, v[1], mean, rstd, AV_NOUT_G, AV_NOUT_B);
    fc_stats_merge(x, 2, mean, rstd);  v[2] = fc_ln_apply(x, v[2], mean, rstd, AV_ININ_G, AV_ININ_B);
    fc_stats_merge(x, 3, mean, rstd);  v[3] = fc_ln_apply(x, v[3], mean, rstd, AV_INOUT_G, AV_INOUT_B);
    const float f0 = sigmoidf_(v[0].x) * v[1].x + sigmoidf_(v[2].x) * v[3].x;       // kernel_updator.py:76-88
    const float f1 = sigmoidf_(v[0].y) * v[1].y + sigmoidf_(v[2].y) * v[3].y;
    fc_bcast_planes(x, 0, x.buf[1], XA_FEAT, f0, f1);
    fc_xwait(x, XA_FEAT);
  }
  FC_TS(11);

  // ---- phase 4: fc_layer (A = features) -> fc_norm -> ReLU = obj0
  fc_gemm(x, c, 7, 1, [&](int) { return x.buf[1]; }, [&](int, int j, const float (&d)[4]) { fc_frag_to_slice(S_FC, j, x.lane, d); });
  __syncthreads();
  FC_TS(12);
  fc_ring_fill(x, c, 8);
  {
    float2 v[1];
    const float2 b = fc_vec2(x, AV_FC_B, 2 * x.cp);
    v[0] = fc_slice2(x, S_FC);
    v[0].x += b.x;
    v[0].y += b.y;
    fc_stats_send<1>(x, XA_ST4, v);
    FC_TS(13);
    float mean, rstd;
    fc_stats_merge(x, 0, mean, rstd);
    float2 o = fc_ln_apply(x, v[0], mean, rstd, AV_FCN_G, AV_FCN_B);                  // kernel_updator.py:91-92
    o.x = fmaxf(o.x, 0.f);
    o.y = fmaxf(o.y, 0.f);
    if (live) *reinterpret_cast<float2 *>(P.obj0 + grow * FC_K + gcol) = o;
    fc_bcast_planes(x, 1, x.buf[0], XA_OBJ, o.x, o.y);
    fc_cluster_arrive();            // last remote store of this CTA issued
    fc_xwait(x, XA_OBJ);
  }
  FC_TS(14);

  // ---- phase 5: in_proj (A = obj0) -> q, k, v rows in global memory (the attention of kernel B reads every tile's k / v)
  fc_gemm(x, c, 8, 3, [&](int) { return x.buf[0]; }, [&](int kc, int j, const float (&d)[4]) {
    const int g = x.lane >> 2, cc = j * 8 + 2 * (x.lane & 3);
    const float2 b = fc_vec2(x, AV_Q_B + (kc - 8), cc);
    float *dst = P.qkv + (size_t)x.row0 * (3 * FC_K) + (kc - 8) * FC_K + FC_SW * (int)x.rank + cc;
    if (g < x.nvalid) *reinterpret_cast<float2 *>(dst + (size_t)g * (3 * FC_K)) = make_float2(d[0] + b.x, d[1] + b.y);
    if (g + 8 < x.nvalid) *reinterpret_cast<float2 *>(dst + (size_t)(g + 8) * (3 * FC_K)) = make_float2(d[2] + b.x, d[3] + b.y);
  });
  FC_TS(30);
  fc_cluster_wait();                // no CTA of the cluster leaves while a peer may still address its shared memory
}

// ======================================================================================================================
// kernel B: attention (head r on CTA r), out_proj + residual + attention_norm, FFN + ffn_norm, cls / mask FC + LN + ReLU,
//           fc_cls, fc_mask, fold (mask kernels x feat_transform) -> a_ext + the mask conv's operand planes
// ring chunks: 0 out_proj | 1..8 ffn.w1 (hidden columns 256 r + 32 i) | 9..16 ffn.w2 (output rows 32 i, K slice 256 r) |
//              17 cls_fc | 18 mask_fc | 19 fc_mask | 20 fc_cls | 21 fold | 22 fold bias row (CTA 0)
// ======================================================================================================================
enum { XB_ATT = 0, XB_ST1, XB_O1, XB_PART, XB_ST3, XB_OBJ, XB_ST4, XB_CLS, XB_MASK, XB_MK, XB_COUNT };

__global__ void __launch_bounds__(FC_NT, 1) vkn_frame_chain_b_kernel(const __grid_constant__ FcParamsB P) {
  extern __shared__ __align__(128) uint8_t fc_smem[];
  const FcCommon &c = P.c;
  FcCtx x;
  fc_setup(x, c, fc_smem, FC_NSL_B);
  const uint32_t S0 = x.sl, S1 = x.sl + FC_SL * 4, S2 = x.sl + 2 * FC_SL * 4;
  const uint32_t BUF_A = x.buf[0], BUF_B = x.buf[1], BUF_H = x.buf[2];
  const bool with_cls = P.cls_out != nullptr;
  {
    const uint32_t xb[XB_COUNT] = {FC_XPL, FC_XST, FC_XPL, (uint32_t)(FC_CL * FC_TM * FC_SW * 4), FC_XST, FC_XPL, 2 * FC_XST,
                                   with_cls ? FC_XPL : 0u, FC_XPL, FC_XPL};
    fc_prologue(x, c, xb);
  }

  const bool live = x.row < x.nvalid;
  const size_t grow = (size_t)(x.row0 + x.row);
  const int gcol = FC_SW * (int)x.rank + 2 * x.cp;
  const int N = c.N, frame_row0 = (int)blockIdx.z * N;
  const float2 ident = live ? __ldcg(reinterpret_cast<const float2 *>(P.obj0 + grow * FC_K + gcol)) : make_float2(0.f, 0.f);

  // ---- phase 0: attention of head `rank` for the tile's queries; K / V of the head in shared memory (over BUF_B + BUF_H)
  {
    float *Ks = reinterpret_cast<float *>(fc_smem + (BUF_B - x.ring));        // [N][33]
    float *Vs = Ks + (size_t)N * 33;                                           // [N][33]
    float *Qs = reinterpret_cast<float *>(fc_smem + (S1 - x.ring));            // [16][32] (slice buffers 1-2 are idle here)
    // probabilities [8 warps][N][2]: the rest of slice buffers 1-2, both staging buffers and the LayerNorm statistics table are
    // contiguous and idle until this CTA's attention output has been handed on (12.4 KB: N <= 192)
    float *Ps = Qs + FC_TM * 32;
    static_assert(2 * FC_SL * 4 - FC_TM * 32 * 4 + 3 * FC_STAGE_B + FC_STATS_B >= (FC_NT / 32) * 2 * 192 * 4, "probability rows do not fit");
    const float *kbase = P.qkv + (size_t)frame_row0 * (3 * FC_K) + FC_K + FC_SW * (int)x.rank;
    for (int idx = x.tid; idx < N * 8; idx += FC_NT) {
      const int j = idx >> 3, d = (idx & 7) * 4;
      const float4 kk = __ldcg(reinterpret_cast<const float4 *>(kbase + (size_t)j * (3 * FC_K) + d));
      const float4 vv = __ldcg(reinterpret_cast<const float4 *>(kbase + (size_t)j * (3 * FC_K) + FC_K + d));
      float *kd = Ks + j * 33 + d, *vd = Vs + j * 33 + d;
      kd[0] = kk.x; kd[1] = kk.y; kd[2] = kk.z; kd[3] = kk.w;
      vd[0] = vv.x; vd[1] = vv.y; vd[2] = vv.z; vd[3] = vv.w;
    }
    for (int idx = x.tid; idx < FC_TM * 32; idx += FC_NT) {
      const int r = idx >> 5, d = idx & 31;
      Qs[idx] = r < x.nvalid ? __ldcg(P.qkv + (size_t)(x.row0 + r) * (3 * FC_K) + FC_SW * (int)x.rank + d) * P.scale : 0.f;
    }
    __syncthreads();
    FC_TS(4);
    // a warp carries its TWO query rows through every pass: each K / V element fetched from shared memory feeds both.
    // Compact loops on purpose: this kernel is a straight-line program every warp runs once, instruction fetch is its top stall.
    const int r0 = 2 * x.warp;
    float ov0 = 0.f, ov1 = 0.f;
    if (r0 < x.nvalid) {                   // warp-uniform (a row beyond nvalid has q = 0: harmless, not stored)
      float *ps = Ps + (size_t)x.warp * ((2 * N + 3) & ~3);    // [N][2]: scores, then probabilities, of the two rows (16-byte aligned)
      const float4 *q0 = reinterpret_cast<const float4 *>(Qs + r0 * 32), *q1 = reinterpret_cast<const float4 *>(Qs + (r0 + 1) * 32);
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll 1
      for (int j = x.lane; j < N; j += 32) {
        const float *kp = Ks + (size_t)j * 33;
        float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < 8; ++d4) {
          const float4 a = q0[d4], b = q1[d4];
          const float k0 = kp[4 * d4], k1 = kp[4 * d4 + 1], k2 = kp[4 * d4 + 2], k3 = kp[4 * d4 + 3];
          s0 = fmaf(a.x, k0, s0);  s1 = fmaf(b.x, k0, s1);
          t0 = fmaf(a.y, k1, t0);  t1 = fmaf(b.y, k1, t1);
          s0 = fmaf(a.z, k2, s0);  s1 = fmaf(b.z, k2, s1);
          t0 = fmaf(a.w, k3, t0);  t1 = fmaf(b.w, k3, t1);
        }
        s0 += t0;
        s1 += t1;
        *reinterpret_cast<float2 *>(ps + 2 * j) = make_float2(s0, s1);
        mx0 = fmaxf(mx0, s0);
        mx1 = fmaxf(mx1, s1);
      }
      mx0 = warp_max(mx0);
      mx1 = warp_max(mx1);
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll 1
      for (int j = x.lane; j < N; j += 32) {
        float2 e = *reinterpret_cast<float2 *>(ps + 2 * j);
        e.x = expf(e.x - mx0);
        e.y = expf(e.y - mx1);
        *reinterpret_cast<float2 *>(ps + 2 * j) = e;
        sum0 += e.x;
        sum1 += e.y;
      }
      sum0 = warp_sum(sum0);
      sum1 = warp_sum(sum1);
      __syncwarp();
      float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
      int j = 0;
#pragma unroll 2
      for (; j + 2 <= N; j += 2) {
        const float4 pp = *reinterpret_cast<const float4 *>(ps + 2 * j);      // (p0[j], p1[j], p0[j+1], p1[j+1])
        const float v0 = Vs[j * 33 + x.lane], v1 = Vs[(j + 1) * 33 + x.lane];
        a0 = fmaf(pp.x, v0, a0);
        b0 = fmaf(pp.y, v0, b0);
        a1 = fmaf(pp.z, v1, a1);
        b1 = fmaf(pp.w, v1, b1);
      }
      if (j < N) {
        const float2 pp = *reinterpret_cast<const float2 *>(ps + 2 * j);
        const float v0 = Vs[j * 33 + x.lane];
        a0 = fmaf(pp.x, v0, a0);
        b0 = fmaf(pp.y, v0, b0);
      }
      ov0 = (a0 + a1) / sum0;
      ov1 = (b0 + b1) / sum1;
    }
    fc_sts(S0 + (uint32_t)(r0 * 33 + x.lane) * 4u, ov0);
    fc_sts(S0 + (uint32_t)((r0 + 1) * 33 + x.lane) * 4u, ov1);
    __syncthreads();
    FC_TS(5);
    const float2 a = fc_slice2(x, S0);
    fc_bcast_planes(x, 0, BUF_A, XB_ATT, a.x, a.y);
    fc_xwait(x, XB_ATT);
  }
  FC_TS(6);

  // ---- phase 1: out_proj + bias + identity -> attention_norm = o1 (kept for the FFN residual)
  fc_gemm(x, c, 0, 1, [&](int) { return BUF_A; }, [&](int, int j, const float (&d)[4]) { fc_frag_to_slice(S1, j, x.lane, d); });
  __syncthreads();
  FC_TS(7);
  fc_ring_fill(x, c, 1);
  float2 o1;
  {
    float2 v[1];
    const float2 b = fc_vec2(x, BV_OUT_B, 2 * x.cp);
    v[0] = fc_slice2(x, S1);
    v[0].x += b.x + ident.x;
    v[0].y += b.y + ident.y;
    fc_stats_send<1>(x, XB_ST1, v);
    FC_TS(8);
    float mean, rstd;
    fc_stats_merge(x, 0, mean, rstd);
    o1 = fc_ln_apply(x, v[0], mean, rstd, BV_AN_G, BV_AN_B);
    fc_bcast_planes(x, 1, BUF_B, XB_O1, o1.x, o1.y);
    fc_xwait(x, XB_O1);
  }
  FC_TS(9);

  // ---- phase 2: FFN first Linear + ReLU: this CTA's 256 hidden channels stay local as the planes of BUF_H
  auto ffn1_epi = [&](int kc, int j, const float (&d)[4]) {
    const int g = x.lane >> 2, cc = j * 8 + 2 * (x.lane & 3), i = kc - 1;
    const float2 b = fc_vec2(x, BV_B1 + i, cc);
    uint32_t w3[3];
    split3_pair(fmaxf(d[0] + b.x, 0.f), fmaxf(d[1] + b.y, 0.f), w3[0], w3[1], w3[2]);
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) sts_u32(BUF_H + fc_a_off(pl, g, 32 * i + cc), w3[pl]);
    split3_pair(fmaxf(d[2] + b.x, 0.f), fmaxf(d[3] + b.y, 0.f), w3[0], w3[1], w3[2]);
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) sts_u32(BUF_H + fc_a_off(pl, g + 8, 32 * i + cc), w3[pl]);
  };
  fc_gemm2(x, c, 1, BUF_B, ffn1_epi);
  __syncthreads();
  FC_TS(10);
  fc_ring_fill(x, c, 5);
  fc_gemm2(x, c, 5, BUF_B, ffn1_epi);
  __syncthreads();
  FC_TS(11);
  fc_ring_fill(x, c, 9);

  // ---- phase 3: FFN second Linear over the local K slice -> partial rows, reduce-scattered to the column owners (over BUF_A)
  auto ffn2_epi = [&](int kc, int j, const float (&d)[4]) {
    const int g = x.lane >> 2, cc = j * 8 + 2 * (x.lane & 3);
    const uint32_t owner = (uint32_t)(kc - 9);
    const uint32_t dst = fc_mapa(BUF_A + (uint32_t)(((int)x.rank * FC_TM + g) * FC_SW + cc) * 4u, owner);
    const uint32_t bar = fc_mapa(x.xbar + 8u * XB_PART, owner);
    fc_st_remote_f2(dst, d[0], d[1], bar);
    fc_st_remote_f2(dst + 8u * FC_SW * 4u, d[2], d[3], bar);
  };
  fc_gemm2(x, c, 9, BUF_H, ffn2_epi);
  __syncthreads();
  FC_TS(12);
  fc_ring_fill(x, c, 13);
  fc_gemm2(x, c, 13, BUF_H, ffn2_epi);
  __syncthreads();
  FC_TS(13);
  fc_ring_fill(x, c, 17);
  fc_xwait(x, XB_PART);
  FC_TS(14);
  {
    float2 v[1];
    v[0] = make_float2(0.f, 0.f);
#pragma unroll
    for (int s = 0; s < FC_CL; ++s) {          // fixed order: deterministic
      const float2 p = fc_lds2(BUF_A + (uint32_t)((s * FC_TM + x.row) * FC_SW + 2 * x.cp) * 4u);
      v[0].x += p.x;
      v[0].y += p.y;
    }
    const float2 b = fc_vec2(x, BV_B2, 2 * x.cp);
    v[0].x += b.x + o1.x;
    v[0].y += b.y + o1.y;
    fc_stats_send<1>(x, XB_ST3, v);
    FC_TS(15);
    float mean, rstd;
    fc_stats_merge(x, 0, mean, rstd);
    const float2 obj = fc_ln_apply(x, v[0], mean, rstd, BV_FN_G, BV_FN_B);             // kernel_update_head.py:214-215
    if (P.obj_out != nullptr && live) *reinterpret_cast<float2 *>(P.obj_out + grow * FC_K + gcol) = obj;
    fc_bcast_planes(x, 0, BUF_B, XB_OBJ, obj.x, obj.y);
    fc_xwait(x, XB_OBJ);
  }
  FC_TS(16);

  // ---- phase 4: cls_fcs[0] / mask_fcs[0] (no bias) + LayerNorm + ReLU
  fc_gemm(x, c, 17, 2, [&](int) { return BUF_B; },
          [&](int kc, int j, const float (&d)[4]) { fc_frag_to_slice(kc == 17 ? S0 : S2, j, x.lane, d); });
  __syncthreads();
  FC_TS(17);
  fc_ring_fill(x, c, 19);
  {
    float2 v[2];
    v[0] = with_cls ? fc_slice2(x, S0) : make_float2(0.f, 0.f);
    v[1] = fc_slice2(x, S2);
    fc_stats_send<2>(x, XB_ST4, v);
    FC_TS(18);
    float mean, rstd;
    fc_stats_merge(x, 0, mean, rstd);
    float2 cf = fc_ln_apply(x, v[0], mean, rstd, BV_CLN_G, BV_CLN_B);
    fc_stats_merge(x, 1, mean, rstd);
    float2 mf = fc_ln_apply(x, v[1], mean, rstd, BV_MLN_G, BV_MLN_B);
    if (with_cls) fc_bcast_planes(x, 1, BUF_A, XB_CLS, fmaxf(cf.x, 0.f), fmaxf(cf.y, 0.f));
    fc_bcast_planes(x, 2, BUF_H, XB_MASK, fmaxf(mf.x, 0.f), fmaxf(mf.y, 0.f));
    if (with_cls) fc_xwait(x, XB_CLS);
    fc_xwait(x, XB_MASK);
  }
  FC_TS(19);

  // ---- phase 5: fc_mask (A = mask branch) -> mask kernels; fc_cls (A = cls branch) -> global
  // fc_mask first: the hand-off of the mask kernels travels while fc_cls is multiplied
  fc_gemm(x, c, 19, 1, [&](int) { return BUF_H; }, [&](int, int j, const float (&d)[4]) { fc_frag_to_slice(S0, j, x.lane, d); });
  __syncthreads();
  {
    const float2 m = fc_slice2(x, S0), b = fc_vec2(x, BV_FCM_B, 2 * x.cp);
    fc_bcast_planes(x, 0, BUF_B, XB_MK, m.x + b.x, m.y + b.y);
    fc_cluster_arrive();            // last remote store of this CTA issued
  }
  FC_TS(20);
  fc_gemm(x, c, 20, 1, [&](int) { return BUF_A; }, [&](int, int j, const float (&d)[4]) {
    const int g = x.lane >> 2, cc = j * 8 + 2 * (x.lane & 3), col = FC_SW * (int)x.rank + cc;
    const float2 b = fc_vec2(x, BV_FCC_B, cc);
    float *dst = P.cls_out + (size_t)x.row0 * P.ncls + col;
    if (g < x.nvalid) {
      if (col < P.ncls) dst[(size_t)g * P.ncls] = d[0] + b.x;
      if (col + 1 < P.ncls) dst[(size_t)g * P.ncls + 1] = d[1] + b.y;
    }
    if (g + 8 < x.nvalid) {
      if (col < P.ncls) dst[(size_t)(g + 8) * P.ncls] = d[2] + b.x;
      if (col + 1 < P.ncls) dst[(size_t)(g + 8) * P.ncls + 1] = d[3] + b.y;
    }
  });
  __syncthreads();
  fc_ring_fill(x, c, 21);
  fc_xwait(x, XB_MK);
  FC_TS(21);

  // ---- phase 6: fold  a = mk . ft_w (+ the bias column mk . ft_b on CTA 0) -> a_ext rows and the mask conv's bf16 planes
  {
    const int fb = (int)blockIdx.z, n0 = (int)blockIdx.y * FC_TM;
    const size_t plane = (size_t)c.B * P.Npad * FC_K;
    fc_gemm(x, c, 21, 2, [&](int) { return BUF_B; }, [&](int kc, int j, const float (&d)[4]) {
      const int g = x.lane >> 2, t4 = x.lane & 3;
      if (kc == 22) {                    // one live weight row: column C of a_ext
        if (t4 == 0) {
          if (g < x.nvalid) P.a_ext[(size_t)(x.row0 + g) * P.lda + FC_K] = d[0];
          if (g + 8 < x.nvalid) P.a_ext[(size_t)(x.row0 + g + 8) * P.lda + FC_K] = d[2];
        }
        return;
      }
      const int col = FC_SW * (int)x.rank + j * 8 + 2 * t4;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = g + 8 * h;
        if (r >= x.nvalid) continue;
        const float v0 = d[2 * h], v1 = d[2 * h + 1];
        *reinterpret_cast<float2 *>(P.a_ext + (size_t)(x.row0 + r) * P.lda + col) = make_float2(v0, v1);
        uint32_t w3[3];
        split3_pair(v0, v1, w3[0], w3[1], w3[2]);
        const size_t o = ((size_t)fb * P.Npad + n0 + r) * FC_K + col;
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) *reinterpret_cast<uint32_t *>(P.a_split + pl * plane + o) = w3[pl];
      }
    });
  }
  FC_TS(30);
  fc_cluster_wait();                // no CTA of the cluster leaves while a peer may still address its shared memory
}

// ---- host ----------------------------------------------------------------------------------------------------------------
static FcChunk fc_chunk_std(const void *w, int ld, int row_off) {       // rows [row_off + 32 r, +32), K = columns [0, 256)
  FcChunk c;
  c.base = (const __nv_bfloat16 *)w + (size_t)row_off * ld;
  c.rank_stride = (long long)32 * ld;
  c.ld = ld;
  c.rows_total = 256;
  c.rows_per_rank = 32;
  c.pad_ = 0;
  return c;
}
static FcVec fc_vec_std(const float *p, int off) {
  FcVec v;
  v.p = p ? p + off : nullptr;
  v.rank_stride = 32;
  v.total = 256;
  return v;
}

template <typename PT>
static int fc_launch(void (*kernel)(PT), const char *name, const PT &p, size_t smem, cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(FC_CL, ceil_div(p.c.N, FC_TM), p.c.B);
  cfg.blockDim = dim3(FC_NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = FC_CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  VKN_LAUNCH_MARK(name, stream);
  VKN_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, p));
  return VKN_OK;
}

constexpr int FC_NCHUNK_A = 11, FC_NCHUNK_B = 23;
constexpr size_t FC_PACK_A = (size_t)FC_CL * FC_NCHUNK_A * FC_CHUNK_B, FC_PACK_B = (size_t)FC_CL * FC_NCHUNK_B * FC_CHUNK_B;

// chunk / vector tables of the two kernels (shared by the launch and by the weight re-layout)
static void fc_tables_a(const VknShape &s, const VknHeadW &w, bool have_xf, FcParamsA &a) {
  const int C = FC_K;
  memset(&a, 0, sizeof(a));
  FcCommon &c = a.c;
  c.N = s.N;
  c.B = s.B;
  c.nslot = FC_NSLOT_A;
  int k = 0;
  c.chunk[k++] = fc_chunk_std(w.ft_w, C, 0);
  if (have_xf) c.chunk[0].rows_total = 0;          // the pooled feature is handed in: no feat_transform GEMM
  c.chunk[k++] = fc_chunk_std(w.upd.inp_w, C, 0);
  c.chunk[k++] = fc_chunk_std(w.upd.inp_w, C, C);
  c.chunk[k++] = fc_chunk_std(w.upd.dyn_w, C, 0);
  c.chunk[k++] = fc_chunk_std(w.upd.dyn_w, C, C);
  c.chunk[k++] = fc_chunk_std(w.upd.ig_w, C, 0);
  c.chunk[k++] = fc_chunk_std(w.upd.ug_w, C, 0);
  c.chunk[k++] = fc_chunk_std(w.upd.fc_w, C, 0);
  c.chunk[k++] = fc_chunk_std(w.attn.in_w, C, 0);
  c.chunk[k++] = fc_chunk_std(w.attn.in_w, C, C);
  c.chunk[k++] = fc_chunk_std(w.attn.in_w, C, 2 * C);
  c.nchunks = k;
  c.vec[AV_FT_B] = fc_vec_std(w.ft_b, 0);
  c.vec[AV_INP_B0] = fc_vec_std(w.upd.inp_b, 0);
  c.vec[AV_INP_B1] = fc_vec_std(w.upd.inp_b, C);
  c.vec[AV_DYN_B0] = fc_vec_std(w.upd.dyn_b, 0);
  c.vec[AV_DYN_B1] = fc_vec_std(w.upd.dyn_b, C);
  c.vec[AV_IG_B] = fc_vec_std(w.upd.ig_b, 0);
  c.vec[AV_UG_B] = fc_vec_std(w.upd.ug_b, 0);
  c.vec[AV_NIN_G] = fc_vec_std(w.upd.norm_in_g, 0);
  c.vec[AV_NIN_B] = fc_vec_std(w.upd.norm_in_b, 0);
  c.vec[AV_NOUT_G] = fc_vec_std(w.upd.norm_out_g, 0);
  c.vec[AV_NOUT_B] = fc_vec_std(w.upd.norm_out_b, 0);
  c.vec[AV_ININ_G] = fc_vec_std(w.upd.inorm_in_g, 0);
  c.vec[AV_ININ_B] = fc_vec_std(w.upd.inorm_in_b, 0);
  c.vec[AV_INOUT_G] = fc_vec_std(w.upd.inorm_out_g, 0);
  c.vec[AV_INOUT_B] = fc_vec_std(w.upd.inorm_out_b, 0);
  c.vec[AV_FC_B] = fc_vec_std(w.upd.fc_b, 0);
  c.vec[AV_FCN_G] = fc_vec_std(w.upd.fc_norm_g, 0);
  c.vec[AV_FCN_B] = fc_vec_std(w.upd.fc_norm_b, 0);
  c.vec[AV_Q_B] = fc_vec_std(w.attn.in_b, 0);
  c.vec[AV_K_B] = fc_vec_std(w.attn.in_b, C);
  c.vec[AV_V_B] = fc_vec_std(w.attn.in_b, 2 * C);
  c.nvec = AV_COUNT;
}
static void fc_tables_b(const VknShape &s, const VknHeadW &w, bool with_cls, FcParamsB &b) {
  const int C = FC_K, F = s.ffn_dim;
  memset(&b, 0, sizeof(b));
  FcCommon &c = b.c;
  c.N = s.N;
  c.B = s.B;
  c.nslot = FC_NSLOT_B;
  int k = 0;
  c.chunk[k++] = fc_chunk_std(w.attn.out_w, C, 0);
  for (int i = 0; i < 8; ++i) {           // hidden columns 256 r + 32 i .. + 32
    FcChunk ch = fc_chunk_std(w.ffn.w1, C, 32 * i);
    ch.rank_stride = (long long)FC_K * C;
    c.chunk[k++] = ch;
  }
  for (int i = 0; i < 8; ++i) {           // output rows 32 i .. + 32, K slice [256 r, 256 r + 256)
    FcChunk ch = fc_chunk_std(w.ffn.w2, F, 32 * i);
    ch.rank_stride = FC_K;
    c.chunk[k++] = ch;
  }
  {
    FcChunk ch = fc_chunk_std(with_cls ? w.cls_fc_w[0] : w.mask_fc_w[0], C, 0);
    if (!with_cls) ch.rows_total = 0;
    c.chunk[k++] = ch;
  }
  c.chunk[k++] = fc_chunk_std(w.mask_fc_w[0], C, 0);
  c.chunk[k++] = fc_chunk_std(w.fc_mask_w, C, 0);
  {
    FcChunk ch = fc_chunk_std(with_cls ? w.fc_cls_w : w.fc_mask_w, C, 0);
    ch.rows_total = with_cls ? s.num_classes : 0;
    c.chunk[k++] = ch;
  }
  c.chunk[k++] = fc_chunk_std(w.ft_wt_ext, C, 0);
  {
    FcChunk ch = fc_chunk_std(w.ft_wt_ext, C, C);       // row C = ft_b: CTA 0 only
    ch.rank_stride = 0;
    ch.rows_total = 1;
    c.chunk[k++] = ch;
  }
  c.nchunks = k;
  c.vec[BV_OUT_B] = fc_vec_std(w.attn.out_b, 0);
  c.vec[BV_AN_G] = fc_vec_std(w.attn.norm_g, 0);
  c.vec[BV_AN_B] = fc_vec_std(w.attn.norm_b, 0);
  for (int i = 0; i < 8; ++i) {
    FcVec v;
    v.p = w.ffn.b1 + 32 * i;
    v.rank_stride = FC_K;
    v.total = F;
    c.vec[BV_B1 + i] = v;
  }
  c.vec[BV_B2] = fc_vec_std(w.ffn.b2, 0);
  c.vec[BV_FN_G] = fc_vec_std(w.ffn.norm_g, 0);
  c.vec[BV_FN_B] = fc_vec_std(w.ffn.norm_b, 0);
  c.vec[BV_CLN_G] = fc_vec_std(with_cls ? w.cls_ln_g[0] : nullptr, 0);
  c.vec[BV_CLN_B] = fc_vec_std(with_cls ? w.cls_ln_b[0] : nullptr, 0);
  c.vec[BV_MLN_G] = fc_vec_std(w.mask_ln_g[0], 0);
  c.vec[BV_MLN_B] = fc_vec_std(w.mask_ln_b[0], 0);
  c.vec[BV_FCM_B] = fc_vec_std(w.fc_mask_b, 0);
  {
    FcVec v = fc_vec_std(with_cls ? w.fc_cls_b : nullptr, 0);
    v.total = s.num_classes;
    c.vec[BV_FCC_B] = v;
  }
  c.nvec = BV_COUNT;
}

// Weight re-layout for VknHeadW.fc_pack: every ring chunk of every cluster rank as one contiguous, zero-padded image
// [32][FC_LD] (kernel A's chunks, then kernel B's), so that the ring is fed by ONE bulk copy per chunk instead of one per row
// (a bulk copy costs ~35 ns of issue whatever its size: measured, profiles/r2_frame_chain.md)
__global__ void __launch_bounds__(256) vkn_frame_chain_pack_kernel(const __grid_constant__ FcCommon c, __nv_bfloat16 *out) {
  const int k = blockIdx.x, rank = blockIdx.y;
  const FcChunk &ch = c.chunk[k];
  int rows = ch.rows_total - rank * ch.rows_per_rank;
  rows = max(0, min(rows, min(ch.rows_per_rank, 32)));
  const __nv_bfloat16 *src = ch.base + (long long)rank * ch.rank_stride;
  uint4 *dst = reinterpret_cast<uint4 *>(out + ((size_t)rank * c.nchunks + k) * (FC_CHUNK_B / 2));
  for (int idx = threadIdx.x; idx < 32 * (FC_LD / 8); idx += 256) {
    const int r = idx / (FC_LD / 8), pc = idx - r * (FC_LD / 8);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows && pc < FC_K / 8) v = *reinterpret_cast<const uint4 *>(src + (size_t)r * ch.ld + pc * 8);
    dst[idx] = v;
  }
}

static bool fc_weights_supported(const VknShape &s, const VknHeadW &w) {
  if (s.C != FC_K || s.num_heads != FC_CL || s.w_dtype != VKN_BF16 || !s.with_ffn || s.ffn_dim != FC_CL * FC_K) return false;
  if ((w.fc_cls_w != nullptr && w.num_cls_fcs != 1) || w.num_mask_fcs != 1) return false;
  if (s.num_classes > 256) return false;
  return true;
}

bool frame_chain_supported(const VknShape &s, const VknHeadW &w) {
  if (const char *e = getenv("VKN_FRAME_CHAIN"))
    if (e[0] == '0') return false;
  if (!fc_weights_supported(s, w)) return false;      // frames_per_set > 1 (clip head) only changes pooling / mask conv: rows are kernel sets
  // K / V of one head must fit the two A buffers they overlay, a lane holds at most 6 keys' probabilities
  if ((size_t)(2 * s.N * 33) * 4 > (size_t)2 * FC_ABUF || s.N > 192) return false;
  return true;
}

size_t frame_chain_pack_bytes(const VknShape &s, const VknHeadW &w) { return fc_weights_supported(s, w) ? FC_PACK_A + FC_PACK_B : 0; }

int launch_frame_chain_pack(const VknShape &s, const VknHeadW &w, void *out, size_t bytes, cudaStream_t stream) {
  if (!fc_weights_supported(s, w)) VKN_FAIL(VKN_E_UNSUPPORTED, "the single-frame row engine does not apply to this head");
  if (!out || bytes < FC_PACK_A + FC_PACK_B) VKN_FAIL(VKN_E_WORKSPACE, "fc_pack buffer: %zu bytes given, %zu needed", bytes, FC_PACK_A + FC_PACK_B);
  if (reinterpret_cast<uintptr_t>(out) & 15) VKN_FAIL(VKN_E_INVALID, "fc_pack buffer must be 16-byte aligned");
  FcParamsA a;
  fc_tables_a(s, w, false, a);
  FcParamsB b;
  fc_tables_b(s, w, w.fc_cls_w != nullptr, b);
  if (a.c.nchunks != FC_NCHUNK_A || b.c.nchunks != FC_NCHUNK_B) VKN_FAIL(VKN_E_INVALID, "frame chain chunk tables out of sync");
  VKN_LAUNCH_MARK("vkn_frame_chain_pack_kernel", stream);
  vkn_frame_chain_pack_kernel<<<dim3(FC_NCHUNK_A, FC_CL), 256, 0, stream>>>(a.c, (__nv_bfloat16 *)out);
  vkn_frame_chain_pack_kernel<<<dim3(FC_NCHUNK_B, FC_CL), 256, 0, stream>>>(b.c, (__nv_bfloat16 *)((char *)out + FC_PACK_A));
  VKN_CUDA_OK(cudaGetLastError());
  return VKN_OK;
}

// The row operators of one stage for P = B * N rows (frames_per_set == 1): pooled sums / counts (or a ready pooled feature) +
// proposal_feat in, obj_feat / cls_score / the mask conv's operands out.  qkv_ws [P][3C] and obj0_ws [P][C] are scratch.
int launch_frame_chain(const VknShape &s, const VknHeadW &w, const float *xp0, const float *cnt, const float *x_feat_in,
                       const float *pf, float *x_feat_out, float *obj0_ws, float *qkv_ws, float *obj_out, float *cls_out, float *a_ext,
                       int lda, void *a_split, int Npad, cudaStream_t stream) {
  static unsigned long long attr_mask = 0;
  const size_t smem_a = fc_smem_bytes(FC_NSLOT_A, FC_NSL_A), smem_b = fc_smem_bytes(FC_NSLOT_B, FC_NSL_B);
  if (first_use_on_device(attr_mask)) {
    VKN_CUDA_OK(cudaFuncSetAttribute(vkn_frame_chain_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    VKN_CUDA_OK(cudaFuncSetAttribute(vkn_frame_chain_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
  }
  const bool with_cls = w.fc_cls_w != nullptr && cls_out != nullptr;
  const bool pack_ok = w.fc_pack != nullptr;
  int handoff_stores = 0;
  if (const char *e = getenv("VKN_FC_BULK")) handoff_stores = e[0] == '0';
  {
    FcParamsA a;
    fc_tables_a(s, w, x_feat_in != nullptr, a);
    a.c.pack = pack_ok ? (const __nv_bfloat16 *)w.fc_pack : nullptr;
    a.c.dbg = debug_ts_slot();
    a.c.handoff_stores = handoff_stores;
    a.xp0 = xp0;
    a.cnt = cnt;
    a.pf = pf;
    a.x_feat_in = x_feat_in;
    a.x_feat_out = x_feat_out;
    a.obj0 = obj0_ws;
    a.qkv = qkv_ws;
    VKN_TRY(fc_launch(vkn_frame_chain_a_kernel, "vkn_frame_chain_a_kernel", a, smem_a, stream));
  }
  {
    FcParamsB b;
    fc_tables_b(s, w, with_cls, b);
    b.c.pack = pack_ok ? (const __nv_bfloat16 *)((const char *)w.fc_pack + FC_PACK_A) : nullptr;
    b.c.dbg = debug_ts_slot();
    b.c.handoff_stores = handoff_stores;
    b.qkv = qkv_ws;
    b.obj0 = obj0_ws;
    b.obj_out = obj_out;
    b.cls_out = with_cls ? cls_out : nullptr;
    b.a_ext = a_ext;
    b.a_split = (__nv_bfloat16 *)a_split;
    b.ncls = s.num_classes;
    b.lda = lda;
    b.Npad = Npad;
    b.scale = 1.0f / sqrtf((float)(FC_K / s.num_heads));
    VKN_TRY(fc_launch(vkn_frame_chain_b_kernel, "vkn_frame_chain_b_kernel", b, smem_b, stream));
  }
  return VKN_OK;
}

}  // namespace vkn
