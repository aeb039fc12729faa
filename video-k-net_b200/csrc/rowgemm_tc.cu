// tcgen05 engine for the row operators of a stage when many rows are in flight (P = frames x kernels >= a few
// hundred): every nn.Linear of KernelUpdator / MultiheadAttention / FFN / the FC heads
// (knet/kernel_updator.py:56-94, knet/det/kernel_update_head.py:204-227) is   Y = epi(X . W^T)   with
//
//   A = X as three bf16 planes hi/mid/lo [3][M][K] (X = hi + mid + lo to 24 bits, written once per row by the
//       producing kernel), K-major, TMA (SWIZZLE_128B), one 3-D box {64 k, 128 rows, 3 planes} per pipeline stage
//   B = W [N][K] bf16 as stored by the module, K-major, TMA (SWIZZLE_128B), box {64 k, BN}
//   D = [128 rows x BN] fp32 in TMEM; the three planes accumulate into the same tile (every bf16 x bf16 product is
//       exact, so the result carries fp32-level accuracy: DESIGN.md section 4)
//
// Persistent: min(#tiles, 148) CTAs walk tiles of 128 rows x BN columns (BN in {32..256} by a cost model in the launcher);
// the TMA ring runs across tiles and the accumulator is double-buffered in TMEM.  Epilogue: thread = row (TMEM lane); bias
// (x rowscale), residual, ReLU, optional fused LayerNorm; fp32 rows and / or bf16 planes for the next GEMM, 32 columns per
// thread in registers, leaving through swizzled per-warp staging boxes and TMA stores (direct 256-bit stores when a tensor
// map cannot describe the output).  Split-K and up to two problems per launch are folded into the tile list.  Weight tiles
// of the first ring fill are issued BEFORE griddepcontrol.wait (PDL).
#include "tc.cuh"
#include "rowops.cuh"

namespace vkn {

struct RgProb {
  CUtensorMap tmA;
  CUtensorMap tmW;
  CUtensorMap tmOut;          // fp32 rows {N, M, K slices}, box {32, 32, 1}, SWIZZLE_128B (valid when tma_out)
  CUtensorMap tmPl;           // bf16 planes {split_C, M, 3}, box {32, 32, 3}, SWIZZLE_64B (valid when tma_pl)
  int tma_out, tma_pl;
  const float *bias, *rowscale, *res, *ln_g, *ln_b, *mul, *add2;
  float *out;
  __nv_bfloat16 *planes;
  long long out_split_stride, plane_elems;
  int ldres, ldmul, ldadd2, ldo, M, N, K, epi, split_N, split_Npad, split_C;
  int mt, nt, tiles;          // row tiles, column tiles, tiles of this problem (mt * nt * ksplit)
};
struct RgBatch {
  RgProb p[2];
  int nprob, ksplit, stages, BN, total_tiles;
  int stg_bytes, pl_off;       // per-warp TMA-store staging: 10 KB (fp32 box + plane boxes side by side, pl_off = 4096) or 6 KB (shared, pl_off = 0)
  uint32_t idesc;
  unsigned long long *dbg;     // optional per-CTA phase timestamps (vkn_debug_timestamps), null in production
};

__device__ __forceinline__ unsigned long long rg_time() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define RG_TS(slot)                                                       \
  do {                                                                    \
    if (batch.dbg != nullptr) batch.dbg[(size_t)blockIdx.x * 16 + (slot)] = rg_time(); \
  } while (0)

constexpr uint32_t RG_A_PLANE = 128u * 128u;        // 128 rows x 64 k x 2 B
constexpr uint32_t RG_A_BYTES = 3u * RG_A_PLANE;

constexpr int RG_THREADS = 320;      // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue

// tile t of the launch -> problem, K slice, row / column origin
struct RgTile {
  int prob, ks, row0, col0;
};
__device__ __forceinline__ RgTile rg_decode(const RgBatch &batch, int t) {
  RgTile r;
  r.prob = 0;
  if (t >= batch.p[0].tiles) {
    t -= batch.p[0].tiles;
    r.prob = 1;
  }
  const RgProb &P = batch.p[r.prob];
  const int per = P.mt * P.nt;
  r.ks = t / per;
  const int rem = t - r.ks * per;
  const int m = rem / P.nt;
  r.row0 = m * 128;
  r.col0 = (rem - m * P.nt) * batch.BN;
  return r;
}


// ---- epilogue pieces (thread = row; 32 consecutive columns in registers) ---------------------------------------
#ifdef VKN_EPI_PROF
__device__ unsigned long long g_epi_prof[16];
#define EPI_T0 const long long ep0__ = clock64()
#define EPI_T(slot)                                                              \
  do {                                                                           \
    if (threadIdx.x == 64) atomicAdd(&g_epi_prof[slot], (unsigned long long)(clock64() - ep0__)); \
  } while (0)
#else
#define EPI_T0
#define EPI_T(slot)
#endif
struct RgRowCtx {
  int row, epi;
  int row_base, ks;           // first row of this warp's 32-row block, K slice (TMA store coordinates)
  uint32_t stg;               // this warp's staging buffer (shared-space address, 1024-byte aligned)
  uint32_t pl_off;            // offset of the plane boxes inside it (0: they reuse the fp32 box after its store was read)
  bool live, out_vec, res_vec, pl_vec, bias_vec;
  float rs;
  size_t prow;
  float *outp;
};

// v[0..32) = accumulator (+ rowscale x bias) (+ residual) of columns [col, col + 32).  The global loads (bias, residual) are
// issued BEFORE the TMEM load so that their L2 latency overlaps it (and, for the first chunk of a tile, the wait for the
// accumulator wait of the next chunk).
__device__ __forceinline__ void rg_chunk_load(const RgProb &P, const RgRowCtx &R, uint32_t taddr, int col, int nc, float (&v)[32]) {
  float add[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) add[e] = 0.f;
  if (R.live) {
    if (R.epi & EPI_BIAS) {
      const float *bp = P.bias + col;
      if (R.bias_vec && nc == 32 && (col & 3) == 0) {
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bp + e));
          add[e] = R.rs * b4.x; add[e + 1] = R.rs * b4.y; add[e + 2] = R.rs * b4.z; add[e + 3] = R.rs * b4.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (e < nc) add[e] = R.rs * __ldg(bp + e);
      }
    }
    if (R.epi & EPI_RES) {
      const float *rp = P.res + (size_t)R.row * P.ldres + col;
      if (R.res_vec && nc == 32) {
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 t4 = __ldcg(reinterpret_cast<const float4 *>(rp + e));
          add[e] += t4.x; add[e + 1] += t4.y; add[e + 2] += t4.z; add[e + 3] += t4.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (e < nc) add[e] += __ldcg(rp + e);
      }
    }
  }
  uint32_t r[32];
  {
    EPI_T0;
    tmem_ld32(taddr, r);
    EPI_T(0);
  }
  {
    EPI_T0;
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]) + add[e];
    EPI_T(1);
  }
}

// sigmoid / elementwise factor / addend after bias, residual and LayerNorm (the KernelUpdator gate, kernel_updator.py:70-88)
__device__ __forceinline__ void rg_post(const RgProb &P, const RgRowCtx &R, int col, int nc, float (&v)[32]) {
  if (!(R.epi & (EPI_SIGMOID | EPI_MUL | EPI_ADD2)) || !R.live) return;
  float m[32];
  if (R.epi & EPI_MUL) {                               // in flight while the sigmoids are evaluated
    const float *mp = P.mul + (size_t)R.row * P.ldmul + col;
#pragma unroll
    for (int e = 0; e < 32; e += 4)
      if (e < nc) {
        const float4 t4 = __ldcg(reinterpret_cast<const float4 *>(mp + e));
        m[e] = t4.x; m[e + 1] = t4.y; m[e + 2] = t4.z; m[e + 3] = t4.w;
      }
  }
  if (R.epi & EPI_SIGMOID) {
    EPI_T0;
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = sigmoid_fast(v[e]);
    EPI_T(8);
  }
  if (R.epi & EPI_MUL) {
    EPI_T0;
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (e < nc) v[e] *= m[e];
    EPI_T(9);
  }
  if (R.epi & EPI_ADD2) {
    EPI_T0;
    const float *ap = P.add2 + (size_t)R.row * P.ldadd2 + col;
#pragma unroll
    for (int e = 0; e < 32; e += 4)
      if (e < nc) {
        const float4 t4 = __ldcg(reinterpret_cast<const float4 *>(ap + e));
        v[e] += t4.x; v[e + 1] += t4.y; v[e + 2] += t4.z; v[e + 3] += t4.w;
      }
    EPI_T(10);
  }
}

// (ReLU) -> fp32 rows and / or bf16 hi/mid/lo planes
__device__ __forceinline__ void rg_chunk_store(const RgProb &P, const RgRowCtx &R, int col, int nc, float (&v)[32]) {
  if (R.epi & EPI_RELU) {
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
  }
  const int lane = threadIdx.x & 31;
  if (!(R.epi & EPI_NOOUT) && P.tma_out) {
    // fp32 rows through a 128B-swizzled [32 rows][128 B] staging box and ONE TMA store: full-line writes instead of 32
    // rows x 32 B per store instruction (the per-SM store rate of the direct path bounded this kernel)
    {
      EPI_T0;
      if (lane == 0) bulk_wait_group_read<0>();
      __syncwarp();
      EPI_T(2);
    }
    EPI_T0;
    const uint32_t rb = R.stg + (uint32_t)lane * 128u;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts_v4(rb + (uint32_t)((j ^ (lane & 7)) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
             __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(&P.tmOut, R.stg, col, R.row_base, R.ks);
      bulk_commit_group();
    }
    EPI_T(11);
  } else if (!(R.epi & EPI_NOOUT) && R.live) {
    float *op = R.outp + (size_t)R.row * P.ldo + col;
    if (R.out_vec && nc == 32) {
#pragma unroll
      for (int e = 0; e < 32; e += 8)
        stg_v8(op + e, __float_as_uint(v[e]), __float_as_uint(v[e + 1]), __float_as_uint(v[e + 2]), __float_as_uint(v[e + 3]),
               __float_as_uint(v[e + 4]), __float_as_uint(v[e + 5]), __float_as_uint(v[e + 6]), __float_as_uint(v[e + 7]));
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (e < nc) op[e] = v[e];
    }
  }
  if ((R.epi & EPI_SPLIT3) && !(R.epi & EPI_SPLIT2H) && col < P.split_C && P.tma_pl) {
    uint32_t w[3][16];
    EPI_T0;
#pragma unroll
    for (int e = 0; e < 32; e += 2) split3_pair(v[e], v[e + 1], w[0][e >> 1], w[1][e >> 1], w[2][e >> 1]);
    EPI_T(12);
    // staging free?  With side-by-side boxes the wait at the top of the fp32 path already covered the previous chunk.
    if (R.pl_off == 0 || (R.epi & EPI_NOOUT) || !P.tma_out) {
      EPI_T0;
      if (lane == 0) bulk_wait_group_read<0>();
      __syncwarp();
      EPI_T(3);
    }
    const uint32_t rb = R.stg + R.pl_off + (uint32_t)lane * 64u;
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int pl = 0; pl < 3; ++pl)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        sts_v4(rb + (uint32_t)pl * 2048u + (uint32_t)((c ^ sw) << 4), w[pl][4 * c], w[pl][4 * c + 1], w[pl][4 * c + 2], w[pl][4 * c + 3]);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(&P.tmPl, R.stg + R.pl_off, col, R.row_base, 0);
      bulk_commit_group();
    }
    EPI_T(13);
  } else if ((R.epi & EPI_SPLIT3) && (R.epi & EPI_SPLIT2H) && col < P.split_C && R.live) {
    // two fp16 planes (the mask conv's fp16 mode): same buffer, same row layout, planes 0 and 1
    const int np = min(32, P.split_C - col);
    __nv_bfloat16 *pp = P.planes + R.prow * P.split_C + col;
    uint32_t w2[2][16];
#pragma unroll
    for (int e = 0; e < 32; e += 2) split2h_pair(v[e], v[e + 1], w2[0][e >> 1], w2[1][e >> 1]);
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
      const uint32_t (&w)[16] = w2[pl];
      __nv_bfloat16 *pt = pp + (size_t)pl * P.plane_elems;
      if (R.pl_vec && np == 32) {
#pragma unroll
        for (int e = 0; e < 16; e += 8) stg_v8(pt + 2 * e, w[e], w[e + 1], w[e + 2], w[e + 3], w[e + 4], w[e + 5], w[e + 6], w[e + 7]);
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (e < np) pt[e] = __ushort_as_bfloat16((uint16_t)(w[e >> 1] >> ((e & 1) * 16)));
      }
    }
  } else if ((R.epi & EPI_SPLIT3) && col < P.split_C && R.live) {
    const int np = min(32, P.split_C - col);
    __nv_bfloat16 *pp = P.planes + R.prow * P.split_C + col;
    uint32_t w3[3][16];
#pragma unroll
    for (int e = 0; e < 32; e += 2) split3_pair(v[e], v[e + 1], w3[0][e >> 1], w3[1][e >> 1], w3[2][e >> 1]);
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      const uint32_t (&w)[16] = w3[pl];
      __nv_bfloat16 *pt = pp + (size_t)pl * P.plane_elems;
      if (R.pl_vec && np == 32) {
#pragma unroll
        for (int e = 0; e < 16; e += 8) stg_v8(pt + 2 * e, w[e], w[e + 1], w[e + 2], w[e + 3], w[e + 4], w[e + 5], w[e + 6], w[e + 7]);
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (e < np) pt[e] = __ushort_as_bfloat16((uint16_t)(w[e >> 1] >> ((e & 1) * 16)));
      }
    }
  }
}

// Epilogue of ONE accumulator tile (rows [row0, row0+128) x columns [col0, col0+BN) of problem P, K slice ks), executed by
// the 8 epilogue warps: thread = row (TMEM lane), the two warps of a lane quarter take alternate 32-column blocks.
// `tacc` = TMEM address of the accumulator buffer (column base), `acc_empty` = barrier that hands it back to the MMA warp
// (one arrive per warp, as soon as its share is in registers), `stg` = this warp's TMA-store staging, `st_base` = fused-
// LayerNorm scratch of this accumulator buffer ([128 rows][8] floats).
__device__ __forceinline__ void rg_epilogue_tile(const RgProb &P, int row0, int col0, int ks, int BN, uint32_t tacc_base,
                                                 uint32_t acc_empty, uint32_t stg, uint32_t pl_off, float *st_base, int warp,
                                                 int lane, int nthreads_epi) {
  const int q = warp & 3;                                 // TMEM lane quarter this warp may read
  const int half = (warp - 2) >> 2;                       // 0: even column blocks, 1: odd
  RgRowCtx R;
  R.epi = P.epi;
  R.row = row0 + q * 32 + lane;
  R.live = R.row < P.M;
  R.row_base = row0 + q * 32;
  R.ks = ks;
  R.stg = stg;
  R.pl_off = pl_off;
  R.outp = P.out + (size_t)ks * P.out_split_stride;
  R.out_vec = (P.ldo % 8 == 0) && ((reinterpret_cast<uintptr_t>(R.outp) & 31) == 0);
  R.res_vec = (P.ldres % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.res) & 15) == 0);
  R.pl_vec = (P.split_C % 16 == 0) && ((reinterpret_cast<uintptr_t>(P.planes) & 31) == 0) && (P.plane_elems % 16 == 0);
  R.bias_vec = (reinterpret_cast<uintptr_t>(P.bias) & 15) == 0;
  R.rs = (R.live && (R.epi & EPI_ROWSCALE)) ? __ldcg(P.rowscale + R.row) : 1.f;
  R.prow = (size_t)R.row;                               // row of the plane buffer this thread writes
  if ((R.epi & EPI_SPLIT3) && P.split_N != P.split_Npad) {
    const int b = R.row / P.split_N;
    R.prow = (size_t)b * P.split_Npad + (R.row - b * P.split_N);
  }
  if (R.epi & (EPI_MUL | EPI_ADD2)) {
    // the factor / addend rows may come from an earlier tile of the SAME step: this very warp stored them (same rows, same
    // column blocks), possibly through TMA -- drain its bulk stores before reading them back
    if (lane == 0) bulk_wait_all();
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
  }
  const uint32_t tacc = tacc_base + ((uint32_t)(q * 32) << 16);
  int last_c0 = half * 32;                              // last column block this warp reads from TMEM
  while (last_c0 + 64 < BN && col0 + last_c0 + 64 < P.N) last_c0 += 64;
  const bool any = half * 32 < BN && col0 + half * 32 < P.N;   // this warp owns at least one column block
  if (R.epi & EPI_LN) {
    // ---- fused LayerNorm: the tile spans the row (host: nt == 1); the two warps of a lane quarter own alternate
    //      32-column blocks, so each thread reduces its blocks (shifted sums: no cancellation), the pair merges
    //      (count, mean, M2) through shared memory, then a second pass over TMEM normalises and stores.
    float K0 = 0.f, S1 = 0.f, S2 = 0.f, cntf = 0.f;
    bool first = true;
    for (int c0 = half * 32; c0 < BN; c0 += 64) {
      const int col = col0 + c0;
      if (col >= P.N) break;
      const int nc = min(32, P.N - col);
      float v[32];
      rg_chunk_load(P, R, tacc + (uint32_t)c0, col, nc, v);
      {   // keep acc + bias + residual in the accumulator itself: the second pass re-reads TMEM, not global memory
        uint32_t vr[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) vr[e] = __float_as_uint(v[e]);
        tmem_st32(tacc + (uint32_t)c0, vr);
      }
      if (first) {
        K0 = v[0];
        first = false;
      }
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (e < nc) {
          const float d = v[e] - K0;
          S1 += d;
          S2 = fmaf(d, d, S2);
        }
      cntf += (float)nc;
    }
    float mean_h = 0.f, m2_h = 0.f;
    if (cntf > 0.f) {
      mean_h = K0 + S1 / cntf;
      m2_h = S2 - S1 * S1 / cntf;
    }
    float *st = st_base + (size_t)(q * 32 + lane) * 8;
    st[half * 4 + 0] = cntf;
    st[half * 4 + 1] = mean_h;
    st[half * 4 + 2] = m2_h;
    {
      EPI_T0;
      named_bar_sync(1, nthreads_epi);
      EPI_T(7);
    }
    const float cb = st[(half ^ 1) * 4 + 0], mb = st[(half ^ 1) * 4 + 1], m2b = st[(half ^ 1) * 4 + 2];
    // merge in a fixed order (half 0 first) so both threads of a row get identical statistics
    const float n0 = half ? cb : cntf, mu0 = half ? mb : mean_h, q0 = half ? m2b : m2_h;
    const float n1 = half ? cntf : cb, mu1 = half ? mean_h : mb, q1 = half ? m2_h : m2b;
    const float nn = n0 + n1, delta = mu1 - mu0;
    const float mean = mu0 + delta * (n1 / nn);
    const float var = (q0 + q1 + delta * delta * (n0 * n1 / nn)) / nn;
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    for (int c0 = half * 32; c0 < BN; c0 += 64) {
      const int col = col0 + c0;
      if (col >= P.N) break;
      const int nc = min(32, P.N - col);
      float v[32];
      {
        uint32_t vr[32];
        tmem_ld32(tacc + (uint32_t)c0, vr);
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(vr[e]);
      }
      if (c0 == last_c0) {                              // second read done: hand the accumulator back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
      }
      if (nc == 32 && ((reinterpret_cast<uintptr_t>(P.ln_g) | reinterpret_cast<uintptr_t>(P.ln_b)) & 15) == 0 && (col & 3) == 0) {
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 g4 = __ldg(reinterpret_cast<const float4 *>(P.ln_g + col + e));
          const float4 b4 = __ldg(reinterpret_cast<const float4 *>(P.ln_b + col + e));
          v[e] = fmaf((v[e] - mean) * rstd, g4.x, b4.x);
          v[e + 1] = fmaf((v[e + 1] - mean) * rstd, g4.y, b4.y);
          v[e + 2] = fmaf((v[e + 2] - mean) * rstd, g4.z, b4.z);
          v[e + 3] = fmaf((v[e + 3] - mean) * rstd, g4.w, b4.w);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (e < nc) v[e] = fmaf((v[e] - mean) * rstd, __ldg(P.ln_g + col + e), __ldg(P.ln_b + col + e));
      }
      {
        EPI_T0;
        rg_post(P, R, col, nc, v);
        EPI_T(4);
      }
      {
        EPI_T0;
        rg_chunk_store(P, R, col, nc, v);
        EPI_T(5);
      }
    }
  } else {
    for (int c0 = half * 32; c0 < BN; c0 += 64) {
      const int col = col0 + c0;
      if (col >= P.N) break;
      const int nc = min(32, P.N - col);
      float v[32];
      rg_chunk_load(P, R, tacc + (uint32_t)c0, col, nc, v);
      if (c0 == last_c0) {                              // this warp's share of the accumulator is in registers
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
      }
      {
        EPI_T0;
        rg_post(P, R, col, nc, v);
        EPI_T(4);
      }
      {
        EPI_T0;
        rg_chunk_store(P, R, col, nc, v);
        EPI_T(5);
      }
    }
  }
  if (!any) {                                           // a warp whose column blocks all lie past N
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(acc_empty);
  }
}

// Persistent: a CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; the TMA ring runs on across tiles and the
// accumulator is double-buffered in TMEM (2 x BN columns), so the tensor core works on tile i+1 while the eight epilogue
// warps drain tile i.  BN = 256 halves the shared-memory operand traffic per FLOP of the 3-plane product (the limiter
// of this kernel: every tcgen05.mma re-reads its A and B tiles from shared memory).
__global__ void __launch_bounds__(RG_THREADS, 1) vkn_rowgemm_tc_kernel(const __grid_constant__ RgBatch batch) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int BN = batch.BN, STG = batch.stages, total = batch.total_tiles;
  const uint32_t w_bytes = (uint32_t)BN * 128u;
  const uint32_t stage_bytes = RG_A_BYTES + w_bytes;
  const uint32_t stg0 = smem_u32(smem + (size_t)STG * stage_bytes);    // 8 epilogue warps x 6 KB TMA-store staging
  uint64_t *bars = (uint64_t *)(smem + (size_t)STG * stage_bytes + 8 * (size_t)batch.stg_bytes);
  const uint32_t bar0 = smem_u32(bars);
  // full[s] = bar0 + 8 s (TMA: A planes + W tile), empty[s] = bar0 + 8 (STG + s) (MMAs retired),
  // acc_full[a] = bar0 + 8 (2 STG + a), acc_empty[a] = bar0 + 8 (2 STG + 2 + a)
  const uint32_t acc_full0 = bar0 + 16 * STG, acc_empty0 = acc_full0 + 16;
  uint32_t *tmem_slot = (uint32_t *)(bars + 2 * STG + 4);
  float *ln_stat = (float *)(tmem_slot + 4);                 // [2][128 rows][2 halves][4]: fused-LayerNorm partials (8 KB)
  const uint32_t smem0 = smem_u32(smem);
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) RG_TS(0);
  const uint32_t ncols = BN < 16 ? 32u : 2u * (uint32_t)BN;     // 64 / 128 / 256 / 512

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&batch.p[0].tmA);
      prefetch_tmap(&batch.p[0].tmW);
      if (batch.nprob > 1) {
        prefetch_tmap(&batch.p[1].tmA);
        prefetch_tmap(&batch.p[1].tmW);
      }
      for (int s = 0; s < STG; ++s) {
        mbar_init(bar0 + 8 * s, 1);
        mbar_init(bar0 + 8 * (STG + s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(acc_full0 + 8 * a, 1);
        mbar_init(acc_empty0 + 8 * a, 8);                   // one arrive per epilogue warp
      }
      fence_barrier_init();
    }
  } else if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), ncols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) RG_TS(1);

  if (warp == 0) {
    if (lane == 0) {
      int s = 0, it = 0;
      uint32_t ph = 0;
      int npre = 0;
      {     // weight tiles of the first ring fill: nobody in the chain writes weights -> before the PDL wait
        const RgTile T = rg_decode(batch, blockIdx.x);
        const RgProb &P = batch.p[T.prob];
        const int nk = ((P.K + 63) / 64) / batch.ksplit;
        npre = nk < STG ? nk : STG;
        for (int i = 0; i < npre; ++i) {
          mbar_expect_tx(bar0 + 8 * i, stage_bytes);
          tma_load_2d(smem0 + i * stage_bytes + RG_A_BYTES, &P.tmW, bar0 + 8 * i, (T.ks * nk + i) * 64, T.col0);
        }
      }
      pdl_wait();                                           // the A planes come from the previous kernel
      RG_TS(2);
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const RgTile T = rg_decode(batch, t);
        const RgProb &P = batch.p[T.prob];
        const int nk = ((P.K + 63) / 64) / batch.ksplit;    // host guarantees divisibility
        const int kb = T.ks * nk;
        for (int i = 0; i < nk; ++i, ++it) {
          if (it >= npre) {
            if (it >= STG) mbar_wait(bar0 + 8 * (STG + s), ph ^ 1u);
            mbar_expect_tx(bar0 + 8 * s, stage_bytes);
            tma_load_2d(smem0 + s * stage_bytes + RG_A_BYTES, &P.tmW, bar0 + 8 * s, (kb + i) * 64, T.col0);
          }
          tma_load_3d(smem0 + s * stage_bytes, &P.tmA, bar0 + 8 * s, (kb + i) * 64, T.row0, 0);
          if (++s == STG) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform issue loop, one elected lane issues (see elect_one() in tc.cuh)
    const uint64_t adesc0 = umma_desc_sw128(smem0, 0, 1024);
    const uint64_t bdesc0 = adesc0 + (uint64_t)(RG_A_BYTES >> 4);
    const uint32_t idesc = batch.idesc;
    int s = 0;
    uint32_t ph = 0, li = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++li) {
      const RgTile T = rg_decode(batch, t);
      const int nk = ((batch.p[T.prob].K + 63) / 64) / batch.ksplit;
      const uint32_t buf = li & 1u;
      mbar_wait(acc_empty0 + 8 * buf, ((li >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t dt = tmem_base + buf * (uint32_t)BN;
      for (int i = 0; i < nk; ++i) {
        mbar_wait(bar0 + 8 * s, ph);
        tc_fence_after();
        if (elect_one()) {
          if (li == 0 && i == 0) RG_TS(3);
          const uint64_t so = (uint64_t)((uint32_t)s * (stage_bytes >> 4));
#pragma unroll
          for (int pl = 2; pl >= 0; --pl) {                 // lo, mid, hi: small terms first
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(dt, adesc0 + so + (uint64_t)(pl * (int)(RG_A_PLANE >> 4) + k * 2), bdesc0 + so + (uint64_t)(k * 2), idesc,
                        (i > 0 || pl < 2 || k > 0) ? 1u : 0u);
          }
          umma_commit(bar0 + 8 * (STG + s));
          if (i == nk - 1) {
            umma_commit(acc_full0 + 8 * buf);
            if (li == 0) RG_TS(4);
          }
        }
        __syncwarp();
        if (++s == STG) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else {
    // ---- epilogue: 8 warps, thread = row (TMEM lane); the two warps of a lane quarter take alternate 32-column
    //      blocks.  A thread owns 32 consecutive columns of its row in registers: bias / residual / ReLU / the plane
    //      split run as 32 independent chains and leave as 16-byte stores (no shared-memory round trip).
    pdl_wait();                                             // residual / rowscale reads, and every global store
    uint32_t li = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++li) {
      const RgTile T = rg_decode(batch, t);
      const RgProb &P = batch.p[T.prob];
      const uint32_t buf = li & 1u;
      mbar_wait(acc_full0 + 8 * buf, (li >> 1) & 1u);
      tc_fence_after();
      if (t + (int)gridDim.x >= total) pdl_trigger();
      if (threadIdx.x == 64 && li == 0) RG_TS(5);
      rg_epilogue_tile(P, T.row0, T.col0, T.ks, BN, tmem_base + buf * (uint32_t)BN, acc_empty0 + 8 * buf,
                       stg0 + (uint32_t)(warp - 2) * (uint32_t)batch.stg_bytes, (uint32_t)batch.pl_off,
                       ln_stat + (size_t)(li & 1u) * 128 * 8, warp, lane, RG_THREADS - 64);
    }
    if (lane == 0) bulk_wait_group_read<0>();               // staging may be released; the writes drain with the grid
    if (threadIdx.x == 64) RG_TS(6);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ncols);
  }
  if (threadIdx.x == 0) RG_TS(7);
}

static int rg_env(const char *name, int dflt) {
  const char *e = getenv(name);
  return (e && atoi(e) > 0) ? atoi(e) : dflt;
}

bool linear_tc_supported(const LinArgs &a) {
  if (a.src.pro != PRO_PLANES) return false;
  if (a.K % 8 != 0 || a.ldw % 8 != 0 || a.src.lda[0] % 8 != 0 || a.src.sum_stride % 8 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(a.w) & 15) || (reinterpret_cast<uintptr_t>(a.src.a[0]) & 15)) return false;
  return true;
}

// LinArgs (PRO_PLANES input, bf16 weights) -> the device-side problem record: tensor maps of the A planes, the weight
// matrix and (when describable) the outputs, for column tiles of BN and ks K-slices.
static int rg_fill_prob(const LinArgs &a, int BN, int ks, RgProb &p) {
  const uint64_t adims[3] = {(uint64_t)a.K, (uint64_t)a.M, 3};
  const uint64_t astr[2] = {(uint64_t)a.src.lda[0] * 2, (uint64_t)a.src.sum_stride * 2};
  const uint32_t abox[3] = {64u, 128u, 3u};
  VKN_TRY(make_tmap_bf16_strided(&p.tmA, a.src.a[0], 3, adims, astr, abox));
  const uint64_t wdims[2] = {(uint64_t)a.K, (uint64_t)a.N};
  const uint64_t wstr[1] = {(uint64_t)a.ldw * 2};
  const uint32_t wbox[2] = {64u, (uint32_t)BN};
  VKN_TRY(make_tmap_bf16_strided(&p.tmW, a.w, 2, wdims, wstr, wbox));
  p.tma_out = 0;
  p.tma_pl = 0;
  const bool tma_ok = getenv("VKN_RG_TMA_STORE") == nullptr || getenv("VKN_RG_TMA_STORE")[0] != '0';
  if (tma_ok && !(a.epi & EPI_NOOUT) && a.out && a.N >= 32 && a.ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 &&
      (ks == 1 || a.out_split_stride % 4 == 0)) {
    const uint64_t od[3] = {(uint64_t)a.N, (uint64_t)a.M, (uint64_t)ks};
    const uint64_t os[2] = {(uint64_t)a.ldo * 4, (uint64_t)(ks > 1 ? a.out_split_stride : (long long)a.M * a.ldo) * 4};
    const uint32_t ob[3] = {32u, 32u, 1u};
    VKN_TRY(make_tmap_store(&p.tmOut, a.out, true, od, os, ob, true));
    p.tma_out = 1;
  }
  if (tma_ok && (a.epi & EPI_SPLIT3) && !(a.epi & EPI_SPLIT2H) && a.split_planes && a.split_N == a.split_Npad && a.split_C % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(a.split_planes) & 15) == 0 && ((long long)a.split_B * a.split_Npad * a.split_C) % 8 == 0) {
    const uint64_t pd[3] = {(uint64_t)a.split_C, (uint64_t)a.M, 3};
    const uint64_t ps[2] = {(uint64_t)a.split_C * 2, (uint64_t)((long long)a.split_B * a.split_Npad * a.split_C) * 2};
    const uint32_t pb[3] = {32u, 32u, 3u};
    VKN_TRY(make_tmap_store(&p.tmPl, a.split_planes, false, pd, ps, pb, false));
    p.tma_pl = 1;
  }
  p.bias = a.bias;
  p.rowscale = a.rowscale;
  p.res = a.res;
  p.ln_g = a.ln_g;
  p.ln_b = a.ln_b;
  if ((a.epi & EPI_LN) && (!a.ln_g || !a.ln_b)) VKN_FAIL(VKN_E_INVALID, "launch_linear_tc: EPI_LN without LayerNorm parameters");
  p.out = a.out;
  p.planes = a.split_planes;
  p.out_split_stride = a.out_split_stride;
  p.plane_elems = (long long)a.split_B * a.split_Npad * a.split_C;
  p.ldres = a.ldres;
  p.mul = a.mul;
  p.ldmul = a.ldmul;
  p.add2 = a.add2;
  p.ldadd2 = a.ldadd2;
  if (((a.epi & EPI_MUL) && (!a.mul || a.ldmul % 4 || (reinterpret_cast<uintptr_t>(a.mul) & 15))) ||
      ((a.epi & EPI_ADD2) && (!a.add2 || a.ldadd2 % 4 || (reinterpret_cast<uintptr_t>(a.add2) & 15))))
    VKN_FAIL(VKN_E_INVALID, "launch_linear_tc: EPI_MUL / EPI_ADD2 need 16-byte aligned fp32 rows");
  if ((a.epi & (EPI_MUL | EPI_ADD2 | EPI_SIGMOID)) && a.N % 4 != 0)
    VKN_FAIL(VKN_E_INVALID, "launch_linear_tc: elementwise epilogue operands need N %% 4 == 0");
  p.ldo = a.ldo;
  p.M = a.M;
  p.N = a.N;
  p.K = a.K;
  p.epi = a.epi;
  p.split_N = a.split_N > 0 ? a.split_N : 1;
  p.split_Npad = a.split_Npad;
  p.split_C = a.split_C;
  p.mt = ceil_div(a.M, 128);
  p.nt = ceil_div(a.N, BN);
  p.tiles = p.mt * p.nt * ks;
  if ((a.epi & EPI_BIAS) && ks > 1) VKN_FAIL(VKN_E_INVALID, "launch_linear_tc: bias with split-K belongs to the consumer");
  if ((a.epi & EPI_SPLIT3) && !a.split_planes) VKN_FAIL(VKN_E_INVALID, "launch_linear_tc: EPI_SPLIT3 without a plane buffer");
  if (!(a.epi & EPI_NOOUT) && !a.out) VKN_FAIL(VKN_E_INVALID, "launch_linear_tc: null output");
  return VKN_OK;
}

// Same contract as launch_linear (smallops.cu) for PRO_PLANES inputs and bf16 weights.
int launch_linear_tc(const LinArgs *probs, int nprob, cudaStream_t stream) {
  if (nprob < 1 || nprob > 2) VKN_FAIL(VKN_E_INVALID, "launch_linear_tc: nprob %d", nprob);
  RgBatch b;
  memset(&b, 0, sizeof(b));
  int maxM = 0, maxN = 0, nkmax = 0;
  const int ks = probs[0].ksplit < 1 ? 1 : probs[0].ksplit;
  for (int i = 0; i < nprob; ++i) {
    const LinArgs &a = probs[i];
    if (!linear_tc_supported(a))
      VKN_FAIL(VKN_E_INVALID, "launch_linear_tc: needs bf16 plane inputs with 16-byte aligned rows (K %d, ldw %d, lda %d)", a.K,
               a.ldw, a.src.lda[0]);
    if ((a.ksplit < 1 ? 1 : a.ksplit) != ks) VKN_FAIL(VKN_E_INVALID, "launch_linear_tc: batched problems must share ksplit");
    const int nk = ceil_div(a.K, 64);
    if (nk % ks != 0) VKN_FAIL(VKN_E_INVALID, "launch_linear_tc: K = %d does not split into %d slices of 64-blocks", a.K, ks);
    if (a.side != nullptr) VKN_FAIL(VKN_E_INVALID, "launch_linear_tc: side outputs are not supported");
    maxM = max(maxM, a.M);
    maxN = max(maxN, a.N);
    nkmax = max(nkmax, nk / ks);
  }
  // Column tile: the candidate with the lowest estimated launch time = waves x tile time.  Tile time in cycles of the
  // shared-memory data pipe (the limiter): per 64-deep K chunk the TMA writes 48 KB of A planes + BN x 128 B of W and
  // the 12 MMAs read 12 x (4 KB + BN x 32 B); plus a fixed prologue/epilogue cost.  Few tiles -> narrow tiles (spread the
  // work), many tiles -> BN = 256 (A planes read once per 256 output columns).
  const int sms = 148;
  int BN = 32;
  {
    double best = 1e30;
    for (int cand = 32; cand <= 256; cand <<= 1) {
      if (cand > 32 && cand / 2 >= maxN) break;                     // tile wider than the problem
      long long tiles = 0;
      for (int i = 0; i < nprob; ++i) tiles += (long long)ceil_div(probs[i].M, 128) * ceil_div(probs[i].N, cand) * ks;
      const double tile_cyc = (double)nkmax * (768.0 + 4.0 * cand) + 1200.0 + 6.0 * cand;
      const double cost = (double)((tiles + sms - 1) / sms) * tile_cyc;
      if (cost < best) {
        best = cost;
        BN = cand;
      }
    }
  }
  BN = rg_env("VKN_RG_BN", BN);
  bool fused_ln = false;
  for (int i = 0; i < nprob; ++i) fused_ln = fused_ln || (probs[i].epi & EPI_LN);
  if (fused_ln) {                                                   // the tile must span the whole row
    if (maxN > 256 || ks != 1) VKN_FAIL(VKN_E_INVALID, "launch_linear_tc: EPI_LN needs N <= 256 and no split-K (N %d, ksplit %d)", maxN, ks);
    BN = maxN <= 32 ? 32 : (maxN <= 64 ? 64 : (maxN <= 128 ? 128 : 256));
  }
  if (BN != 32 && BN != 64 && BN != 128 && BN != 256) VKN_FAIL(VKN_E_INVALID, "VKN_RG_BN must be 32, 64, 128 or 256");
  const size_t stage_bytes = (size_t)RG_A_BYTES + (size_t)BN * 128;
  int stages = rg_env("VKN_RG_STAGES", 4);
  int stg_bytes = 10240;                                              // fp32 box + plane boxes side by side when they fit
  auto smem_of = [&](int st) { return (size_t)st * stage_bytes + 8 * (size_t)stg_bytes + 1024 + (2 * st + 4) * 8 + 16 + 2 * 128 * 8 * 4 + 64; };
  if (smem_of(2) > 227 * 1024) stg_bytes = 6144;                      // BN = 256: keep two pipeline stages, share the box
  while (stages > 1 && smem_of(stages) > 227 * 1024) --stages;
  size_t smem = smem_of(stages);
  b.stg_bytes = stg_bytes;
  b.pl_off = stg_bytes == 10240 ? 4096 : 0;
  b.nprob = nprob;
  b.ksplit = ks;
  b.stages = stages;
  b.BN = BN;
  b.idesc = make_idesc_bf16(128, BN, 0, 0);
  b.dbg = debug_ts_slot();
  for (int i = 0; i < nprob; ++i) {
    VKN_TRY(rg_fill_prob(probs[i], BN, ks, b.p[i]));
    b.total_tiles += b.p[i].tiles;
  }
  if (nprob == 1) {
    b.p[1] = b.p[0];
    b.p[1].tiles = 0;
  }
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) {
    VKN_CUDA_OK(cudaFuncSetAttribute(vkn_rowgemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  dim3 grid(b.total_tiles < sms ? b.total_tiles : sms);
  VKN_LAUNCH_MARK("vkn_rowgemm_tc_kernel", stream);
  VKN_CUDA_OK(launch_chain(vkn_rowgemm_tc_kernel, grid, dim3(RG_THREADS), smem, stream, b));
  return VKN_OK;
}


// =====================================================================================================================
// Chain kernel: a whole SEQUENCE of row operators of a stage in ONE launch.
//
// The row operators of a stage (KernelUpdator -> in-proj | out-proj -> FFN -> FC heads -> fold; knet/kernel_updator.py:56-94,
// knet/det/kernel_update_head.py:201-227) only ever couple the rows of one kernel set through the attention, so between
// two attentions a 128-row tile depends on nothing but itself.  Launched as ~14 separate GEMM / row kernels each step ran
// one TMA -> MMA -> epilogue pass per CTA and then waited for the whole grid (measured: 61 launches, 59 % of the step for
// 9 % of its FLOPs).  Here a CTA owns a row tile and walks the step program by itself:
//   GEMM step  = up to two problems x column tiles of 256, the same TMA ring / tcgen05 / double-buffered TMEM pipeline and
//                the same epilogue as vkn_rowgemm_tc_kernel (fused LayerNorm, planes for the next GEMM, TMA stores);
//   row step   = a row transform by the epilogue warps (gate arithmetic: 4 LayerNorms + 2 sigmoids, products) -> planes;
//   step edge  = the epilogue warps drain their TMA stores, fence the proxies, meet on a named barrier and signal
//                `step_done`; the TMA warp loads the next step's A planes (this tile's rows, just written, L2-resident)
//                only after that, while its weight tile is already in flight.
// Nothing but the CTA's own 384 threads ever waits: no grid-wide dependency, no launch gap, weights stream from L2.
constexpr int CH_MAXP = 16, CH_MAXR = 4, CH_MAXS = 14;
struct ChRow {
  RowSrc src;
  float *out;
  __nv_bfloat16 *planes;
  long long plane_stride;
  int ldo, ldp, K, pad_;
};
struct ChStep {
  int kind, n, first, pad_;      // kind 0: GEMM problems p[first, first + n); kind 1: row op r[first]
};
struct ChainProg {
  RgProb p[CH_MAXP];
  ChRow r[CH_MAXR];
  ChStep s[CH_MAXS];
  int nsteps, M, mtiles, stages, stg_bytes, pl_off;
  int tpc;                       // row tiles a CTA interleaves per pass of the program: 2 (pairs) or 1
  uint32_t idesc;
  unsigned long long *dbg;       // optional: %globaltimer at every step edge of the CTA's first tile (vkn_debug_timestamps)
};
static_assert(sizeof(ChainProg) <= 32000, "chain program must fit the kernel parameter space");

constexpr int CH_BN = 256;

// ---- chain-kernel epilogue -----------------------------------------------------------------------------------------
// Same arithmetic as rg_epilogue_tile, organised around the latency that bounded it (measured, tools/epi_prof.py: every
// global load an epilogue thread waits for costs ~1.2 us while the TMA ring keeps the L2 path busy, and there were 1-3 per
// 32-column block):
//   * bias / LayerNorm gamma / beta of the tile are staged in shared memory ONCE per tile, before the accumulator wait;
//   * the per-row operand of a block (residual in the first pass, elementwise factor in the second) is prefetched into
//     registers one block ahead: its latency hides behind the store phase of the previous block (or the accumulator wait).
// ch_epilogue_pre runs before the wait for the accumulator, ch_epilogue_tile after it.
struct ChPre {
  float ra[32];          // prefetched operand block (residual or factor) of the thread's row
};

__device__ __forceinline__ void ch_load_op(const float *base, int ld, int row, int col, int nc, bool live, float (&ra)[32]) {
  if (!live) return;
  const float *p = base + (size_t)row * ld + col;
  if (nc == 32) {
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      const float4 t4 = __ldcg(reinterpret_cast<const float4 *>(p + e));
      ra[e] = t4.x; ra[e + 1] = t4.y; ra[e + 2] = t4.z; ra[e + 3] = t4.w;
    }
  } else {
#pragma unroll
    for (int e = 0; e < 32; ++e) ra[e] = (e < nc) ? __ldcg(p + e) : 0.f;
  }
}

// vec_s: [3][256] floats (bias, gamma, beta of columns col0 .. col0+255).  All 256 epilogue threads.
__device__ __forceinline__ void ch_epilogue_pre(const RgProb &P, int row0, int col0, float *vec_s, int warp, int lane, ChPre &pre) {
  const int t = (warp - 2) * 32 + lane, col = col0 + t;
  named_bar_sync(1, RG_THREADS - 64);                    // every warp is done with the previous tile's vectors
  vec_s[t] = ((P.epi & EPI_BIAS) && col < P.N) ? __ldg(P.bias + col) : 0.f;
  if (P.epi & EPI_LN) {
    vec_s[256 + t] = col < P.N ? __ldg(P.ln_g + col) : 0.f;
    vec_s[512 + t] = col < P.N ? __ldg(P.ln_b + col) : 0.f;
  }
  const int q = warp & 3, half = (warp - 2) >> 2;
  const int row = row0 + q * 32 + lane, c = col0 + half * 32;
  if (c < P.N) {
    const int nc = min(32, P.N - c);
    if (P.epi & EPI_RES) ch_load_op(P.res, P.ldres, row, c, nc, row < P.M, pre.ra);
    else if ((P.epi & EPI_MUL) && !(P.epi & EPI_LN)) ch_load_op(P.mul, P.ldmul, row, c, nc, row < P.M, pre.ra);
  }
}

__device__ __forceinline__ void ch_epilogue_tile(const RgProb &P, int row0, int col0, uint32_t tacc_base, uint32_t acc_empty,
                                                 uint32_t stg, uint32_t pl_off, float *st_base, const float *vec_s, int warp,
                                                 int lane, ChPre &pre) {
  constexpr int BN = CH_BN;
  const int q = warp & 3, half = (warp - 2) >> 2;
  RgRowCtx R;
  R.epi = P.epi;
  R.row = row0 + q * 32 + lane;
  R.live = R.row < P.M;
  R.row_base = row0 + q * 32;
  R.ks = 0;
  R.stg = stg;
  R.pl_off = pl_off;
  R.outp = P.out;
  R.out_vec = (P.ldo % 8 == 0) && ((reinterpret_cast<uintptr_t>(R.outp) & 31) == 0);
  R.res_vec = true;
  R.pl_vec = (P.split_C % 16 == 0) && ((reinterpret_cast<uintptr_t>(P.planes) & 31) == 0) && (P.plane_elems % 16 == 0);
  R.bias_vec = true;
  R.rs = (R.live && (R.epi & EPI_ROWSCALE)) ? __ldcg(P.rowscale + R.row) : 1.f;
  R.prow = (size_t)R.row;
  if ((R.epi & EPI_SPLIT3) && P.split_N != P.split_Npad) {
    const int b = R.row / P.split_N;
    R.prow = (size_t)b * P.split_Npad + (R.row - b * P.split_N);
  }
  named_bar_sync(1, RG_THREADS - 64);                    // the staged vectors are visible
  const uint32_t tacc = tacc_base + ((uint32_t)(q * 32) << 16);
  int last_c0 = half * 32;
  while (last_c0 + 64 < BN && col0 + last_c0 + 64 < P.N) last_c0 += 64;
  const bool any = half * 32 < BN && col0 + half * 32 < P.N;
  const bool has_res = (R.epi & EPI_RES) != 0, has_mul = (R.epi & EPI_MUL) != 0;

  // v = acc + rowscale * bias (+ prefetched residual); then prefetch the next block's operand
  auto load_block = [&](int c0, float (&v)[32]) {
    uint32_t r[32];
    tmem_ld32(tacc + (uint32_t)c0, r);
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      const float4 b4 = *reinterpret_cast<const float4 *>(vec_s + c0 + e);
      v[e] = fmaf(R.rs, b4.x, __uint_as_float(r[e]));
      v[e + 1] = fmaf(R.rs, b4.y, __uint_as_float(r[e + 1]));
      v[e + 2] = fmaf(R.rs, b4.z, __uint_as_float(r[e + 2]));
      v[e + 3] = fmaf(R.rs, b4.w, __uint_as_float(r[e + 3]));
    }
  };
  // sigmoid / factor (prefetched) / addend, then ReLU + stores
  auto finish_block = [&](int c0, int col, int nc, float (&v)[32], bool mul_next_valid, int next_c0) {
    if (R.live) {
      if (R.epi & EPI_SIGMOID) {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = sigmoid_fast(v[e]);
      }
      if (has_mul) {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] *= pre.ra[e];
      }
    }
    if (has_mul && mul_next_valid) {
      const int ncol = col0 + next_c0;
      ch_load_op(P.mul, P.ldmul, R.row, ncol, min(32, P.N - ncol), R.live, pre.ra);
    }
    if ((R.epi & EPI_ADD2) && R.live) {
      const float *ap = P.add2 + (size_t)R.row * P.ldadd2 + col;
#pragma unroll
      for (int e = 0; e < 32; e += 4)
        if (e < nc) {
          const float4 t4 = __ldcg(reinterpret_cast<const float4 *>(ap + e));
          v[e] += t4.x; v[e + 1] += t4.y; v[e + 2] += t4.z; v[e + 3] += t4.w;
        }
    }
    rg_chunk_store(P, R, col, nc, v);
  };

  if (R.epi & EPI_LN) {
    float K0 = 0.f, S1 = 0.f, S2 = 0.f, cntf = 0.f;
    bool first = true;
    for (int c0 = half * 32; c0 < BN; c0 += 64) {
      const int col = col0 + c0;
      if (col >= P.N) break;
      const int nc = min(32, P.N - col);
      float v[32];
      load_block(c0, v);
      if (has_res) {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] += pre.ra[e];
      }
      // next operand block: the residual of the next block of this pass, or the factor of the first block of pass 2
      if (c0 + 64 < BN && col + 64 < P.N) {
        if (has_res) ch_load_op(P.res, P.ldres, R.row, col + 64, min(32, P.N - col - 64), R.live, pre.ra);
      } else if (has_mul) {
        ch_load_op(P.mul, P.ldmul, R.row, col0 + half * 32, min(32, P.N - col0 - half * 32), R.live, pre.ra);
      }
      {   // keep acc + bias + residual in the accumulator itself: the second pass re-reads TMEM, not global memory
        uint32_t vr[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) vr[e] = __float_as_uint(v[e]);
        tmem_st32(tacc + (uint32_t)c0, vr);
      }
      if (first) {
        K0 = v[0];
        first = false;
      }
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (e < nc) {
          const float d = v[e] - K0;
          S1 += d;
          S2 = fmaf(d, d, S2);
        }
      cntf += (float)nc;
    }
    float mean_h = 0.f, m2_h = 0.f;
    if (cntf > 0.f) {
      mean_h = K0 + S1 / cntf;
      m2_h = S2 - S1 * S1 / cntf;
    }
    float *st = st_base + (size_t)(q * 32 + lane) * 8;
    st[half * 4 + 0] = cntf;
    st[half * 4 + 1] = mean_h;
    st[half * 4 + 2] = m2_h;
    named_bar_sync(1, RG_THREADS - 64);
    const float cb = st[(half ^ 1) * 4 + 0], mb = st[(half ^ 1) * 4 + 1], m2b = st[(half ^ 1) * 4 + 2];
    const float n0 = half ? cb : cntf, mu0 = half ? mb : mean_h, q0 = half ? m2b : m2_h;
    const float n1 = half ? cntf : cb, mu1 = half ? mean_h : mb, q1 = half ? m2_h : m2b;
    const float nn = n0 + n1, delta = mu1 - mu0;
    const float mean = mu0 + delta * (n1 / nn);
    const float var = (q0 + q1 + delta * delta * (n0 * n1 / nn)) / nn;
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    for (int c0 = half * 32; c0 < BN; c0 += 64) {
      const int col = col0 + c0;
      if (col >= P.N) break;
      const int nc = min(32, P.N - col);
      float v[32];
      {
        uint32_t vr[32];
        tmem_ld32(tacc + (uint32_t)c0, vr);
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(vr[e]);
      }
      if (c0 == last_c0) {                              // second read done: hand the accumulator back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
      }
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 g4 = *reinterpret_cast<const float4 *>(vec_s + 256 + c0 + e);
        const float4 b4 = *reinterpret_cast<const float4 *>(vec_s + 512 + c0 + e);
        v[e] = fmaf((v[e] - mean) * rstd, g4.x, b4.x);
        v[e + 1] = fmaf((v[e + 1] - mean) * rstd, g4.y, b4.y);
        v[e + 2] = fmaf((v[e + 2] - mean) * rstd, g4.z, b4.z);
        v[e + 3] = fmaf((v[e + 3] - mean) * rstd, g4.w, b4.w);
      }
      finish_block(c0, col, nc, v, c0 + 64 < BN && col + 64 < P.N, c0 + 64);
    }
  } else {
    for (int c0 = half * 32; c0 < BN; c0 += 64) {
      const int col = col0 + c0;
      if (col >= P.N) break;
      const int nc = min(32, P.N - col);
      float v[32];
      load_block(c0, v);
      if (c0 == last_c0) {                              // this warp's share of the accumulator is in registers
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
      }
      const bool more = c0 + 64 < BN && col + 64 < P.N;
      if (has_res) {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] += pre.ra[e];
        if (more) ch_load_op(P.res, P.ldres, R.row, col + 64, min(32, P.N - col - 64), R.live, pre.ra);
      }
      finish_block(c0, col, nc, v, more, c0 + 64);
    }
  }
  if (!any) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(acc_empty);
  }
}


// one row transform of the chain: warp per row, rows of the tile dealt round-robin to the 8 epilogue warps
__device__ __forceinline__ void ch_row_step(const ChRow &Rw, int row0, int M, int ew, int lane) {
  for (int rr = ew; rr < 128; rr += 8) {
    const int row = row0 + rr;
    if (row >= M) break;
    RowRaw r;
    float v[KPL];
    row_load<4, true>(Rw.src, row, 0, Rw.K, lane, r);
    row_finish<4>(Rw.src, Rw.K, lane, r, v);
    if (Rw.out != nullptr) store_pairs(Rw.out + (size_t)row * Rw.ldo, Rw.K, lane, v);
    if (Rw.planes != nullptr) {
#pragma unroll
      for (int p = 0; p < KPL / 2; ++p) {
        const int k = kidx(lane, 2 * p);
        float x0 = v[2 * p], x1 = v[2 * p + 1];
#pragma unroll
        for (int t = 0; t < 3; ++t) {          // hi, then the residuals: v == hi + mid + lo to 24 bits
          const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
          x0 -= __bfloat162float(h0);
          x1 -= __bfloat162float(h1);
          if (k < Rw.K)
            *reinterpret_cast<uint32_t *>(Rw.planes + (size_t)t * Rw.plane_stride + (size_t)row * Rw.ldp + k) =
                (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        }
      }
    }
  }
}

// spin until the CTA's epilogue warps have completed `need` steps (bounded: a protocol bug traps instead of hanging)
__device__ __forceinline__ void ch_wait_steps(uint32_t ctr, uint32_t need) {
  for (uint32_t it = 0; it < SPIN_LIMIT; ++it) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(ctr) : "memory");
    if (v >= need) return;
  }
  __trap();
}

__global__ void __launch_bounds__(RG_THREADS, 1) vkn_chain_tc_kernel(const __grid_constant__ ChainProg prog) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int STG = prog.stages, nsteps = prog.nsteps, tpc = prog.tpc, npairs = (prog.mtiles + tpc - 1) / tpc;
  constexpr uint32_t stage_bytes = RG_A_BYTES + (uint32_t)CH_BN * 128u;
  const uint32_t stg0 = smem_u32(smem + (size_t)STG * stage_bytes);
  uint64_t *bars = (uint64_t *)(smem + (size_t)STG * stage_bytes + 8 * (size_t)prog.stg_bytes);
  const uint32_t bar0 = smem_u32(bars);
  // full[s] = bar0 + 8 s, empty[s] = bar0 + 8 (STG + s), acc_full[h], acc_empty[h]; then the two step counters
  const uint32_t acc_full0 = bar0 + 16 * STG, acc_empty0 = acc_full0 + 16;
  // steps completed by row tile X / Y of the current pair (monotonic counters: a parity barrier could be lapped by two
  // back-to-back row steps)
  const uint32_t step_ctr = acc_empty0 + 16;
  uint32_t *tmem_slot = (uint32_t *)(bars + 2 * STG + 5);
  float *ln_stat = (float *)(tmem_slot + 4);                 // [2][128 rows][8]: fused-LayerNorm partials (8 KB)
  const uint32_t smem0 = smem_u32(smem);
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  // vkn_debug_timestamps: 64 words per CTA -- [0] start, [1 + 2 si + h] edge of step si / tile h (first pair), [32..] cycle
  // counters of the three roles (waits and work), see tools/chain_timeline.py
  unsigned long long *dbg = prog.dbg ? prog.dbg + (size_t)blockIdx.x * 64 : nullptr;
#define CH_ACC(acc, stmt)               \
  do {                                  \
    if (dbg != nullptr) {               \
      const long long c0__ = clock64(); \
      stmt;                             \
      acc += clock64() - c0__;          \
    } else {                            \
      stmt;                             \
    }                                   \
  } while (0)
  // epilogue-side counters live in shared memory (updated by one thread): registers are the scarce resource there
  long long *epi_ctr = reinterpret_cast<long long *>(ln_stat + 2 * 128 * 8);   // 4 counters behind the LayerNorm scratch
  // [3][256]: bias / gamma / beta of the current tile (16-byte aligned: read as float4)
  float *vec_s = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(epi_ctr + 4) + 15) & ~(uintptr_t)15);
#define CH_ACC_S(slot, stmt)                                        \
  do {                                                              \
    if (dbg != nullptr) {                                           \
      const long long c0__ = clock64();                             \
      stmt;                                                         \
      if (threadIdx.x == 64) epi_ctr[slot] += clock64() - c0__;     \
    } else {                                                        \
      stmt;                                                         \
    }                                                               \
  } while (0)
  if (threadIdx.x < 4) epi_ctr[threadIdx.x] = 0;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < STG; ++s) {
        mbar_init(bar0 + 8 * s, 1);
        mbar_init(bar0 + 8 * (STG + s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(acc_full0 + 8 * a, 1);
        mbar_init(acc_empty0 + 8 * a, 8);                   // one arrive per epilogue warp
      }
      asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(step_ctr), "r"(0u) : "memory");
      fence_barrier_init();
    }
  } else if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 2u * CH_BN);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                               // the first step reads what the previous kernel wrote
  if (dbg != nullptr && threadIdx.x == 0) dbg[0] = rg_time();

  // A CTA owns PAIRS of row tiles (X = 2p, Y = 2p + 1) and interleaves them through every step: while the epilogue warps
  // drain X's accumulator (TMEM buffer 0) the tensor core multiplies Y's (buffer 1) and vice versa, and a tile's step
  // edge (stores complete -> the next step's A planes loaded back from L2) hides behind the other tile's work.  An odd
  // last tile pairs with a phantom (rows >= M: zero-filled loads, no stores).
  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0, it = 0, g = 0;
      long long c_empty = 0, c_steps = 0;
      for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
        for (int si = 0; si < nsteps; ++si, ++g) {
          const int kind = prog.s[si].kind, np = prog.s[si].n, first = prog.s[si].first;
          if (kind != 0) continue;                          // row step: nothing to load
          for (int pi = 0; pi < np; ++pi) {
            const RgProb &P = prog.p[first + pi];
            const int nk = (P.K + 63) / 64;
            for (int j = 0; j < P.nt; ++j)
              for (int h = 0; h < tpc; ++h)
                for (int i = 0; i < nk; ++i, ++it) {
                  if (it >= (uint32_t)STG) CH_ACC(c_empty, mbar_wait(bar0 + 8 * (STG + s), ph ^ 1u));
                  mbar_expect_tx(bar0 + 8 * s, stage_bytes);
                  tma_load_2d(smem0 + s * stage_bytes + RG_A_BYTES, &P.tmW, bar0 + 8 * s, i * 64, j * CH_BN);
                  CH_ACC(c_steps, ch_wait_steps(step_ctr + 4 * h, g));   // this tile's A planes were written by its previous steps
                  tma_load_3d(smem0 + s * stage_bytes, &P.tmA, bar0 + 8 * s, i * 64, (tpc * pr + h) * 128, 0);
                  if (++s == STG) {
                    s = 0;
                    ph ^= 1u;
                  }
                }
          }
        }
      }
      if (dbg != nullptr) {
        dbg[48] = (unsigned long long)c_empty;
        dbg[49] = (unsigned long long)c_steps;
      }
    }
  } else if (warp == 1) {
    const uint64_t adesc0 = umma_desc_sw128(smem0, 0, 1024);
    const uint64_t bdesc0 = adesc0 + (uint64_t)(RG_A_BYTES >> 4);
    const uint32_t idesc = prog.idesc;
    int s = 0;
    uint32_t ph = 0, li = 0;                                // li counts accumulator uses: even = tile X, odd = tile Y
    long long c_full = 0, c_acc = 0;
    for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
      for (int si = 0; si < nsteps; ++si) {
        if (prog.s[si].kind != 0) continue;
        const int np = prog.s[si].n, first = prog.s[si].first;
        for (int pi = 0; pi < np; ++pi) {
          const int nk = (prog.p[first + pi].K + 63) / 64, nt = prog.p[first + pi].nt;
          for (int jh = 0; jh < tpc * nt; ++jh, ++li) {
            const uint32_t buf = li & 1u;
            CH_ACC(c_acc, mbar_wait(acc_empty0 + 8 * buf, ((li >> 1) & 1u) ^ 1u));
            tc_fence_after();
            const uint32_t dt = tmem_base + buf * (uint32_t)CH_BN;
            for (int i = 0; i < nk; ++i) {
              CH_ACC(c_full, mbar_wait(bar0 + 8 * s, ph));
              tc_fence_after();
              if (elect_one()) {
                const uint64_t so = (uint64_t)((uint32_t)s * (stage_bytes >> 4));
#pragma unroll
                for (int pl = 2; pl >= 0; --pl) {             // lo, mid, hi: small terms first
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16(dt, adesc0 + so + (uint64_t)(pl * (int)(RG_A_PLANE >> 4) + k * 2), bdesc0 + so + (uint64_t)(k * 2), idesc,
                              (i > 0 || pl < 2 || k > 0) ? 1u : 0u);
                }
                umma_commit(bar0 + 8 * (STG + s));
                if (i == nk - 1) umma_commit(acc_full0 + 8 * buf);
              }
              __syncwarp();
              if (++s == STG) {
                s = 0;
                ph ^= 1u;
              }
            }
          }
        }
      }
    }
    if (dbg != nullptr && lane == 0) {
      dbg[40] = (unsigned long long)c_full;
      dbg[41] = (unsigned long long)c_acc;
    }
  } else {
    const int ew = warp - 2;
    const uint32_t stg = stg0 + (uint32_t)ew * (uint32_t)prog.stg_bytes;
    uint32_t li = 0, done = 0;
    bool first_pair = true;
    // step edge of row tile h: everything it wrote is complete and visible to the TMA loads / L2 reads of its next step
#define CH_EDGE(h, si)                                                                                                  \
  do {                                                                                                                  \
    const long long e0 = dbg ? clock64() : 0;                                                                           \
    if (lane == 0) bulk_wait_all();                                                                                     \
    __syncwarp();                                                                                                       \
    asm volatile("fence.proxy.async;" ::: "memory");                                                                    \
    named_bar_sync(1, RG_THREADS - 64);                                                                                 \
    if (threadIdx.x == 64) {                                                                                            \
      asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(step_ctr + 4 * (h)), "r"(done + 1) : "memory");     \
      if (dbg != nullptr) {                                                                                             \
        epi_ctr[2] += clock64() - e0;                                                                                   \
        if (first_pair && (si) < 15) dbg[1 + 2 * (si) + (h)] = rg_time();                                               \
      }                                                                                                                 \
    }                                                                                                                   \
  } while (0)
    for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
      for (int si = 0; si < nsteps; ++si, ++done) {
        const int kind = prog.s[si].kind, np = prog.s[si].n, first = prog.s[si].first;
        if (kind == 0) {
          for (int pi = 0; pi < np; ++pi) {
            const RgProb &P = prog.p[first + pi];
            for (int j = 0; j < P.nt; ++j) {
              const bool last = (pi == np - 1) && (j == P.nt - 1);
              for (int h = 0; h < tpc; ++h, ++li) {
                const uint32_t buf = li & 1u;
                ChPre pre;
                if (P.epi & (EPI_MUL | EPI_ADD2)) {
                  // the factor / addend rows may come from an earlier tile of the SAME step: this very warp stored them
                  // (same rows, same column blocks) through TMA -- drain its bulk stores before reading them back
                  if (lane == 0) bulk_wait_all();
                  __syncwarp();
                }
                ch_epilogue_pre(P, (tpc * pr + h) * 128, j * CH_BN, vec_s, warp, lane, pre);
                CH_ACC_S(0, mbar_wait(acc_full0 + 8 * buf, (li >> 1) & 1u));
                tc_fence_after();
                CH_ACC_S(1, ch_epilogue_tile(P, (tpc * pr + h) * 128, j * CH_BN, tmem_base + buf * (uint32_t)CH_BN, acc_empty0 + 8 * buf,
                                 stg, (uint32_t)prog.pl_off, ln_stat + (size_t)buf * 128 * 8, vec_s, warp, lane, pre));
                if (last) CH_EDGE(h, si);
              }
            }
          }
        } else {
          for (int h = 0; h < tpc; ++h) {
            CH_ACC_S(3, ch_row_step(prog.r[first], (tpc * pr + h) * 128, prog.M, ew, lane));
            CH_EDGE(h, si);
          }
        }
      }
      first_pair = false;
    }
    if (dbg != nullptr && threadIdx.x == 64) {
      dbg[32] = (unsigned long long)epi_ctr[0];
      dbg[33] = (unsigned long long)epi_ctr[1];
      dbg[34] = (unsigned long long)epi_ctr[2];
      dbg[35] = (unsigned long long)epi_ctr[3];
      dbg[31] = rg_time();
    }
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2u * CH_BN);
  }
}

#ifdef VKN_EPI_PROF
extern "C" int vkn_debug_epi_prof(unsigned long long *out8, int reset) {
  if (out8) cudaMemcpyFromSymbol(out8, g_epi_prof, sizeof(g_epi_prof));
  if (reset) {
    unsigned long long z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_epi_prof, z, sizeof(z));
  }
  return 0;
}
#endif

// ---- host: chain builder -------------------------------------------------------------------------------------------
struct ChainBuild {
  ChainProg prog;
  int np, nr;
};
static ChainBuild *g_chain_pool = nullptr;     // one reusable host record (the program is copied into the launch)

ChainBuild *chain_begin(int M) {
  if (!g_chain_pool) g_chain_pool = (ChainBuild *)aligned_alloc(64, (sizeof(ChainBuild) + 63) / 64 * 64);
  ChainBuild *b = g_chain_pool;
  memset(b, 0, sizeof(*b));
  b->prog.M = M;
  b->prog.mtiles = ceil_div(M, 128);
  return b;
}
bool chain_empty(const ChainBuild *b) { return b->prog.nsteps == 0; }

int chain_add_gemm(ChainBuild *b, const LinArgs *probs, int nprob) {
  if (b->prog.nsteps >= CH_MAXS || b->np + nprob > CH_MAXP) VKN_FAIL(VKN_E_INVALID, "chain: too many steps / problems");
  ChStep &st = b->prog.s[b->prog.nsteps++];
  st.kind = 0;
  st.n = nprob;
  st.first = b->np;
  for (int i = 0; i < nprob; ++i) {
    const LinArgs &a = probs[i];
    if (!linear_tc_supported(a)) VKN_FAIL(VKN_E_INVALID, "chain: GEMM steps need bf16 plane inputs with 16-byte aligned rows");
    if (a.M != b->prog.M) VKN_FAIL(VKN_E_INVALID, "chain: every step must span the same %d rows (got %d)", b->prog.M, a.M);
    if (a.ksplit > 1 || a.side != nullptr) VKN_FAIL(VKN_E_INVALID, "chain: split-K / side outputs are not part of the chain form");
    if ((a.epi & EPI_LN) && a.N > CH_BN) VKN_FAIL(VKN_E_INVALID, "chain: EPI_LN needs N <= %d", CH_BN);
    VKN_TRY(rg_fill_prob(a, CH_BN, 1, b->prog.p[b->np++]));
  }
  return VKN_OK;
}

int chain_add_rowprep(ChainBuild *b, const RowSrc &src, float *out, int ldo, void *planes, int ldp, long long plane_stride, int K) {
  if (b->prog.nsteps >= CH_MAXS || b->nr >= CH_MAXR) VKN_FAIL(VKN_E_INVALID, "chain: too many steps / row operators");
  if (src.pro == PRO_PLANES || K > KC || (K & 1)) VKN_FAIL(VKN_E_INVALID, "chain: row step needs fp32 sources and an even K <= %d", KC);
  ChStep &st = b->prog.s[b->prog.nsteps++];
  st.kind = 1;
  st.n = 1;
  st.first = b->nr;
  ChRow &r = b->prog.r[b->nr++];
  r.src = src;
  r.out = out;
  r.ldo = ldo;
  r.planes = (__nv_bfloat16 *)planes;
  r.ldp = ldp;
  r.plane_stride = plane_stride;
  r.K = K;
  return VKN_OK;
}

int chain_launch(ChainBuild *b, cudaStream_t stream) {
  if (b->prog.nsteps == 0) return VKN_OK;
  ChainProg &g = b->prog;
  constexpr size_t stage_bytes = (size_t)RG_A_BYTES + (size_t)CH_BN * 128;
  g.stages = 2;
  g.stg_bytes = 6144;
  g.pl_off = 0;
  g.idesc = make_idesc_bf16(128, CH_BN, 0, 0);
  g.dbg = debug_ts_slot();
  const size_t smem = (size_t)g.stages * stage_bytes + 8 * (size_t)g.stg_bytes + 1024 + (2 * g.stages + 5) * 8 + 16 + 2 * 128 * 8 * 4 + 64 + 32 + 3 * 256 * 4 + 16;
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) {
    VKN_CUDA_OK(cudaFuncSetAttribute(vkn_chain_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  if (smem > 227 * 1024) VKN_FAIL(VKN_E_INVALID, "chain: shared memory budget exceeded");
  // Pairs of row tiles per CTA (one tile's epilogue and step edge hide behind the other's MMAs) when there are enough
  // tiles to keep every SM busy that way; otherwise one tile per CTA (twice the CTAs, accumulator double-buffered across
  // column tiles).  VKN_CHAIN_TPC overrides.
  g.tpc = g.mtiles > 148 ? 2 : 1;
  if (const char *e = getenv("VKN_CHAIN_TPC")) g.tpc = atoi(e) == 2 ? 2 : 1;
  const int npairs = (g.mtiles + g.tpc - 1) / g.tpc;
  int sms = 148;                                             // VKN_CHAIN_CTAS: leave SMs to the kernels of other streams
  if (const char *e = getenv("VKN_CHAIN_CTAS")) sms = atoi(e) > 0 && atoi(e) <= 148 ? atoi(e) : 148;
  dim3 grid(npairs < sms ? npairs : sms);
  VKN_LAUNCH_MARK("vkn_chain_tc_kernel", stream);
  VKN_CUDA_OK(launch_chain(vkn_chain_tc_kernel, grid, dim3(RG_THREADS), smem, stream, g));
  g.nsteps = 0;
  b->np = 0;
  b->nr = 0;
  return VKN_OK;
}

}  // namespace vkn
