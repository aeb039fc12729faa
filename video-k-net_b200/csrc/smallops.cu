// The N-kernel "rows" operators of the KernelUpdateHead stage: every Linear of KernelUpdator, MHSA,
// FFN and the cls/mask heads runs through ONE fused kernel whose prologue applies the row-wise
// transform feeding it (LayerNorm / ReLU / gate arithmetic, warp-shuffle reductions, staged through
// shared memory) and whose epilogue adds bias / residual / ReLU (or emits the bf16 hi/mid/lo planes the
// tcgen05 mask-conv engine consumes).
//
// Latency structure (these kernels are latency-, not bandwidth-bound): the whole [BN x K-chunk] weight
// tile is fetched with 16-byte cp.async BEFORE the programmatic-dependency wait, so under PDL the weight
// stream of kernel i+1 overlaps the tail of kernel i; the row panel (which depends on kernel i) is built
// afterwards, then one barrier, then the full-K FMA loop out of shared memory.
//
// Reference math: knet/kernel_updator.py:56-94, knet/det/kernel_update_head.py:201-227.
#include <stdlib.h>

#include "common.cuh"
#include "rowops.cuh"

namespace vkn {

constexpr int NT = 256;   // threads per CTA (8 warps: two per scheduler, the kernels are latency-bound)
constexpr int AS_LD = KC + 4;   // panel row stride (floats): bank offset 4 per row -> conflict-free float4 reads

template <typename WT>
struct WTile {
  static constexpr int LD = KC + (sizeof(WT) == 2 ? 8 : 4);   // elements; row stride = 528 B (bf16) / 1040 B (f32)
};

// ---- row operator: one warp per row -----------------------------------------------------------------
// Materialises a prologue ONCE per row (the fused Linear recomputes it in each of its column-block CTAs, which
// is fine for a LayerNorm but not for the KernelUpdator gate: 4 LayerNorms + 2 sigmoids per element).  Output:
// fp32 rows and / or the three bf16 planes a tensor-core Linear consumes with PRO_PLANES.
constexpr int ROW_NT = 128;
__global__ void __launch_bounds__(ROW_NT) vkn_rowop_kernel(const __grid_constant__ RowSrc src, float *out, int ldo,
                                                          __nv_bfloat16 *planes, int ldp, long long plane_stride,
                                                          int M, int K) {
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (ROW_NT / 32) + warp;
  if (row < M) {
    for (int k0 = 0; k0 < K; k0 += KC) {
      const int klen = min(KC, K - k0);
      RowRaw r;
      float v[KPL];
      row_load(src, row, k0, klen, lane, r);
      row_finish(src, klen, lane, r, v);
      if (out != nullptr) store_pairs(out + (size_t)row * ldo + k0, klen, lane, v);
      if (planes != nullptr) {
#pragma unroll
        for (int p = 0; p < KPL / 2; ++p) {
          const int k = kidx(lane, 2 * p);
          float x0 = v[2 * p], x1 = v[2 * p + 1];
#pragma unroll
          for (int t = 0; t < 3; ++t) {          // hi, then the residuals: v == hi + mid + lo to 24 bits
            const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
            x0 -= __bfloat162float(h0);
            x1 -= __bfloat162float(h1);
            if (k < klen)
              *reinterpret_cast<uint32_t *>(planes + (size_t)t * plane_stride + (size_t)row * ldp + k0 + k) =
                  (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          }
        }
      }
    }
  }
  pdl_trigger();
}

// ---- fused rows x Linear ----------------------------------------------------------------------
struct LinBatch {
  LinArgs p[2];
  int nwbuf;      // weight-tile buffers carved from shared memory: 1 (K fits one chunk) or 2
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define VKN_TS(slot)                                                                                   \
  do {                                                                                                 \
    if (A.dbg != nullptr && tid == 0)                                                                  \
      A.dbg[(((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (slot)] = gtime(); \
  } while (0)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// D(16x8, f32) += A(16x16, bf16 row) * B(16x8, bf16 col)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void lin_epilogue(const LinArgs &A, float *outp, const float *bias_s, int col0, int row, int col,
                                             float v) {
  if (row >= A.M || col >= A.N) return;
  if (A.epi & EPI_BIAS) v += ((A.epi & EPI_ROWSCALE) ? __ldg(A.rowscale + row) : 1.f) * bias_s[col - col0];
  if (A.epi & EPI_RES) v += __ldg(A.res + (size_t)row * A.ldres + col);
  if (A.epi & EPI_RELU) v = fmaxf(v, 0.f);
  if (!(A.epi & EPI_NOOUT)) outp[(size_t)row * A.ldo + col] = v;
  if ((A.epi & EPI_SPLIT3) && col < A.split_C) {
    // v == hi + mid + lo to 24 bits; every bf16 x bf16 product in the mask conv is then exact
    const int b = row / A.split_N, n = row - b * A.split_N;
    const size_t plane = (size_t)A.split_B * A.split_Npad * A.split_C;
    const size_t o = ((size_t)b * A.split_Npad + n) * A.split_C + col;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(hi);
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
    A.split_planes[o] = hi;
    A.split_planes[plane + o] = mid;
    A.split_planes[2 * plane + o] = lo;
  }
}

// Shared-memory layout.  fp32 weights: fp32 row panel As[BM][AS_LD] + Ws, FFMA main loop.
// bf16 weights: the fp32 rows are split into three bf16 planes (v = hi + mid + lo to 24 bits) and the
// main loop runs on tensor cores (mma.sync m16n8k16, fp32 accumulate): every product hi/mid/lo x w is
// exact, so the result carries fp32-level accuracy while the FMA work leaves the CUDA cores.
constexpr int PL_LD = KC + 8;   // plane row stride in bf16 (528 B: conflict-free 32-bit fragment loads)

template <typename WT, int BM, int BN>
struct LinSmem {
  static constexpr bool TC = sizeof(WT) == 2;
  static constexpr size_t panel_bytes = TC ? (size_t)3 * BM * PL_LD * 2 : (size_t)BM * AS_LD * 4;
  static constexpr size_t w_one = (size_t)BN * WTile<WT>::LD * sizeof(WT);
  static constexpr size_t w_bytes = 2 * w_one;                         // double-buffered across K chunks (multi-chunk launches)
  static constexpr size_t red_bytes = TC ? (size_t)(NT / 32) * 32 * 4 * 4 : 0;
  static constexpr size_t vec_bytes = (size_t)(9 * KC + BN) * 4;   // LN gamma/beta x4, pre-LN bias, epilogue bias
  static constexpr size_t total = panel_bytes + w_bytes + red_bytes + vec_bytes;
  static constexpr size_t total_single = panel_bytes + w_one + red_bytes + vec_bytes;   // launches whose K fits one chunk
};

template <typename WT, int BM, int BN, int NSRC>
__global__ void __launch_bounds__(NT, BN > 32 ? 2 : (NSRC == 4 ? 2 : 3))
    vkn_linear_kernel(const __grid_constant__ LinBatch batch) {
  using SM = LinSmem<WT, BM, BN>;
  constexpr int WLD = WTile<WT>::LD;
  constexpr int EPV = 16 / sizeof(WT);           // elements per 16-byte cp.async
  extern __shared__ __align__(16) uint8_t lin_smem[];
  float(*As)[AS_LD] = reinterpret_cast<float(*)[AS_LD]>(lin_smem);                 // fp32-weight path
  __nv_bfloat16 *Pl = reinterpret_cast<__nv_bfloat16 *>(lin_smem);                 // bf16 path: [3][BM][PL_LD]
  WT(*Ws0)[WLD] = reinterpret_cast<WT(*)[WLD]>(lin_smem + SM::panel_bytes);
  const size_t wtot = (size_t)batch.nwbuf * SM::w_one;
  float *red = reinterpret_cast<float *>(lin_smem + SM::panel_bytes + wtot);
  float *vecs = reinterpret_cast<float *>(lin_smem + SM::panel_bytes + wtot + SM::red_bytes);   // [9][KC] + [BN]

  const int ks_total = batch.p[0].ksplit;
  const LinArgs &A = batch.p[blockIdx.z / ks_total];
  const int ks = blockIdx.z % ks_total;
  const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
  if (row0 >= A.M || col0 >= A.N) {
    pdl_trigger();
    return;
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  VKN_TS(0);                        // entry

  int kper = (A.K + ks_total - 1) / ks_total;
  kper = (kper + 31) / 32 * 32;
  const int kbeg = ks * kper, kend = min(A.K, kbeg + kper);

  // FFMA path accumulators (fp32 weights)
  constexpr int TM = BM / (NT / 16), TN = BN / 16;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  // tensor-core path: 16x8 output tiles over the 8 warps (K split across warps when there are < 8 tiles)
  constexpr int TILES_N = BN / 8, NTILES = (BM / 16) * TILES_N;
  constexpr int KSW = NTILES >= 8 ? 1 : 8 / NTILES;      // warps sharing one tile along K
  constexpr int TPW = NTILES >= 8 ? NTILES / 8 : 1;      // tiles per warp
  float tacc[TPW][4];          // final accumulators
  float pacc[TPW][3][4];       // one independent MMA chain per bf16 plane (hi / mid / lo)
#pragma unroll
  for (int t = 0; t < TPW; ++t)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      tacc[t][e] = 0.f;
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) pacc[t][pl][e] = 0.f;
    }
  const int kh = KSW > 1 ? warp / NTILES : 0;

  const WT *Wp = reinterpret_cast<const WT *>(A.w);
  const uint32_t ws_base = (uint32_t)__cvta_generic_to_shared(&Ws0[0][0]);

  // ---- parameter vectors (LayerNorm affine, biases) are weights too: stage them in shared memory before the
  //      PDL wait so that no dependent global round trip is left inside the LN / epilogue code.
  const RowSrc &S = A.src;
  {
    const uint32_t v0 = (uint32_t)__cvta_generic_to_shared(vecs);
    const int nln = (S.pro == PRO_GATE) ? 4 : ((S.pro == PRO_LN || S.pro == PRO_LN_RELU) ? 1 : 0);
    const int kq = min(A.K, KC) / 4;                       // 16-byte pieces per vector (LN modes have K <= KC)
    for (int idx = tid; idx < nln * 2 * kq; idx += NT) {
      const int vsel = idx / kq, pc = idx - vsel * kq;     // vsel: 0..nln-1 gamma, nln..2nln-1 beta
      const float *src = (vsel < nln ? S.ln_g[vsel] : S.ln_b[vsel - nln]) + pc * 4;
      const int slot = vsel < nln ? vsel : 4 + (vsel - nln);
      cp_async16(v0 + (uint32_t)(slot * KC + pc * 4) * 4u, src, 16u);
    }
    if (A.epi & EPI_BIAS) {
      for (int pc = tid; pc < BN / 4; pc += NT) {
        const int c = col0 + pc * 4;
        const int valid = c < A.N ? min(4, A.N - c) : 0;
        cp_async16(v0 + (uint32_t)(9 * KC + pc * 4) * 4u, valid ? A.bias + c : A.bias, (uint32_t)valid * 4u);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");      // group "vectors" (older than the weight tile group)
  }
  const float *bias_s = vecs + 9 * KC;

  // weight tile of one K chunk: BN rows x kclen columns, all 16-byte pieces in flight at once (zero-filled
  // beyond N / K).  Weights are never written inside the chain -> safe before the PDL wait.
  auto issue_w = [&](int kc0, int buf) {
    const int kclen = min(KC, kend - kc0);
    const int kpad = (kclen + 31) & ~31;
    const uint32_t ws0 = ws_base + (uint32_t)(buf * SM::w_one);
    if (kpad == KC) {                  // common case: constant trip counts, no integer division
      constexpr int PPR = KC / EPV;    // pieces per row: 32 (bf16) / 64 (f32)
      constexpr int RPP = NT / PPR;    // rows per pass
      const int pc = tid % PPR, nb = tid / PPR;
#pragma unroll
      for (int n0 = 0; n0 < BN; n0 += RPP) {
        const int n = n0 + nb;
        const bool live = col0 + n < A.N;
        const WT *src = live ? Wp + (size_t)(col0 + n) * A.ldw + kc0 + pc * EPV : Wp;
        cp_async16(ws0 + (uint32_t)(n * WLD + pc * EPV) * (uint32_t)sizeof(WT), src, live ? 16u : 0u);
      }
    } else {
      const int pieces_per_row = kpad / EPV;
      const int total = BN * pieces_per_row;
      for (int idx = tid; idx < total; idx += NT) {
        const int n = idx / pieces_per_row, pc = idx - n * pieces_per_row;
        const int gk = kc0 + pc * EPV;
        const bool live = (col0 + n < A.N) && (gk < kend);
        const int valid = live ? min(EPV, kend - gk) : 0;
        const WT *src = live ? Wp + (size_t)(col0 + n) * A.ldw + gk : Wp;
        cp_async16(ws0 + (uint32_t)(n * WLD + pc * EPV) * (uint32_t)sizeof(WT), src, (uint32_t)(valid * sizeof(WT)));
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue_w(kbeg, 0);

  int cbuf = 0;
  for (int kc0 = kbeg; kc0 < kend; kc0 += KC, cbuf ^= 1) {
    const int kclen = min(KC, kend - kc0);
    const int kpad = (kclen + 31) & ~31;
    WT(*Ws)[WLD] = reinterpret_cast<WT(*)[WLD]>(lin_smem + SM::panel_bytes + (size_t)cbuf * SM::w_one);
    const bool more = kc0 + KC < kend;
    const bool planes_in = SM::TC && A.src.pro == PRO_PLANES;
    if (kc0 == kbeg) {
      VKN_TS(1);                     // prefetches issued
      pdl_wait();                    // everything below reads what the previous kernel produced
      VKN_TS(2);                     // dependency resolved
    }
    if (planes_in) {
      // the producer already wrote the three bf16 planes of these rows: plain 16-byte copies, zero-filled tails
      const __nv_bfloat16 *src_pl = reinterpret_cast<const __nv_bfloat16 *>(A.src.a[0]);
      const uint32_t pl0 = (uint32_t)__cvta_generic_to_shared(Pl);
      const int ppr = kpad / 8;                                  // 16-byte pieces per row
      for (int idx = tid; idx < 3 * BM * ppr; idx += NT) {
        const int pc = idx % ppr, r = (idx / ppr) % BM, t = idx / (ppr * BM);
        const int row = row0 + r, gk = kc0 + pc * 8;
        const bool live = row < A.M && gk < kend;
        const int valid = live ? min(8, kend - gk) : 0;
        const __nv_bfloat16 *src = live ? src_pl + (size_t)t * A.src.sum_stride + (size_t)row * A.src.lda[0] + gk : src_pl;
        cp_async16(pl0 + (uint32_t)(((size_t)t * BM + r) * PL_LD + pc * 8) * 2u, src, (uint32_t)valid * 2u);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (more) issue_w(kc0 + KC, cbuf ^ 1);          // next chunk's weights stream while this chunk is built / multiplied
    if (kc0 == kbeg && !planes_in) {
      // groups in flight: vectors, W(0) [, W(1)]: the vector group must have landed before the LN code runs
      if (more) asm volatile("cp.async.wait_group 2;" ::: "memory");
      else asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncthreads();
    }
    // ---- panel: transformed rows row0..row0+BM, columns kc0..kc0+kclen (zero padded to kpad).
    //      Each warp owns RPW rows; the raw loads of GRP rows are issued before any reduction.
    if (!planes_in) {
      constexpr int NW = NT / 32, RPW = BM / NW, GRP = 2;
      static_assert(BM % NW == 0 && RPW % GRP == 0, "rows per warp must be a multiple of the load group");
#pragma unroll 1
      for (int q0 = 0; q0 < RPW; q0 += GRP) {
        RowRawT<NSRC> raw[GRP];
#pragma unroll
        for (int q = 0; q < GRP; ++q) {
          const int row = row0 + warp + NW * (q0 + q);
          if (row < A.M) row_load(S, row, kc0, kclen, lane, raw[q]);
        }
#pragma unroll
        for (int q = 0; q < GRP; ++q) {
          const int r = warp + NW * (q0 + q), row = row0 + r;
          float v[KPL];
          if (row < A.M) {
            row_finish(S, kclen, lane, raw[q], v, vecs);
          } else {
#pragma unroll
            for (int i = 0; i < KPL; ++i) v[i] = 0.f;
          }
          if (SM::TC) {
#pragma unroll
            for (int p = 0; p < KPL / 2; ++p) {
              const int k = kidx(lane, 2 * p);
              uint32_t w3[3];
#pragma unroll
              for (int t = 0; t < 3; ++t) w3[t] = 0u;
              float x0 = v[2 * p], x1 = v[2 * p + 1];
#pragma unroll
              for (int t = 0; t < 3; ++t) {          // hi, then the residuals
                const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                x0 -= __bfloat162float(h0);
                x1 -= __bfloat162float(h1);
                w3[t] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
              }
              if (k < kpad) {
#pragma unroll
                for (int t = 0; t < 3; ++t)
                  *reinterpret_cast<uint32_t *>(Pl + ((size_t)t * BM + r) * PL_LD + k) = w3[t];
              }
            }
          } else {
#pragma unroll
            for (int p = 0; p < KPL / 2; ++p) {
              const int k = kidx(lane, 2 * p);
              *reinterpret_cast<float2 *>(&As[r][k]) = make_float2(v[2 * p], v[2 * p + 1]);
            }
          }
          if (A.side != nullptr && blockIdx.x == 0 && row < A.M)
            store_pairs(A.side + (size_t)row * A.ldside + kc0, kclen, lane, v);
        }
      }
    }
    VKN_TS(3);                       // panel built
    if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");    // this chunk's tile landed; the next may be in flight
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    VKN_TS(4);                       // tile + panel visible
    if (!more) pdl_trigger();   // last chunk staged: let the next kernel start its prefetch
    if (SM::TC) {
      const int g = lane >> 2, t4 = lane & 3;
      const int nsteps = kpad / 16;
      const int s_beg = kh * nsteps / KSW, s_end = (kh + 1) * nsteps / KSW;
#pragma unroll
      for (int tt = 0; tt < TPW; ++tt) {
        const int tile = KSW > 1 ? warp % NTILES : warp * TPW + tt;
        const int tm = tile / TILES_N, tn = tile - tm * TILES_N;
        const __nv_bfloat16 *wrow = reinterpret_cast<const __nv_bfloat16 *>(&Ws[tn * 8 + g][0]) + 2 * t4;
        const __nv_bfloat16 *prow = Pl + (size_t)(tm * 16 + g) * PL_LD + 2 * t4;
#pragma unroll 2
        for (int st = s_beg; st < s_end; ++st) {
          const int k0 = st * 16;
          const uint32_t b0 = *reinterpret_cast<const uint32_t *>(wrow + k0);
          const uint32_t b1 = *reinterpret_cast<const uint32_t *>(wrow + k0 + 8);
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            const __nv_bfloat16 *pp = prow + (size_t)t * BM * PL_LD + k0;
            uint32_t a[4];
            a[0] = *reinterpret_cast<const uint32_t *>(pp);
            a[1] = *reinterpret_cast<const uint32_t *>(pp + 8 * PL_LD);
            a[2] = *reinterpret_cast<const uint32_t *>(pp + 8);
            a[3] = *reinterpret_cast<const uint32_t *>(pp + 8 * PL_LD + 8);
            mma_bf16_16816(pacc[tt][t], a, b0, b1);
          }
        }
      }
    } else {
      const int tx = tid & 15, ty = tid >> 4;
#pragma unroll 2
      for (int kk = 0; kk < kpad; kk += 8) {
        float a[TM][8], b[TN][8];
#pragma unroll
        for (int i = 0; i < TM; ++i) load8(&As[ty + (NT / 16) * i][kk], a[i]);
#pragma unroll
        for (int j = 0; j < TN; ++j) load8(reinterpret_cast<const WT *>(&Ws[tx + 16 * j][kk]), b[j]);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[i][j] = fmaf(a[i][e], b[j][e], acc[i][j]);
      }
    }
    if (more) __syncthreads();
  }

  VKN_TS(5);                         // main loop done
  float *outp = A.out + (size_t)ks * A.out_split_stride;
  if (SM::TC) {
#pragma unroll
    for (int t = 0; t < TPW; ++t)
#pragma unroll
      for (int e = 0; e < 4; ++e) tacc[t][e] = (pacc[t][2][e] + pacc[t][1][e]) + pacc[t][0][e];   // small terms first
    if (KSW > 1) {     // fold the K-halves: warps kh > 0 hand their fragments to the kh == 0 warp of the tile
      if (kh > 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) red[(warp * 32 + lane) * 4 + e] = tacc[0][e];
      }
      __syncthreads();
      if (kh == 0) {
#pragma unroll
        for (int o = 1; o < KSW; ++o)
#pragma unroll
          for (int e = 0; e < 4; ++e) tacc[0][e] += red[((warp + o * NTILES) * 32 + lane) * 4 + e];
      }
    }
    if (kh == 0) {
      const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
      for (int tt = 0; tt < TPW; ++tt) {
        const int tile = KSW > 1 ? warp % NTILES : warp * TPW + tt;
        const int tm = tile / TILES_N, tn = tile - tm * TILES_N;
        const int r = row0 + tm * 16 + g, c = col0 + tn * 8 + 2 * t4;
        lin_epilogue(A, outp, bias_s, col0, r, c, tacc[tt][0]);
        lin_epilogue(A, outp, bias_s, col0, r, c + 1, tacc[tt][1]);
        lin_epilogue(A, outp, bias_s, col0, r + 8, c, tacc[tt][2]);
        lin_epilogue(A, outp, bias_s, col0, r + 8, c + 1, tacc[tt][3]);
      }
    }
  } else {
    const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) lin_epilogue(A, outp, bias_s, col0, row0 + ty + (NT / 16) * i, col0 + tx + 16 * j, acc[i][j]);
  }
  VKN_TS(6);                         // stores issued
}

static int check_src(const RowSrc &s, int K) {
  if (s.pro != PRO_COPY && s.pro != PRO_PLANES && K > KC)
    VKN_FAIL(VKN_E_UNSUPPORTED, "row transform %d needs K <= %d (got %d)", s.pro, KC, K);
  if (s.nsum < 1) VKN_FAIL(VKN_E_INVALID, "RowSrc.nsum must be >= 1");
  return VKN_OK;
}

template <typename WT, int BM, int BN, int NSRC>
static int launch_linear_n(const LinBatch &b, dim3 grid, cudaStream_t stream) {
  const size_t smem = b.nwbuf == 2 ? LinSmem<WT, BM, BN>::total : LinSmem<WT, BM, BN>::total_single;
  if (smem > 227 * 1024) VKN_FAIL(VKN_E_UNSUPPORTED, "linear tile %dx%d needs %zu bytes of shared memory", BM, BN, smem);
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) {
    VKN_CUDA_OK(cudaFuncSetAttribute(vkn_linear_kernel<WT, BM, BN, NSRC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(LinSmem<WT, BM, BN>::total < 227 * 1024 ? LinSmem<WT, BM, BN>::total : 227 * 1024)));
  }
  VKN_CUDA_OK(launch_chain(vkn_linear_kernel<WT, BM, BN, NSRC>, grid, dim3(NT), smem, stream, b));
  return VKN_OK;
}

// NSRC = 4 only for the KernelUpdator gate prologue (four LayerNorm'd sources); every other prologue needs at
// most two row sources, which keeps the register footprint at 3 CTAs per SM.
template <typename WT, int BM, int BN>
static int launch_linear_t(const LinBatch &b, dim3 grid, cudaStream_t stream) {
  const bool gate = b.p[0].src.pro == PRO_GATE || b.p[1].src.pro == PRO_GATE;
  return gate ? launch_linear_n<WT, BM, BN, 4>(b, grid, stream) : launch_linear_n<WT, BM, BN, 2>(b, grid, stream);
}

int launch_linear(const LinArgs *probs, int nprob, int w_dtype, cudaStream_t stream) {
  if (nprob < 1 || nprob > 2) VKN_FAIL(VKN_E_INVALID, "launch_linear: nprob %d", nprob);
  LinBatch b;
  int maxM = 0, maxN = 0;
  for (int i = 0; i < nprob; ++i) {
    b.p[i] = probs[i];
    if (b.p[i].ksplit < 1) b.p[i].ksplit = 1;
    if (b.p[i].ksplit != b.p[0].ksplit) VKN_FAIL(VKN_E_INVALID, "launch_linear: batched problems must share ksplit");
    VKN_TRY(check_src(b.p[i].src, b.p[i].K));
    if (b.p[i].src.pro == PRO_PLANES && w_dtype != VKN_BF16)
      VKN_FAIL(VKN_E_INVALID, "launch_linear: plane inputs need the tensor-core (bf16 weight) path");
    if (b.p[i].ldw % 8 != 0 || (reinterpret_cast<uintptr_t>(b.p[i].w) & 15))
      VKN_FAIL(VKN_E_INVALID, "launch_linear: weight rows must be 16-byte aligned (ldw %d)", b.p[i].ldw);
    maxM = max(maxM, b.p[i].M);
    maxN = max(maxN, b.p[i].N);
  }
  if (nprob == 1) b.p[1] = b.p[0];
  {
    int kmax = 0;
    for (int i = 0; i < nprob; ++i) {
      int kper = ceil_div(b.p[i].K, b.p[i].ksplit);
      kper = (kper + 31) / 32 * 32;
      kmax = max(kmax, kper);
    }
    b.nwbuf = kmax > KC ? 2 : 1;
  }
  {
    unsigned long long *ts = debug_ts_slot();
    b.p[0].dbg = ts;
    b.p[1].dbg = ts;
  }
  const int ks = b.p[0].ksplit;
  // Tile choice (measured on B200, tools/kernel_times.py + VKN_LINEAR_BN sweep): always 16-row tiles (32/64-row
  // tiles lose to more waves of 16-row CTAs: serial row groups per warp); 64 columns is the best or within 2 % of
  // the best column width from 100 to 800 rows (32 columns recompute the row prologue in 8 CTAs instead of 4;
  // 128 columns leave too few CTAs for a single frame).
  int bn = maxN >= 64 ? 64 : 32;
  if (const char *e = getenv("VKN_LINEAR_BN")) {
    const int v = atoi(e);
    if (v == 32 || v == 64 || (v == 128 && w_dtype == VKN_BF16)) bn = v;
  }
  static const char *names[3] = {"vkn_linear_kernel<16x32>", "vkn_linear_kernel<16x64>", "vkn_linear_kernel<16x128>"};
  VKN_LAUNCH_MARK(names[bn == 32 ? 0 : (bn == 64 ? 1 : 2)], stream);
  dim3 grid(ceil_div(maxN, bn), ceil_div(maxM, 16), nprob * ks);
#define VKN_LIN_DISPATCH(BN_)                                                                   \
  return w_dtype == VKN_BF16 ? launch_linear_t<__nv_bfloat16, 16, BN_>(b, grid, stream)          \
                             : launch_linear_t<float, 16, BN_>(b, grid, stream)
  if (bn == 32) VKN_LIN_DISPATCH(32);
  if (bn == 64) VKN_LIN_DISPATCH(64);
  VKN_LIN_DISPATCH(128);
#undef VKN_LIN_DISPATCH
}

int launch_rowop(const RowSrc &src, float *out, int ldo, int M, int K, cudaStream_t stream) {
  VKN_TRY(check_src(src, K));
  if (src.pro == PRO_PLANES) VKN_FAIL(VKN_E_INVALID, "launch_rowop: plane inputs are only consumed by launch_linear");
  return launch_rowprep(src, out, ldo, nullptr, 0, 0, M, K, stream);
}

int launch_rowprep(const RowSrc &src, float *out, int ldo, void *planes, int ldp, long long plane_stride, int M, int K,
                   cudaStream_t stream) {
  VKN_TRY(check_src(src, K));
  if (src.pro == PRO_PLANES) VKN_FAIL(VKN_E_INVALID, "launch_rowprep: plane inputs are only consumed by launch_linear");
  VKN_LAUNCH_MARK("vkn_rowop_kernel", stream);
  VKN_CUDA_OK(launch_chain(vkn_rowop_kernel, dim3(ceil_div(M, ROW_NT / 32)), dim3(ROW_NT), 0, stream, src, out, ldo,
                           (__nv_bfloat16 *)planes, ldp, plane_stride, M, K));
  return VKN_OK;
}

// ---- attention among the N kernels of a frame ---------------------------------------------------
// One CTA = (query block, head, frame).  K and V of the head sit in shared memory ([N][hd+1], bank
// conflict free); one warp per query: lane-parallel scores, shuffle softmax, lane-per-channel PV.
// torch semantics (F.multi_head_attention_forward): q scaled by 1/sqrt(hd) BEFORE q.k^T.
constexpr int ATT_QB = 8;

__global__ void __launch_bounds__(NT) vkn_attention_kernel(const float *__restrict__ q, int ldq,
                                                           const float *__restrict__ k, int ldk,
                                                           const float *__restrict__ v, int ldv,
                                                           float *__restrict__ out, int ldo, int N, int hd,
                                                           float scale, __nv_bfloat16 *__restrict__ planes,
                                                           long long plane_stride, int qb) {
  extern __shared__ float smem[];
  const int hs = hd + 1;
  float *Ks = smem;                    // [N][hs]
  float *Vs = Ks + (size_t)N * hs;     // [N][hs]
  float *Ps = Vs + (size_t)N * hs;             // [NT/32 warps][N]
  float *Qs = Ps + (NT / 32) * (size_t)N;      // [NT/32 warps][32]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * qb;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t rowb = (size_t)b * N;
  pdl_wait();
  if ((hd & 3) == 0 && (ldk & 3) == 0 && (ldv & 3) == 0) {
    const int hd4 = hd >> 2;
    for (int idx = tid; idx < N * hd4; idx += NT) {
      const int j = idx / hd4, d = (idx - j * hd4) * 4;
      const float4 kk = __ldg(reinterpret_cast<const float4 *>(k + (rowb + j) * ldk + h * hd + d));
      const float4 vv = __ldg(reinterpret_cast<const float4 *>(v + (rowb + j) * ldv + h * hd + d));
      float *kd = Ks + j * hs + d, *vd = Vs + j * hs + d;
      kd[0] = kk.x; kd[1] = kk.y; kd[2] = kk.z; kd[3] = kk.w;
      vd[0] = vv.x; vd[1] = vv.y; vd[2] = vv.z; vd[3] = vv.w;
    }
  } else {
    for (int idx = tid; idx < N * hd; idx += NT) {
      const int j = idx / hd, d = idx - j * hd;
      Ks[j * hs + d] = __ldg(k + (rowb + j) * ldk + h * hd + d);
      Vs[j * hs + d] = __ldg(v + (rowb + j) * ldv + h * hd + d);
    }
  }
  __syncthreads();
  pdl_trigger();
  float *ps = Ps + (size_t)warp * N;
  float *qs = Qs + warp * 32;
  for (int qi = q0 + warp; qi < min(N, q0 + qb); qi += NT / 32) {
    if (lane < hd) qs[lane] = __ldg(q + (rowb + qi) * ldq + h * hd + lane) * scale;
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) {
      float s0 = 0.f, s1 = 0.f;
      int d = 0;
      for (; d + 2 <= hd; d += 2) {
        s0 = fmaf(qs[d], Ks[j * hs + d], s0);
        s1 = fmaf(qs[d + 1], Ks[j * hs + d + 1], s1);
      }
      if (d < hd) s0 = fmaf(qs[d], Ks[j * hs + d], s0);
      const float s = s0 + s1;
      ps[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float e = expf(ps[j] - mx);
      ps[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    if (lane < hd) {
      float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
      int j = 0;
      for (; j + 4 <= N; j += 4) {
        o0 = fmaf(ps[j], Vs[j * hs + lane], o0);
        o1 = fmaf(ps[j + 1], Vs[(j + 1) * hs + lane], o1);
        o2 = fmaf(ps[j + 2], Vs[(j + 2) * hs + lane], o2);
        o3 = fmaf(ps[j + 3], Vs[(j + 3) * hs + lane], o3);
      }
      for (; j < N; ++j) o0 = fmaf(ps[j], Vs[j * hs + lane], o0);
      const float ov = ((o0 + o1) + (o2 + o3)) / sum;
      if (out != nullptr) out[(rowb + qi) * ldo + h * hd + lane] = ov;
      if (planes != nullptr) {     // A operand of the out-projection on the tcgen05 row engine
        float xr = ov;
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
          const __nv_bfloat16 hb = __float2bfloat16_rn(xr);
          xr -= __bfloat162float(hb);
          planes[(size_t)pl * plane_stride + (rowb + qi) * ldo + h * hd + lane] = hb;
        }
      }
    }
    __syncwarp();
  }
}

// Batched form (many (head, frame) pairs in flight): a warp carries FOUR queries through each pass, so every K / V
// element fetched from shared memory feeds four FMAs and the four queries' q / p values arrive as one broadcast
// 16-byte load ([d][4] and [key][4] layouts).  The single-query kernel above is bound by its shared-memory loads
// (~330 per query); this one needs ~90.  Same arithmetic per (query, key, channel), same softmax order.
constexpr int ATT4_QB = 32;    // queries per CTA: 8 warps x 4

__global__ void __launch_bounds__(NT) vkn_attention4_kernel(const float *__restrict__ q, int ldq,
                                                            const float *__restrict__ k, int ldk,
                                                            const float *__restrict__ v, int ldv,
                                                            float *__restrict__ out, int ldo, int N, int hd,
                                                            float scale, __nv_bfloat16 *__restrict__ planes,
                                                            long long plane_stride) {
  extern __shared__ float smem[];
  const int hs = hd + 1;
  float *Ks = smem;                                  // [N][hs]
  float *Vs = Ks + (size_t)N * hs;                   // [N][hs]
  float *Ps = Vs + (size_t)N * hs;                   // [8 warps][N][4]
  Ps = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(Ps) + 15) & ~(uintptr_t)15);
  float *Qs = Ps + (NT / 32) * (size_t)N * 4;        // [8 warps][32][4]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT4_QB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t rowb = (size_t)b * N;
  pdl_wait();
  if ((hd & 3) == 0 && (ldk & 3) == 0 && (ldv & 3) == 0) {
    const int hd4 = hd >> 2;
    for (int idx = tid; idx < N * hd4; idx += NT) {
      const int j = idx / hd4, d = (idx - j * hd4) * 4;
      const float4 kk = __ldg(reinterpret_cast<const float4 *>(k + (rowb + j) * ldk + h * hd + d));
      const float4 vv = __ldg(reinterpret_cast<const float4 *>(v + (rowb + j) * ldv + h * hd + d));
      float *kd = Ks + j * hs + d, *vd = Vs + j * hs + d;
      kd[0] = kk.x; kd[1] = kk.y; kd[2] = kk.z; kd[3] = kk.w;
      vd[0] = vv.x; vd[1] = vv.y; vd[2] = vv.z; vd[3] = vv.w;
    }
  } else {
    for (int idx = tid; idx < N * hd; idx += NT) {
      const int j = idx / hd, d = idx - j * hd;
      Ks[j * hs + d] = __ldg(k + (rowb + j) * ldk + h * hd + d);
      Vs[j * hs + d] = __ldg(v + (rowb + j) * ldv + h * hd + d);
    }
  }
  __syncthreads();
  pdl_trigger();
  float4 *ps4 = reinterpret_cast<float4 *>(Ps) + (size_t)warp * N;
  float4 *qs4 = reinterpret_cast<float4 *>(Qs) + warp * 32;
  const int qbase = q0 + warp * 4;
  if (qbase >= N) return;
  const int nq = min(4, N - qbase);
  if (lane < hd) {
    float qv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) qv[i] = i < nq ? __ldg(q + (rowb + qbase + i) * ldq + h * hd + lane) * scale : 0.f;
    qs4[lane] = make_float4(qv[0], qv[1], qv[2], qv[3]);
  }
  __syncwarp();
  // pass 1: raw scores of 4 queries x (4 keys per lane per sweep); running max per query
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  for (int j0 = 0; j0 < N; j0 += 128) {
    float sc[4][4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i) sc[t][i] = 0.f;
    int jr[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) jr[t] = min(j0 + lane + 32 * t, N - 1) * hs;
    for (int d = 0; d < hd; ++d) {
      const float4 q4 = qs4[d];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float kv = Ks[jr[t] + d];
        sc[t][0] = fmaf(q4.x, kv, sc[t][0]);
        sc[t][1] = fmaf(q4.y, kv, sc[t][1]);
        sc[t][2] = fmaf(q4.z, kv, sc[t][2]);
        sc[t][3] = fmaf(q4.w, kv, sc[t][3]);
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = j0 + lane + 32 * t;
      if (j < N) {
        ps4[j] = make_float4(sc[t][0], sc[t][1], sc[t][2], sc[t][3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) mx[i] = fmaxf(mx[i], sc[t][i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) mx[i] = warp_max(mx[i]);
  // pass 2: exponentials (each lane revisits the keys it wrote) and their sums
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = lane; j < N; j += 32) {
    float4 e = ps4[j];
    e.x = expf(e.x - mx[0]); e.y = expf(e.y - mx[1]); e.z = expf(e.z - mx[2]); e.w = expf(e.w - mx[3]);
    ps4[j] = e;
    sum[0] += e.x; sum[1] += e.y; sum[2] += e.z; sum[3] += e.w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) sum[i] = warp_sum(sum[i]);
  __syncwarp();
  // pass 3: lane = channel; every V element feeds the four queries
  if (lane < hd) {
    float o[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i][0] = o[i][1] = 0.f;
    int j = 0;
    for (; j + 2 <= N; j += 2) {
      const float4 p0 = ps4[j], p1 = ps4[j + 1];
      const float v0 = Vs[j * hs + lane], v1 = Vs[(j + 1) * hs + lane];
      o[0][0] = fmaf(p0.x, v0, o[0][0]); o[1][0] = fmaf(p0.y, v0, o[1][0]);
      o[2][0] = fmaf(p0.z, v0, o[2][0]); o[3][0] = fmaf(p0.w, v0, o[3][0]);
      o[0][1] = fmaf(p1.x, v1, o[0][1]); o[1][1] = fmaf(p1.y, v1, o[1][1]);
      o[2][1] = fmaf(p1.z, v1, o[2][1]); o[3][1] = fmaf(p1.w, v1, o[3][1]);
    }
    if (j < N) {
      const float4 p0 = ps4[j];
      const float v0 = Vs[j * hs + lane];
      o[0][0] = fmaf(p0.x, v0, o[0][0]); o[1][0] = fmaf(p0.y, v0, o[1][0]);
      o[2][0] = fmaf(p0.z, v0, o[2][0]); o[3][0] = fmaf(p0.w, v0, o[3][0]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < nq) {
        const float ov = (o[i][0] + o[i][1]) / sum[i];
        const size_t off = (rowb + qbase + i) * ldo + h * hd + lane;
        if (out != nullptr) out[off] = ov;
        if (planes != nullptr) {
          float xr = ov;
#pragma unroll
          for (int pl = 0; pl < 3; ++pl) {
            const __nv_bfloat16 hb = __float2bfloat16_rn(xr);
            xr -= __bfloat162float(hb);
            planes[(size_t)pl * plane_stride + off] = hb;
          }
        }
      }
    }
  }
}

int launch_attention(const float *q, int ldq, const float *k, int ldk, const float *v, int ldv, float *out,
                     int ldo, int B, int N, int C, int heads, cudaStream_t stream, void *planes, long long plane_stride) {
  if (heads < 1 || C % heads != 0) VKN_FAIL(VKN_E_INVALID, "attention: C %d not divisible by heads %d", C, heads);
  const int hd = C / heads;
  if (hd > 32) VKN_FAIL(VKN_E_UNSUPPORTED, "attention: head_dim %d > 32", hd);
  {
    // frame batches: the tcgen05 kernel (one CTA per frame, both products on tensor cores); a few frames: the SIMT kernels
    // below spread (query block, head, frame) over many CTAs.  VKN_ATT_TC=0 forces SIMT, VKN_ATT_TC_MIN sets the crossover.
    int tc_min = 8;
    if (const char *e = getenv("VKN_ATT_TC_MIN")) tc_min = atoi(e);
    const char *e = getenv("VKN_ATT_TC");
    if (!(e && e[0] == '0') && B >= tc_min && attention_tc_supported(N, C, heads, q, ldq, k, ldk, v, ldv, out, ldo, planes, plane_stride))
      return launch_attention_tc(q, ldq, k, ldk, v, ldv, out, ldo, B, N, C, heads, stream, planes, plane_stride);
  }
  const size_t smem = ((size_t)2 * N * (hd + 1) + (NT / 32) * (size_t)N + (NT / 32) * 32) * sizeof(float);
  static unsigned long long attr_set = 0;
  if (first_use_on_device(attr_set)) {
    VKN_CUDA_OK(cudaFuncSetAttribute(vkn_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  if (smem > 160 * 1024) VKN_FAIL(VKN_E_UNSUPPORTED, "attention: N %d too large for shared memory", N);
  // queries per CTA: one per warp for a single frame (latency), more when many (head, frame) pairs are in flight --
  // every CTA stages the K and V of its head, so fewer, longer CTAs stop re-reading them (one wave of <= ~296 CTAs)
  int qb = ATT_QB;
  (void)B;   // (measured: longer CTAs lose -- the kernel is bound by its per-query instruction stream, not by K/V staging)
  if (const char *e = getenv("VKN_ATT_QB")) qb = atoi(e) >= ATT_QB ? atoi(e) / ATT_QB * ATT_QB : qb;
  bool four = (long long)ceil_div(N, ATT4_QB) * heads * B >= 120;     // enough CTAs for every SM: the batched form
  if (const char *e = getenv("VKN_ATT4")) four = e[0] == '1';
  if (four) {
    const size_t smem4 = ((size_t)2 * N * (hd + 1) + 4 + (NT / 32) * (size_t)N * 4 + (NT / 32) * 32 * 4) * sizeof(float);
    static unsigned long long attr4 = 0;
    if (first_use_on_device(attr4)) {
      VKN_CUDA_OK(cudaFuncSetAttribute(vkn_attention4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    if (smem4 > 200 * 1024) VKN_FAIL(VKN_E_UNSUPPORTED, "attention: N %d too large for shared memory", N);
    VKN_LAUNCH_MARK("vkn_attention4_kernel", stream);
    VKN_CUDA_OK(launch_chain(vkn_attention4_kernel, dim3(ceil_div(N, ATT4_QB), heads, B), dim3(NT), smem4, stream, q, ldq, k,
                             ldk, v, ldv, out, ldo, N, hd, 1.0f / sqrtf((float)hd), (__nv_bfloat16 *)planes, plane_stride));
    return VKN_OK;
  }
  dim3 grid(ceil_div(N, qb), heads, B);
  VKN_LAUNCH_MARK("vkn_attention_kernel", stream);
  VKN_CUDA_OK(launch_chain(vkn_attention_kernel, grid, dim3(NT), smem, stream, q, ldq, k, ldk, v, ldv, out, ldo, N, hd,
                           1.0f / sqrtf((float)hd), (__nv_bfloat16 *)planes, plane_stride, qb));
  return VKN_OK;
}

}  // namespace vkn
