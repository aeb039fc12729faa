// Row-wise transforms shared by the fused row kernels (smallops.cu) and the chain kernel (rowgemm_tc.cu): a warp owns a
// row, lane l holds element pairs; LayerNorm / gate arithmetic by warp shuffles.  Template flag CG = read the sources with
// ld.global.cg (L2-coherent): required when the rows were written earlier by the SAME kernel (the chain kernel), where
// the non-coherent path (__ldg) could return stale lines.
#pragma once
#include "common.cuh"

namespace vkn {

constexpr int KC = 256;       // K-chunk a warp holds in registers (row panel width)
constexpr int KPL = KC / 32;  // panel elements per lane

// ---- row-wise prologue ------------------------------------------------------------------------
// A warp owns a row; lane l holds the element PAIRS k = 64 p + 2 l, +1 (p = 0..3): 8-byte loads/stores,
// and the pair is what one 32-bit bf16x2 word of the tensor-core operand planes holds.
__device__ __forceinline__ int kidx(int lane, int i) { return ((i >> 1) << 6) + (lane << 1) + (i & 1); }

__device__ __forceinline__ void ln_inplace(float (&v)[KPL], int K, int lane, const float *g, const float *b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < KPL; ++i) s += v[i];  // out-of-range slots hold 0
  const float mean = warp_sum(s) / (float)K;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
    const float d = (kidx(lane, i) < K) ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)K + 1e-5f);
#pragma unroll
  for (int p = 0; p < KPL / 2; ++p) {
    const int k = kidx(lane, 2 * p);
    if (k < K) {
      const float2 gg = *reinterpret_cast<const float2 *>(g + k);     // global or shared (generic load)
      const float2 bb = *reinterpret_cast<const float2 *>(b + k);
      v[2 * p] = (v[2 * p] - mean) * rstd * gg.x + bb.x;
      v[2 * p + 1] = (v[2 * p + 1] - mean) * rstd * gg.y + bb.y;
    }
  }
}

// Four LayerNorms of the same row at once (the KernelUpdator gate prologue): the four pairs of warp
// reductions are interleaved so their shuffle latencies overlap instead of adding up.
__device__ __forceinline__ void ln4_inplace(float (&v)[4][KPL], int K, int lane, const float *const (&g)[4],
                                            const float *const (&b)[4]) {
  float s[4], q[4], mean[4], rstd[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    s[t] = 0.f;
#pragma unroll
    for (int i = 0; i < KPL; ++i) s[t] += v[t][i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int t = 0; t < 4; ++t) s[t] += __shfl_xor_sync(0xffffffffu, s[t], o);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    mean[t] = s[t] / (float)K;
    q[t] = 0.f;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const float d = (kidx(lane, i) < K) ? v[t][i] - mean[t] : 0.f;
      q[t] += d * d;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int t = 0; t < 4; ++t) q[t] += __shfl_xor_sync(0xffffffffu, q[t], o);
#pragma unroll
  for (int t = 0; t < 4; ++t) rstd[t] = 1.0f / sqrtf(q[t] / (float)K + 1e-5f);
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int p = 0; p < KPL / 2; ++p) {
      const int k = kidx(lane, 2 * p);
      if (k < K) {
        const float2 gg = *reinterpret_cast<const float2 *>(g[t] + k);
        const float2 bb = *reinterpret_cast<const float2 *>(b[t] + k);
        v[t][2 * p] = (v[t][2 * p] - mean[t]) * rstd[t] * gg.x + bb.x;
        v[t][2 * p + 1] = (v[t][2 * p + 1] - mean[t]) * rstd[t] * gg.y + bb.y;
      }
    }
}

// v[2p], v[2p+1] (+)= a[k], a[k+1]   (k even, klen even: a pair never straddles the end)
template <bool ACC, bool CG = false>
__device__ __forceinline__ void fetch_pairs(const float *a, int klen, int lane, float (&v)[KPL]) {
#pragma unroll
  for (int p = 0; p < KPL / 2; ++p) {
    const int k = kidx(lane, 2 * p);
    float2 t = make_float2(0.f, 0.f);
    if (k < klen) t = CG ? __ldcg(reinterpret_cast<const float2 *>(a + k)) : __ldg(reinterpret_cast<const float2 *>(a + k));
    if (ACC) {
      v[2 * p] += t.x;
      v[2 * p + 1] += t.y;
    } else {
      v[2 * p] = t.x;
      v[2 * p + 1] = t.y;
    }
  }
}

// Row transform in two steps so that the global loads of ALL rows a warp owns are in flight before the
// first LayerNorm reduction starts (the kernels are latency-bound).
//   row_load  : raw values of up to 4 sources -> registers (sum of slices + bias + residual folded in)
//   row_finish: LN / ReLU / gate arithmetic (warp-shuffle reductions)
// LN-type modes require k0 == 0 and klen == K (host-checked).
template <int NSRC>
struct RowRawT {
  float v[NSRC][KPL];
};
using RowRaw = RowRawT<4>;

template <int NSRC, bool CG = false>
__device__ __forceinline__ void row_load(const RowSrc &s, int row, int k0, int klen, int lane, RowRawT<NSRC> &r) {
  if (s.pro == PRO_MUL) {
    fetch_pairs<false, CG>(s.a[0] + (size_t)row * s.lda[0] + k0, klen, lane, r.v[0]);
    fetch_pairs<false, CG>(s.a[1] + (size_t)row * s.lda[1] + k0, klen, lane, r.v[1]);
    return;
  }
  if constexpr (NSRC == 4) {
    if (s.pro == PRO_GATE) {
      fetch_pairs<false, CG>(s.a[0] + (size_t)row * s.lda[0], klen, lane, r.v[0]);   // update gate pre-activation
      fetch_pairs<false, CG>(s.a[1] + (size_t)row * s.lda[1], klen, lane, r.v[1]);   // param_out
      fetch_pairs<false, CG>(s.a[2] + (size_t)row * s.lda[2], klen, lane, r.v[2]);   // input gate pre-activation
      fetch_pairs<false, CG>(s.a[3] + (size_t)row * s.lda[3], klen, lane, r.v[3]);   // input_out
      return;
    }
  }
  // PRO_COPY / PRO_LN / PRO_LN_RELU: fixed-order sum of slices (+ bias + residual)
  float(&v)[KPL] = r.v[0];
  const float *a0 = s.a[0] + (size_t)row * s.lda[0] + k0;
  fetch_pairs<false, CG>(a0, klen, lane, v);
  int sl = 1;
  for (; sl + 4 <= s.nsum; sl += 4) {      // four independent slices in flight, fixed association order
    float t[4][KPL];
#pragma unroll
    for (int u = 0; u < 4; ++u) fetch_pairs<false, CG>(a0 + (size_t)(sl + u) * s.sum_stride, klen, lane, t[u]);
#pragma unroll
    for (int i = 0; i < KPL; ++i) v[i] += (t[0][i] + t[1][i]) + (t[2][i] + t[3][i]);
  }
  for (; sl < s.nsum; ++sl) fetch_pairs<true, CG>(a0 + (size_t)sl * s.sum_stride, klen, lane, v);
  if (s.pbias) fetch_pairs<true, CG>(s.pbias + k0, klen, lane, v);
  if (s.pres) fetch_pairs<true, CG>(s.pres + (size_t)(s.pres_mod > 0 ? row % s.pres_mod : row) * s.ldpres + k0, klen, lane, v);
}

// lnv: optional shared-memory copy of the LayerNorm vectors ([i] gamma at lnv + i*KC, beta at lnv + (4+i)*KC)
template <int NSRC>
__device__ __forceinline__ void row_finish(const RowSrc &s, int klen, int lane, RowRawT<NSRC> &r, float (&v)[KPL],
                                           const float *lnv = nullptr) {
  if (s.pro == PRO_MUL) {
#pragma unroll
    for (int i = 0; i < KPL; ++i) v[i] = r.v[0][i] * r.v[1][i];
    return;
  }
  if constexpr (NSRC == 4) {
  if (s.pro == PRO_GATE) {
    const float *const g4[4] = {lnv ? lnv : s.ln_g[0], lnv ? lnv + KC : s.ln_g[1], lnv ? lnv + 2 * KC : s.ln_g[2],
                                lnv ? lnv + 3 * KC : s.ln_g[3]};
    const float *const b4[4] = {lnv ? lnv + 4 * KC : s.ln_b[0], lnv ? lnv + 5 * KC : s.ln_b[1],
                                lnv ? lnv + 6 * KC : s.ln_b[2], lnv ? lnv + 7 * KC : s.ln_b[3]};
    ln4_inplace(r.v, klen, lane, g4, b4);
#pragma unroll
    for (int i = 0; i < KPL; ++i)
      v[i] = (kidx(lane, i) < klen) ? sigmoidf_(r.v[0][i]) * r.v[1][i] + sigmoidf_(r.v[2][i]) * r.v[3][i] : 0.f;
    return;
  }
  }
#pragma unroll
  for (int i = 0; i < KPL; ++i) v[i] = r.v[0][i];
  if (s.pro == PRO_LN || s.pro == PRO_LN_RELU) {
    ln_inplace(v, klen, lane, lnv ? lnv : s.ln_g[0], lnv ? lnv + 4 * KC : s.ln_b[0]);
    if (s.pro == PRO_LN_RELU) {
#pragma unroll
      for (int i = 0; i < KPL; ++i) v[i] = fmaxf(v[i], 0.f);
    }
  }
}

__device__ __forceinline__ void store_pairs(float *dst, int klen, int lane, const float (&v)[KPL]) {
#pragma unroll
  for (int p = 0; p < KPL / 2; ++p) {
    const int k = kidx(lane, 2 * p);
    if (k < klen) *reinterpret_cast<float2 *>(dst + k) = make_float2(v[2 * p], v[2 * p + 1]);
  }
}


}  // namespace vkn
