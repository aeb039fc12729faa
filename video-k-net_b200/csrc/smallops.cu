// The N-kernel "rows" operators of the KernelUpdateHead stage: every Linear of KernelUpdator, MHSA,
// FFN and the cls/mask heads runs through ONE fused kernel whose prologue applies the row-wise
// transform feeding it (LayerNorm / ReLU / gate arithmetic, warp-shuffle reductions, staged through
// shared memory) and whose epilogue adds bias / residual / ReLU.  Weights are streamed with 16-byte
// vector loads; activations stay fp32.
//
// Reference math: knet/kernel_updator.py:56-94, knet/det/kernel_update_head.py:201-227.
#include "common.cuh"

namespace vkn {

constexpr int KC = 256;   // K-chunk held in the shared-memory row panel
constexpr int BK = 32;    // K-step of the weight tile
constexpr int NT = 128;   // threads per CTA
constexpr int KPL = KC / 32;  // panel elements per lane

// ---- row-wise prologue ------------------------------------------------------------------------
__device__ __forceinline__ void ln_inplace(float (&v)[KPL], int K, int lane, const float *g, const float *b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < KPL; ++i) s += v[i];  // out-of-range lanes hold 0
  const float mean = warp_sum(s) / (float)K;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
    const int k = lane + 32 * i;
    const float d = (k < K) ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)K + 1e-5f);
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
    const int k = lane + 32 * i;
    if (k < K) v[i] = (v[i] - mean) * rstd * __ldg(g + k) + __ldg(b + k);
  }
}

__device__ __forceinline__ void fetch_plain(const float *a, int lda, int row, int k0, int klen, int lane,
                                            float (&v)[KPL]) {
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
    const int k = lane + 32 * i;
    v[i] = (k < klen) ? __ldg(a + (size_t)row * lda + k0 + k) : 0.f;
  }
}

// One warp produces the transformed values of one row for columns k0 .. k0+klen (klen <= KC).
// LN-type modes require k0 == 0 and klen == K (host-checked).
__device__ __forceinline__ void row_transform(const RowSrc &s, int row, int k0, int klen, int lane,
                                              float (&v)[KPL]) {
  if (s.pro == PRO_MUL) {
    float b[KPL];
    fetch_plain(s.a[0], s.lda[0], row, k0, klen, lane, v);
    fetch_plain(s.a[1], s.lda[1], row, k0, klen, lane, b);
#pragma unroll
    for (int i = 0; i < KPL; ++i) v[i] *= b[i];
    return;
  }
  if (s.pro == PRO_GATE) {
    float t[KPL], acc[KPL];
    fetch_plain(s.a[0], s.lda[0], row, 0, klen, lane, v);       // update gate pre-activation
    ln_inplace(v, klen, lane, s.ln_g[0], s.ln_b[0]);
    fetch_plain(s.a[1], s.lda[1], row, 0, klen, lane, t);       // param_out
    ln_inplace(t, klen, lane, s.ln_g[1], s.ln_b[1]);
#pragma unroll
    for (int i = 0; i < KPL; ++i) acc[i] = sigmoidf_(v[i]) * t[i];
    fetch_plain(s.a[2], s.lda[2], row, 0, klen, lane, v);       // input gate pre-activation
    ln_inplace(v, klen, lane, s.ln_g[2], s.ln_b[2]);
    fetch_plain(s.a[3], s.lda[3], row, 0, klen, lane, t);       // input_out
    ln_inplace(t, klen, lane, s.ln_g[3], s.ln_b[3]);
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const int k = lane + 32 * i;
      v[i] = (k < klen) ? acc[i] + sigmoidf_(v[i]) * t[i] : 0.f;
    }
    return;
  }
  // PRO_COPY / PRO_LN / PRO_LN_RELU: sum of slices (+ bias + residual) first
#pragma unroll
  for (int i = 0; i < KPL; ++i) v[i] = 0.f;
  for (int sl = 0; sl < s.nsum; ++sl) {
    const float *a = s.a[0] + (size_t)sl * s.sum_stride + (size_t)row * s.lda[0] + k0;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const int k = lane + 32 * i;
      if (k < klen) v[i] += __ldg(a + k);
    }
  }
  if (s.pbias) {
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const int k = lane + 32 * i;
      if (k < klen) v[i] += __ldg(s.pbias + k0 + k);
    }
  }
  if (s.pres) {
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const int k = lane + 32 * i;
      if (k < klen) v[i] += __ldg(s.pres + (size_t)row * s.ldpres + k0 + k);
    }
  }
  if (s.pro == PRO_LN || s.pro == PRO_LN_RELU) {
    ln_inplace(v, klen, lane, s.ln_g[0], s.ln_b[0]);
    if (s.pro == PRO_LN_RELU) {
#pragma unroll
      for (int i = 0; i < KPL; ++i) v[i] = fmaxf(v[i], 0.f);
    }
  }
}

// ---- standalone row operator (materialises a prologue result) ----------------------------------
__global__ void __launch_bounds__(NT) vkn_rowop_kernel(const __grid_constant__ RowSrc src, float *out, int ldo,
                                                       int M, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (NT / 32) + warp;
  if (row >= M) return;
  for (int k0 = 0; k0 < K; k0 += KC) {
    const int klen = min(KC, K - k0);
    float v[KPL];
    row_transform(src, row, k0, klen, lane, v);
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const int k = lane + 32 * i;
      if (k < klen) out[(size_t)row * ldo + k0 + k] = v[i];
    }
  }
}

// ---- fused rows x Linear ----------------------------------------------------------------------
struct LinBatch {
  LinArgs p[2];
};

template <typename WT, int BM, int BN>
__global__ void __launch_bounds__(NT) vkn_linear_kernel(const __grid_constant__ LinBatch batch) {
  constexpr int TM = BM / (NT / 16);
  constexpr int TN = BN / 16;
  static_assert(TM >= 1 && TN >= 1, "tile too small");
  __shared__ __align__(16) float As[BM][KC + 4];
  __shared__ __align__(16) float Ws[BN][BK + 4];

  const int ks_total = batch.p[0].ksplit;
  const LinArgs &A = batch.p[blockIdx.z / ks_total];
  const int ks = blockIdx.z % ks_total;
  const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
  if (row0 >= A.M || col0 >= A.N) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tx = tid & 15, ty = tid >> 4;

  int kper = (A.K + ks_total - 1) / ks_total;
  kper = (kper + BK - 1) / BK * BK;
  const int kbeg = ks * kper, kend = min(A.K, kbeg + kper);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const WT *Wp = reinterpret_cast<const WT *>(A.w);

  for (int kc0 = kbeg; kc0 < kend; kc0 += KC) {
    const int kclen = min(KC, kend - kc0);
    // ---- panel: transformed rows row0..row0+BM, columns kc0..kc0+kclen (zero padded)
    for (int r = warp; r < BM; r += NT / 32) {
      const int row = row0 + r;
      float v[KPL];
      if (row < A.M) {
        row_transform(A.src, row, kc0, kclen, lane, v);
      } else {
#pragma unroll
        for (int i = 0; i < KPL; ++i) v[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < KPL; ++i) As[r][lane + 32 * i] = v[i];
      if (A.side != nullptr && blockIdx.x == 0 && row < A.M) {
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
          const int k = lane + 32 * i;
          if (k < kclen) A.side[(size_t)row * A.ldside + kc0 + k] = v[i];
        }
      }
    }
    __syncthreads();
    for (int k0 = 0; k0 < kclen; k0 += BK) {
      // ---- weight tile: BN rows x BK columns, 8 consecutive k per thread (16-byte bf16 loads)
#pragma unroll
      for (int pass = 0; pass < BN / (NT / 4); ++pass) {
        const int n = pass * (NT / 4) + (tid >> 2);
        const int kk = (tid & 3) * 8;
        const int gk = kc0 + k0 + kk;
        float w[8];
        if (col0 + n < A.N && gk + 8 <= kend) {
          load8(Wp + (size_t)(col0 + n) * A.ldw + gk, w);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            w[e] = (col0 + n < A.N && gk + e < kend) ? to_f32(Wp[(size_t)(col0 + n) * A.ldw + gk + e]) : 0.f;
        }
        *reinterpret_cast<float4 *>(&Ws[n][kk]) = make_float4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<float4 *>(&Ws[n][kk + 4]) = make_float4(w[4], w[5], w[6], w[7]);
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; kk += 4) {
        float4 a4[TM], b4[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a4[i] = *reinterpret_cast<const float4 *>(&As[ty + (NT / 16) * i][k0 + kk]);
#pragma unroll
        for (int j = 0; j < TN; ++j) b4[j] = *reinterpret_cast<const float4 *>(&Ws[tx + 16 * j][kk]);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) {
            acc[i][j] = fmaf(a4[i].x, b4[j].x, acc[i][j]);
            acc[i][j] = fmaf(a4[i].y, b4[j].y, acc[i][j]);
            acc[i][j] = fmaf(a4[i].z, b4[j].z, acc[i][j]);
            acc[i][j] = fmaf(a4[i].w, b4[j].w, acc[i][j]);
          }
      }
      __syncthreads();
    }
  }

  float *outp = A.out + (size_t)ks * A.out_split_stride;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = row0 + ty + (NT / 16) * i;
    if (row >= A.M) continue;
    const float rs = (A.epi & EPI_ROWSCALE) ? __ldg(A.rowscale + row) : 1.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = col0 + tx + 16 * j;
      if (col >= A.N) continue;
      float v = acc[i][j];
      if (A.epi & EPI_BIAS) v += rs * __ldg(A.bias + col);
      if (A.epi & EPI_RES) v += __ldg(A.res + (size_t)row * A.ldres + col);
      if (A.epi & EPI_RELU) v = fmaxf(v, 0.f);
      outp[(size_t)row * A.ldo + col] = v;
    }
  }
}

static int check_src(const RowSrc &s, int K) {
  if (s.pro != PRO_COPY && K > KC) VKN_FAIL(VKN_E_UNSUPPORTED, "row transform %d needs K <= %d (got %d)", s.pro, KC, K);
  if (s.nsum < 1) VKN_FAIL(VKN_E_INVALID, "RowSrc.nsum must be >= 1");
  return VKN_OK;
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
    if (b.p[i].ldw % 8 != 0 || (reinterpret_cast<uintptr_t>(b.p[i].w) & 15))
      VKN_FAIL(VKN_E_INVALID, "launch_linear: weight rows must be 16-byte aligned (ldw %d)", b.p[i].ldw);
    maxM = max(maxM, b.p[i].M);
    maxN = max(maxN, b.p[i].N);
  }
  if (nprob == 1) b.p[1] = b.p[0];
  const int ks = b.p[0].ksplit;
  // tile choice: wide outputs get the 32x64 tile, the C x C layers the 16x32 tile (more CTAs in flight)
  const bool big = maxN >= 1024;
  dim3 grid, block(NT);
  VKN_LAUNCH_MARK(big ? "vkn_linear_kernel<32x64>" : "vkn_linear_kernel<16x32>", stream);
  if (big) {
    grid = dim3(ceil_div(maxN, 64), ceil_div(maxM, 32), nprob * ks);
    if (w_dtype == VKN_BF16) vkn_linear_kernel<__nv_bfloat16, 32, 64><<<grid, block, 0, stream>>>(b);
    else vkn_linear_kernel<float, 32, 64><<<grid, block, 0, stream>>>(b);
  } else {
    grid = dim3(ceil_div(maxN, 32), ceil_div(maxM, 16), nprob * ks);
    if (w_dtype == VKN_BF16) vkn_linear_kernel<__nv_bfloat16, 16, 32><<<grid, block, 0, stream>>>(b);
    else vkn_linear_kernel<float, 16, 32><<<grid, block, 0, stream>>>(b);
  }
  VKN_CUDA_OK(cudaGetLastError());
  return VKN_OK;
}

int launch_rowop(const RowSrc &src, float *out, int ldo, int M, int K, cudaStream_t stream) {
  VKN_TRY(check_src(src, K));
  VKN_LAUNCH_MARK("vkn_rowop_kernel", stream);
  vkn_rowop_kernel<<<ceil_div(M, NT / 32), NT, 0, stream>>>(src, out, ldo, M, K);
  VKN_CUDA_OK(cudaGetLastError());
  return VKN_OK;
}

// ---- attention among the N kernels of a frame ---------------------------------------------------
// One CTA = (query block, head, frame).  K and V of the head sit in shared memory ([N][hd+1], bank
// conflict free); one warp per query: lane-parallel scores, shuffle softmax, lane-per-channel PV.
// torch semantics (F.multi_head_attention_forward): q scaled by 1/sqrt(hd) BEFORE q.k^T.
constexpr int ATT_QB = 16;

__global__ void __launch_bounds__(NT) vkn_attention_kernel(const float *__restrict__ q, int ldq,
                                                           const float *__restrict__ k, int ldk,
                                                           const float *__restrict__ v, int ldv,
                                                           float *__restrict__ out, int ldo, int N, int hd,
                                                           float scale) {
  extern __shared__ float smem[];
  const int hs = hd + 1;
  float *Ks = smem;                    // [N][hs]
  float *Vs = Ks + (size_t)N * hs;     // [N][hs]
  float *Ps = Vs + (size_t)N * hs;     // [4 warps][N]
  float *Qs = Ps + 4 * (size_t)N;      // [4 warps][32]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT_QB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t rowb = (size_t)b * N;
  for (int idx = tid; idx < N * hd; idx += NT) {
    const int j = idx / hd, d = idx - j * hd;
    Ks[j * hs + d] = __ldg(k + (rowb + j) * ldk + h * hd + d);
    Vs[j * hs + d] = __ldg(v + (rowb + j) * ldv + h * hd + d);
  }
  __syncthreads();
  float *ps = Ps + (size_t)warp * N;
  float *qs = Qs + warp * 32;
  for (int qi = q0 + warp; qi < min(N, q0 + ATT_QB); qi += NT / 32) {
    if (lane < hd) qs[lane] = __ldg(q + (rowb + qi) * ldq + h * hd + lane) * scale;
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) {
      float s = 0.f;
      for (int d = 0; d < hd; ++d) s = fmaf(qs[d], Ks[j * hs + d], s);
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
      float o = 0.f;
      for (int j = 0; j < N; ++j) o = fmaf(ps[j], Vs[j * hs + lane], o);
      out[(rowb + qi) * ldo + h * hd + lane] = o / sum;
    }
    __syncwarp();
  }
}

int launch_attention(const float *q, int ldq, const float *k, int ldk, const float *v, int ldv, float *out,
                     int ldo, int B, int N, int C, int heads, cudaStream_t stream) {
  if (heads < 1 || C % heads != 0) VKN_FAIL(VKN_E_INVALID, "attention: C %d not divisible by heads %d", C, heads);
  const int hd = C / heads;
  if (hd > 32) VKN_FAIL(VKN_E_UNSUPPORTED, "attention: head_dim %d > 32", hd);
  const size_t smem = ((size_t)2 * N * (hd + 1) + 4 * (size_t)N + 4 * 32) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    VKN_CUDA_OK(cudaFuncSetAttribute(vkn_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  if (smem > 160 * 1024) VKN_FAIL(VKN_E_UNSUPPORTED, "attention: N %d too large for shared memory", N);
  dim3 grid(ceil_div(N, ATT_QB), heads, B);
  VKN_LAUNCH_MARK("vkn_attention_kernel", stream);
  vkn_attention_kernel<<<grid, NT, smem, stream>>>(q, ldq, k, ldk, v, ldv, out, ldo, N, hd,
                                                   1.0f / sqrtf((float)hd));
  VKN_CUDA_OK(cudaGetLastError());
  return VKN_OK;
}

}  // namespace vkn
