// CUDA-core fp32 engines for the two big contractions of a stage.  They serve fp32 storage
// (VKN_F32), shapes the tcgen05 engine does not take, and are the on-device cross-check of the
// tensor-core engine.
//
//   pooling  (knet/det/kernel_update_head.py:190-195, feat_transform folded out):
//       xp0[b,n,c] = sum_p 1[mask[b,n,p] > thr] * x[b,c,p]        cnt[b,n] = sum_p 1[...]
//   mask conv (knet/det/kernel_update_head.py:247-260, feat_transform folded in):
//       out[b,n,p] = sum_c a_ext[b*N+n, c] * x[b,c,p] + a_ext[b*N+n, C]
#include "common.cuh"

namespace vkn {

constexpr int GT = 256;       // threads per CTA
constexpr int TILE_N = 128;   // kernels per CTA tile
constexpr int TILE_C = 64;    // channels per CTA tile (pooling)
constexpr int TILE_P = 64;    // pixels per CTA tile (mask conv)
constexpr int PK = 32;        // inner step
constexpr int POOL_CHUNK = 256;  // pixels reduced by one pooling CTA

// ---- pooling -----------------------------------------------------------------------------------
template <typename XT>
__global__ void __launch_bounds__(GT) vkn_pool_simt_kernel(const XT *__restrict__ x, const XT *__restrict__ mask,
                                                           float *__restrict__ partials,
                                                           float *__restrict__ cnt_partials, int B, int N, int C,
                                                           int HW, float thr) {
  __shared__ __align__(16) float Ms[TILE_N][PK + 4];
  __shared__ __align__(16) float Xs[TILE_C][PK + 4];
  const int chunk = blockIdx.x, b = blockIdx.z;
  const int cblocks = C / TILE_C;
  const int cb = blockIdx.y % cblocks, nb = blockIdx.y / cblocks;
  const int n0 = nb * TILE_N, c0 = cb * TILE_C;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int pbeg = chunk * POOL_CHUNK, pend = min(HW, pbeg + POOL_CHUNK);
  pdl_wait();

  float acc[8][4];
  float cacc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    cacc[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  }
  const XT *mb = mask + (size_t)b * N * HW;
  const XT *xb = x + (size_t)b * C * HW;
  for (int p0 = pbeg; p0 < pend; p0 += PK) {
    {
      const int kk = tid & 31;
      const int p = p0 + kk;
#pragma unroll
      for (int i = 0; i < TILE_N / 8; ++i) {
        const int r = (tid >> 5) + 8 * i;
        const int n = n0 + r;
        float m = 0.f;
        if (n < N && p < pend) m = (to_f32(mb[(size_t)n * HW + p]) > thr) ? 1.f : 0.f;
        Ms[r][kk] = m;
      }
#pragma unroll
      for (int i = 0; i < TILE_C / 8; ++i) {
        const int r = (tid >> 5) + 8 * i;
        float v = 0.f;
        if (p < pend) v = to_f32(xb[(size_t)(c0 + r) * HW + p]);
        Xs[r][kk] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < PK; kk += 4) {
      float4 a4[8], b4[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a4[i] = *reinterpret_cast<const float4 *>(&Ms[ty + 16 * i][kk]);
#pragma unroll
      for (int j = 0; j < 4; ++j) b4[j] = *reinterpret_cast<const float4 *>(&Xs[tx + 16 * j][kk]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        cacc[i] += (a4[i].x + a4[i].y) + (a4[i].z + a4[i].w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = fmaf(a4[i].x, b4[j].x, acc[i][j]);
          acc[i][j] = fmaf(a4[i].y, b4[j].y, acc[i][j]);
          acc[i][j] = fmaf(a4[i].z, b4[j].z, acc[i][j]);
          acc[i][j] = fmaf(a4[i].w, b4[j].w, acc[i][j]);
        }
      }
    }
    __syncthreads();
  }
  pdl_trigger();
  float *po = partials + ((size_t)chunk * B + b) * N * C;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = n0 + ty + 16 * i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) po[(size_t)n * C + c0 + tx + 16 * j] = acc[i][j];
    if (cb == 0 && tx == 0) cnt_partials[((size_t)chunk * B + b) * N + n] = cacc[i];
  }
}

int pool_simt_chunks(const VknShape &s) { return ceil_div(s.H * s.W, POOL_CHUNK); }

int launch_pool_simt(const VknShape &s, const void *x, const void *mask, float *partials, float *cnt_partials,
                     int *nchunks, cudaStream_t stream) {
  const int HW = s.H * s.W;
  if (s.C % TILE_C != 0) VKN_FAIL(VKN_E_UNSUPPORTED, "pool: C %d must be a multiple of %d", s.C, TILE_C);
  *nchunks = pool_simt_chunks(s);
  dim3 grid(*nchunks, (s.C / TILE_C) * ceil_div(s.N, TILE_N), s.B);
  VKN_LAUNCH_MARK("vkn_pool_simt_kernel", stream);
  if (s.x_dtype == VKN_BF16)
    VKN_CUDA_OK(launch_chain(vkn_pool_simt_kernel<__nv_bfloat16>, grid, dim3(GT), 0, stream, (const __nv_bfloat16 *)x,
                             (const __nv_bfloat16 *)mask, partials, cnt_partials, s.B, s.N, s.C, HW, s.mask_thr_logit));
  else
    VKN_CUDA_OK(launch_chain(vkn_pool_simt_kernel<float>, grid, dim3(GT), 0, stream, (const float *)x,
                             (const float *)mask, partials, cnt_partials, s.B, s.N, s.C, HW, s.mask_thr_logit));
  return VKN_OK;
}

// partials [nchunks][B*F frames][N*C] -> xp0 [B][N*C];  cnt_partials [nchunks][B*F][N] -> cnt [B][N].
// A kernel set b folds its nchunks * F slices (F = frames_per_set; the clip head averages the pooled
// feature over its frames, knet_vis/tracker/kernel_update_head.py:245-246 -> scale = 1/F).
// One CTA = 128 outputs (float4 per lane) of one set; its 8 warps each sum every 8th slice with four
// independent 16-byte loads in flight, then the 8 warp sums are combined in a FIXED order
// (deterministic, no atomics).
__global__ void __launch_bounds__(256) vkn_pool_reduce_kernel(const float *__restrict__ partials,
                                                              const float *__restrict__ cnt_partials, int nchunks,
                                                              int B, int F, int N, int NC, float scale,
                                                              float *__restrict__ xp0, float *__restrict__ cnt,
                                                              __nv_bfloat16 *__restrict__ planes) {
  __shared__ float4 red[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int idx = blockIdx.x * 128 + lane * 4;
  const int nsl = nchunks * F;
  pdl_wait();
  pdl_trigger();      // the consumer's prologue (weight ring of the cluster chain, ~3 us) overlaps this kernel; it waits for our
                      // completion at its own dependency wait
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  auto slice = [&](int sl) -> size_t {       // slice (chunk, f) of set b -> frame-slice index in the partials
    if (F == 1) return (size_t)sl * B + b;
    const int ch = sl / F, f = sl - ch * F;
    return ((size_t)ch * B * F + (size_t)b * F + f);
  };
  if (idx < NC) {
    float4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    int sl = warp;
    for (; sl + 56 < nsl; sl += 64) {        // eight independent 16-byte loads in flight per thread, same association order
      float4 t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = __ldg(reinterpret_cast<const float4 *>(partials + slice(sl + 8 * u) * NC + idx));
#pragma unroll
      for (int u = 0; u < 8; ++u) { a[u & 3].x += t[u].x; a[u & 3].y += t[u].y; a[u & 3].z += t[u].z; a[u & 3].w += t[u].w; }
    }
    for (; sl + 24 < nsl; sl += 32) {
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = __ldg(reinterpret_cast<const float4 *>(partials + slice(sl + 8 * u) * NC + idx));
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u].x += t[u].x; a[u].y += t[u].y; a[u].z += t[u].z; a[u].w += t[u].w; }
    }
    for (int u = 0; sl < nsl; sl += 8, ++u) {
      const float4 t = __ldg(reinterpret_cast<const float4 *>(partials + slice(sl) * NC + idx));
      a[u & 3].x += t.x; a[u & 3].y += t.y; a[u & 3].z += t.z; a[u & 3].w += t.w;
    }
    acc = make_float4((a[0].x + a[1].x) + (a[2].x + a[3].x), (a[0].y + a[1].y) + (a[2].y + a[3].y),
                      (a[0].z + a[1].z) + (a[2].z + a[3].z), (a[0].w + a[1].w) + (a[2].w + a[3].w));
  }
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && idx < NC) {
    float4 t = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 u = red[w][lane];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    const float v4[4] = {t.x * scale, t.y * scale, t.z * scale, t.w * scale};
    *reinterpret_cast<float4 *>(xp0 + (size_t)b * NC + idx) = make_float4(v4[0], v4[1], v4[2], v4[3]);
    if (planes != nullptr) {       // A operand of the tcgen05 row GEMM that follows: bf16 hi/mid/lo planes [3][B*N][C]
      uint16_t h[3][4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float xr = v4[e];
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
          const __nv_bfloat16 hb = __float2bfloat16_rn(xr);
          xr -= __bfloat162float(hb);
          h[pl][e] = __bfloat16_as_ushort(hb);
        }
      }
#pragma unroll
      for (int pl = 0; pl < 3; ++pl)
        *reinterpret_cast<uint2 *>(planes + (size_t)pl * B * NC + (size_t)b * NC + idx) =
            make_uint2((uint32_t)h[pl][0] | ((uint32_t)h[pl][1] << 16), (uint32_t)h[pl][2] | ((uint32_t)h[pl][3] << 16));
    }
  }
  // pixel counts of the set: CTAs x < ceil(N/32) each own 32 kernels; warp w sums slices w, w+8, ...
  // (lane = kernel), then the 8 warp sums are combined in fixed order.
  __shared__ float cred[8][32];
  const int n = blockIdx.x * 32 + lane;
  if (blockIdx.x * 32 < N) {
    float c0 = 0.f, c1 = 0.f;
    if (n < N) {
      int sl = warp;
      for (; sl + 8 < nsl; sl += 16) {
        c0 += __ldg(cnt_partials + slice(sl) * N + n);
        c1 += __ldg(cnt_partials + slice(sl + 8) * N + n);
      }
      if (sl < nsl) c0 += __ldg(cnt_partials + slice(sl) * N + n);
    }
    cred[warp][lane] = c0 + c1;
    __syncthreads();
    if (warp == 0 && n < N) {
      float t = cred[0][lane];
#pragma unroll
      for (int w = 1; w < 8; ++w) t += cred[w][lane];
      cnt[(size_t)b * N + n] = t * scale;
    }
  }
}

// Few slices per set (frame batches: 148 / B pixel chunks per frame): one thread per 4 outputs sums its slices in
// order -- same fixed order as the kernel above (slice 0, 1, 2, ... -> identical bits), no shared-memory stage, every
// thread busy.  Requires F == 1.
__global__ void __launch_bounds__(256) vkn_pool_reduce_flat_kernel(const float *__restrict__ partials,
                                                                   const float *__restrict__ cnt_partials, int nchunks,
                                                                   int B, int N, int NC, float *__restrict__ xp0,
                                                                   float *__restrict__ cnt,
                                                                   __nv_bfloat16 *__restrict__ planes) {
  pdl_wait();
  const size_t total4 = (size_t)B * NC / 4;
  const size_t i4 = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i4 < total4) {
    const size_t stride4 = total4;                         // one slice = [B][N][C]
    const float4 *p = reinterpret_cast<const float4 *>(partials) + i4;
    float4 t[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < nchunks) t[c] = __ldg(p + (size_t)c * stride4);
    float4 a = t[0];
#pragma unroll
    for (int c = 1; c < 8; ++c)
      if (c < nchunks) { a.x += t[c].x; a.y += t[c].y; a.z += t[c].z; a.w += t[c].w; }
    reinterpret_cast<float4 *>(xp0)[i4] = a;
    if (planes != nullptr) {
      const float v4[4] = {a.x, a.y, a.z, a.w};
      uint16_t h[3][4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float xr = v4[e];
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
          const __nv_bfloat16 hb = __float2bfloat16_rn(xr);
          xr -= __bfloat162float(hb);
          h[pl][e] = __bfloat16_as_ushort(hb);
        }
      }
#pragma unroll
      for (int pl = 0; pl < 3; ++pl)
        *reinterpret_cast<uint2 *>(planes + (size_t)pl * B * NC + i4 * 4) =
            make_uint2((uint32_t)h[pl][0] | ((uint32_t)h[pl][1] << 16), (uint32_t)h[pl][2] | ((uint32_t)h[pl][3] << 16));
    }
  }
  const size_t ic = (size_t)blockIdx.x * 256 + threadIdx.x;   // pixel counts: the first B * N threads
  if (ic < (size_t)B * N) {
    float c = 0.f;
    for (int ch = 0; ch < nchunks; ++ch) c += __ldg(cnt_partials + (size_t)ch * B * N + ic);
    cnt[ic] = c;
  }
  pdl_trigger();
}

int launch_pool_reduce(const VknShape &s, const float *partials, const float *cnt_partials, int nchunks,
                       float *xp0, float *cnt, cudaStream_t stream, void *planes) {
  const int F = s.frames_per_set > 1 ? s.frames_per_set : 1;
  const int NC = s.N * s.C;
  if (F == 1 && nchunks <= 8 && NC % 4 == 0) {
    const size_t total4 = (size_t)s.B * NC / 4;
    VKN_LAUNCH_MARK("vkn_pool_reduce_flat_kernel", stream);
    VKN_CUDA_OK(launch_chain(vkn_pool_reduce_flat_kernel, dim3((unsigned)((total4 + 255) / 256)), dim3(256), 0, stream, partials,
                             cnt_partials, nchunks, s.B, s.N, NC, xp0, cnt, (__nv_bfloat16 *)planes));
    return VKN_OK;
  }
  VKN_LAUNCH_MARK("vkn_pool_reduce_kernel", stream);
  int gx = ceil_div(NC, 128);
  if (gx < ceil_div(s.N, 32)) gx = ceil_div(s.N, 32);
  VKN_CUDA_OK(launch_chain(vkn_pool_reduce_kernel, dim3(gx, s.B), dim3(256), 0, stream, partials, cnt_partials, nchunks,
                           s.B, F, s.N, NC, 1.0f / (float)F, xp0, cnt, (__nv_bfloat16 *)planes));
  return VKN_OK;
}

// ---- static kernels -> mask-conv operand (vkn_init_proposals) ----------------------------------------------------
__global__ void __launch_bounds__(256) vkn_pack_kernels_kernel(const float *__restrict__ w, const float *__restrict__ b,
                                                               int N, int C, float *__restrict__ a_ext, int lda,
                                                               __nv_bfloat16 *__restrict__ planes, int Npad) {
  pdl_wait();
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx < N * C) {
    const int n = idx / C, c = idx - n * C;
    const float v = w[idx];
    a_ext[(size_t)n * lda + c] = v;
    if (c == 0) a_ext[(size_t)n * lda + C] = b ? b[n] : 0.f;
    if (planes != nullptr) {
      const size_t plane = (size_t)Npad * C, o = (size_t)n * C + c;
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const float r1 = v - __bfloat162float(hi);
      const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
      planes[o] = hi;
      planes[plane + o] = mid;
      planes[2 * plane + o] = __float2bfloat16_rn(r1 - __bfloat162float(mid));
    }
  }
  pdl_trigger();
}

int launch_pack_kernels(const float *w, const float *b, int N, int C, float *a_ext, int lda, void *planes, int Npad,
                        cudaStream_t stream) {
  VKN_LAUNCH_MARK("vkn_pack_kernels_kernel", stream);
  VKN_CUDA_OK(launch_chain(vkn_pack_kernels_kernel, dim3(ceil_div(N * C, 256)), dim3(256), 0, stream, w, b, N, C, a_ext, lda,
                           (__nv_bfloat16 *)planes, Npad));
  return VKN_OK;
}

// ---- dynamic mask conv ---------------------------------------------------------------------------
template <typename XT, bool VEC>
__global__ void __launch_bounds__(GT) vkn_maskgemm_simt_kernel(const XT *__restrict__ x,
                                                               const float *__restrict__ a_ext, int lda,
                                                               XT *__restrict__ out, int N, int C, int HW, int F) {
  __shared__ __align__(16) float As[TILE_N][PK + 4];
  __shared__ __align__(16) float Xs[PK][TILE_P + 4];
  const int p0 = blockIdx.x * TILE_P, n0 = blockIdx.y * TILE_N, b = blockIdx.z;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const XT *xb = x + (size_t)b * C * HW;
  const float *ab = a_ext + (size_t)(b / F) * N * lda;      // frame b uses the kernels of set b / F
  pdl_wait();
  for (int c0 = 0; c0 < C; c0 += PK) {
    {  // A tile: 128 rows x 32 channels, 8 consecutive channels per thread
      const int kk = (tid & 3) * 8;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = (tid >> 2) + 64 * i;
        const int n = n0 + r;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (n < N) {
          v0 = *reinterpret_cast<const float4 *>(ab + (size_t)n * lda + c0 + kk);
          v1 = *reinterpret_cast<const float4 *>(ab + (size_t)n * lda + c0 + kk + 4);
        }
        *reinterpret_cast<float4 *>(&As[r][kk]) = v0;
        *reinterpret_cast<float4 *>(&As[r][kk + 4]) = v1;
      }
      // x tile: 32 channels x 64 pixels, pixel-contiguous
      const int px = tid & 63;
#pragma unroll
      for (int i = 0; i < PK / 4; ++i) {
        const int kc = (tid >> 6) + 4 * i;
        float v = 0.f;
        if (p0 + px < HW) v = to_f32(xb[(size_t)(c0 + kc) * HW + p0 + px]);
        Xs[kc][px] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < PK; kk += 4) {
      float4 a4[8], b4[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a4[i] = *reinterpret_cast<const float4 *>(&As[ty + 16 * i][kk]);
#pragma unroll
      for (int e = 0; e < 4; ++e) b4[e] = *reinterpret_cast<const float4 *>(&Xs[kk + e][tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float av[4] = {a4[i].x, a4[i].y, a4[i].z, a4[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[i][0] = fmaf(av[e], b4[e].x, acc[i][0]);
          acc[i][1] = fmaf(av[e], b4[e].y, acc[i][1]);
          acc[i][2] = fmaf(av[e], b4[e].z, acc[i][2]);
          acc[i][3] = fmaf(av[e], b4[e].w, acc[i][3]);
        }
      }
    }
    __syncthreads();
  }
  pdl_trigger();
  XT *ob = out + (size_t)b * N * HW;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = n0 + ty + 16 * i;
    if (n >= N) continue;
    const float bias = __ldg(ab + (size_t)n * lda + C);
    const int p = p0 + tx * 4;
    float v[4] = {acc[i][0] + bias, acc[i][1] + bias, acc[i][2] + bias, acc[i][3] + bias};
    if (VEC) {
      if (p < HW) store4(ob + (size_t)n * HW + p, v);  // HW % 4 == 0: a 4-group never straddles the end
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (p + e < HW) store_as<XT>(ob + (size_t)n * HW + p + e, v[e]);
    }
  }
}

int launch_maskgemm_simt(const VknShape &s, const void *x, const float *a_ext, int lda, void *out,
                         cudaStream_t stream) {
  const int HW = s.H * s.W;
  if (s.C % PK != 0) VKN_FAIL(VKN_E_UNSUPPORTED, "mask gemm: C %d must be a multiple of %d", s.C, PK);
  if (lda % 4 != 0) VKN_FAIL(VKN_E_INVALID, "mask gemm: lda %d must be a multiple of 4", lda);
  const int F = s.frames_per_set > 1 ? s.frames_per_set : 1;
  dim3 grid(ceil_div(HW, TILE_P), ceil_div(s.N, TILE_N), s.B * F);
  VKN_LAUNCH_MARK("vkn_maskgemm_simt_kernel", stream);
  const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (s.x_dtype == VKN_BF16) {
    auto xp = (const __nv_bfloat16 *)x;
    auto op = (__nv_bfloat16 *)out;
    if (vec) VKN_CUDA_OK(launch_chain(vkn_maskgemm_simt_kernel<__nv_bfloat16, true>, grid, dim3(GT), 0, stream, xp, a_ext, lda, op, s.N, s.C, HW, F));
    else VKN_CUDA_OK(launch_chain(vkn_maskgemm_simt_kernel<__nv_bfloat16, false>, grid, dim3(GT), 0, stream, xp, a_ext, lda, op, s.N, s.C, HW, F));
  } else {
    auto xp = (const float *)x;
    auto op = (float *)out;
    if (vec) VKN_CUDA_OK(launch_chain(vkn_maskgemm_simt_kernel<float, true>, grid, dim3(GT), 0, stream, xp, a_ext, lda, op, s.N, s.C, HW, F));
    else VKN_CUDA_OK(launch_chain(vkn_maskgemm_simt_kernel<float, false>, grid, dim3(GT), 0, stream, xp, a_ext, lda, op, s.N, s.C, HW, F));
  }
  VKN_CUDA_OK(cudaGetLastError());
  return VKN_OK;
}

}  // namespace vkn
