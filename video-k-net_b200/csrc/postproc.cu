// Post-loop mask path: the last-stage x`mask_upsample_stride` bilinear upsample of
// `_mask_forward` (knet/det/kernel_iter_head.py:122-128) + `rescale_masks` (knet/det/kernel_update_head.py:443-458:
// sigmoid -> bilinear to batch_input_shape -> crop [:h,:w] -> bilinear to ori_shape) + the `> mask_thr` of get_seg_masks
// (:460-467), fused into ONE kernel: the reference materialises three full-resolution fp32 tensors per mask; here a CTA
// rebuilds, in shared memory, exactly the part of each intermediate image its 32x64 output tile depends on
// (logits -> sigmoid(up) -> batch-shape -> output) and writes probabilities and / or thresholded bytes once.
// Every interpolation uses torch's align_corners=False arithmetic (area_pixel_compute_source_index):
//   src = scale * (dst + 0.5) - 0.5, clamped at 0;  i0 = int(src), i1 = i0 + (i0 < in - 1), lambda = src - i0.
// HBM-bound on its output (K * Ho * Wo bytes or floats); no tensor-core work.
#include "common.cuh"

namespace vkn {

constexpr int RS_TY = 64, RS_TX = 128, RS_NT = 256;      // largest output tile; halved (down to 4 x 8) until the cone fits
constexpr int RS_MAX_ROWS = 128;         // rows per staged region (y-tap tables)
constexpr int RS_MAX_REGION = 6144;      // floats per staged region (three regions: 72 KB of dynamic shared memory)

struct RsParams {
  int K, H, W, up, S1h, S1w, Hb, Wb, h, w, Ho, Wo, ty, tx;
  int offS, off2, offT;      // float offsets of the staged regions inside dynamic shared memory (sized to the cone)
  float sc1h, sc1w, sc2h, sc2w, sc3h, sc3w, thr;
};

__device__ __forceinline__ float rs_src(float scale, int dst) {
  const float s = scale * ((float)dst + 0.5f) - 0.5f;
  return s < 0.f ? 0.f : s;
}
// first / last input index touched by output indices [d0, d1] (src is monotone in dst)
__device__ __forceinline__ void rs_range(float scale, int d0, int d1, int in, int &lo, int &hi) {
  lo = (int)rs_src(scale, d0);
  hi = (int)rs_src(scale, d1);
  hi = hi < in - 1 ? hi + 1 : in - 1;
  if (lo > in - 1) lo = in - 1;
}
// one axis of torch's bilinear (align_corners=False): first tap (relative to the staged region), second-tap step, weights
struct RsTap {
  int i0, step;
  float l0, l1;
};
__device__ __forceinline__ RsTap rs_tap(float scale, int dst, int in, int lo) {
  const float r = rs_src(scale, dst);
  const int i = (int)r;
  RsTap t;
  t.i0 = i - lo;
  t.step = i < in - 1 ? 1 : 0;
  t.l1 = r - (float)i;
  t.l0 = 1.f - t.l1;
  return t;
}

// y taps of a pass, one float4 per output row: {first-row offset (floats, as bits), second-row step (floats, as bits), l0, l1}
__device__ __forceinline__ void rs_fill_ytab(float4 *tab, float scale, int dst0, int n, int in, int lo, int pitch) {
  for (int i = threadIdx.x; i < n; i += RS_NT) {
    const RsTap t = rs_tap(scale, dst0 + i, in, lo);
    tab[i] = make_float4(__int_as_float(t.i0 * pitch), __int_as_float(t.step * pitch), t.l0, t.l1);
  }
}
__device__ __forceinline__ float rs_lerp_tab(const float *col, const float4 yt, int xstep, float xl0, float xl1) {
  const float *p = col + __float_as_int(yt.x);
  const float *q = p + __float_as_int(yt.y);
  return yt.z * (xl0 * p[0] + xl1 * p[xstep]) + yt.w * (xl0 * q[0] + xl1 * q[xstep]);
}

// Thread mapping without integer divisions: a thread owns columns x = cx, cx + 64, ... (its x taps are computed once per
// column) and walks rows y = ry, ry + 4, ...; the y taps of every pass are tabulated once per CTA.  `ident3`: the last
// resize is the identity (ori_shape == img_shape, the common case), so the batch-shape image IS the output and is never
// staged.  ~25 instructions per output pixel and pass (the first version spent 87, issue-bound: profiles/).
template <typename T>
__global__ void __launch_bounds__(RS_NT) vkn_rescale_masks_kernel(const T *__restrict__ masks, float *__restrict__ probs,
                                                                  uint8_t *__restrict__ bits, const RsParams P, int ident3) {
  extern __shared__ float rs_smem[];
  float *regL = rs_smem;                        // logits (only when up > 1)
  float *regS = rs_smem + P.offS;               // sigmoid(upsampled logits)
  float *reg2 = rs_smem + P.off2;               // batch-shape image, cropped (unused when the last resize is the identity)
  float4 *tab1 = reinterpret_cast<float4 *>(rs_smem + P.offT);     // y taps: logits -> S rows
  float4 *tab2 = tab1 + RS_MAX_ROWS;                               //         S -> batch-shape rows
  float4 *tab3 = tab2 + RS_MAX_ROWS;                               //         batch-shape -> output rows
  const int k = blockIdx.z;
  const int Y0 = blockIdx.y * P.ty, X0 = blockIdx.x * P.tx;
  const int Y1 = min(Y0 + P.ty, P.Ho) - 1, X1 = min(X0 + P.tx, P.Wo) - 1;
  const int cx = threadIdx.x & 63, ry = threadIdx.x >> 6;
  const float thr = P.thr;
  const int Wo = P.Wo;
  // dependency cone of the tile, level by level
  int y2lo, y2hi, x2lo, x2hi, yslo, yshi, xslo, xshi, yllo = 0, ylhi = 0, xllo = 0, xlhi = 0;
  if (ident3) {
    y2lo = Y0; y2hi = Y1; x2lo = X0; x2hi = X1;
  } else {
    rs_range(P.sc3h, Y0, Y1, P.h, y2lo, y2hi);
    rs_range(P.sc3w, X0, X1, P.w, x2lo, x2hi);
  }
  rs_range(P.sc2h, y2lo, y2hi, P.S1h, yslo, yshi);
  rs_range(P.sc2w, x2lo, x2hi, P.S1w, xslo, xshi);
  const int r2 = y2hi - y2lo + 1, c2 = x2hi - x2lo + 1, rS = yshi - yslo + 1, cS = xshi - xslo + 1;
  const T *mk = masks + (size_t)k * P.H * P.W;
  int cL = 0;
  if (P.up > 1) {
    rs_range(P.sc1h, yslo, yshi, P.H, yllo, ylhi);
    rs_range(P.sc1w, xslo, xshi, P.W, xllo, xlhi);
    cL = xlhi - xllo + 1;
    rs_fill_ytab(tab1, P.sc1h, yslo, rS, P.H, yllo, cL);
  }
  rs_fill_ytab(tab2, P.sc2h, y2lo, r2, P.S1h, yslo, cS);
  if (!ident3) rs_fill_ytab(tab3, P.sc3h, Y0, Y1 - Y0 + 1, P.h, y2lo, c2);
  pdl_wait();
  if (P.up > 1) {
    const int rL = ylhi - yllo + 1;
    for (int x = cx; x < cL; x += 64)
      for (int y = ry; y < rL; y += 4) regL[y * cL + x] = to_f32(mk[(size_t)(yllo + y) * P.W + xllo + x]);
    __syncthreads();
    for (int x = cx; x < cS; x += 64) {                       // kernel_iter_head.py:122-128, then the sigmoid of :446
      const RsTap tx = rs_tap(P.sc1w, xslo + x, P.W, xllo);
      for (int y = ry; y < rS; y += 4) {
        const float v = rs_lerp_tab(regL + tx.i0, tab1[y], tx.step, tx.l0, tx.l1);
        regS[y * cS + x] = 1.0f / (1.0f + expf(-v));
      }
    }
  } else {
    for (int x = cx; x < cS; x += 64)
      for (int y = ry; y < rS; y += 4) {
        const float v = to_f32(mk[(size_t)(yslo + y) * P.W + xslo + x]);
        regS[y * cS + x] = 1.0f / (1.0f + expf(-v));
      }
  }
  __syncthreads();
  pdl_trigger();
  if (ident3) {                                                 // :445-449; the resize to ori_shape is the identity
    for (int x = cx; x < c2; x += 64) {
      const RsTap tx = rs_tap(P.sc2w, x2lo + x, P.S1w, xslo);
      const float *col = regS + tx.i0;
      size_t o = ((size_t)k * P.Ho + Y0 + ry) * Wo + X0 + x;
      for (int y = ry; y < r2; y += 4, o += (size_t)4 * Wo) {
        const float v = rs_lerp_tab(col, tab2[y], tx.step, tx.l0, tx.l1);
        if (probs != nullptr) probs[o] = v;
        if (bits != nullptr) bits[o] = v > thr ? 1 : 0;         // :462
      }
    }
    return;
  }
  for (int x = cx; x < c2; x += 64) {                           // :445-449 (to batch_input_shape; the crop only bounds reads)
    const RsTap tx = rs_tap(P.sc2w, x2lo + x, P.S1w, xslo);
    const float *col = regS + tx.i0;
    for (int y = ry; y < r2; y += 4) reg2[y * c2 + x] = rs_lerp_tab(col, tab2[y], tx.step, tx.l0, tx.l1);
  }
  __syncthreads();
  const int tw = X1 - X0 + 1, th = Y1 - Y0 + 1;
  for (int x = cx; x < tw; x += 64) {                           // :451-457 (to ori_shape)
    const RsTap tx = rs_tap(P.sc3w, X0 + x, P.w, x2lo);
    const float *col = reg2 + tx.i0;
    size_t o = ((size_t)k * P.Ho + Y0 + ry) * Wo + X0 + x;
    for (int y = ry; y < th; y += 4, o += (size_t)4 * Wo) {
      const float v = rs_lerp_tab(col, tab3[y], tx.step, tx.l0, tx.l1);
      if (probs != nullptr) probs[o] = v;
      if (bits != nullptr) bits[o] = v > thr ? 1 : 0;           // :462
    }
  }
}

// host-side bound of the dependency cone: the input extent an output span of `tile` consecutive indices touches
// (src is affine in dst with slope `scale`; + first tap, second tap, fractional alignment)
static int h_extent(float scale, int tile, int in) {
  const int e = (int)(scale * (float)tile) + 3;
  return e < in ? e : in;
}

int launch_rescale_masks(const void *masks, int dtype, int K, int H, int W, int up, int Hb, int Wb, int h, int w, int Ho,
                         int Wo, float thr, float *probs, uint8_t *bits, cudaStream_t stream) {
  if (K < 1 || H < 1 || W < 1 || up < 1 || Hb < 1 || Wb < 1 || h < 1 || w < 1 || Ho < 1 || Wo < 1)
    VKN_FAIL(VKN_E_INVALID, "rescale_masks: sizes must be positive");
  if (h > Hb || w > Wb) VKN_FAIL(VKN_E_INVALID, "rescale_masks: img_shape (%d, %d) exceeds batch_input_shape (%d, %d)", h, w, Hb, Wb);
  if (!probs && !bits) VKN_FAIL(VKN_E_INVALID, "rescale_masks: no output requested");
  RsParams P;
  P.K = K; P.H = H; P.W = W; P.up = up; P.S1h = H * up; P.S1w = W * up; P.Hb = Hb; P.Wb = Wb; P.h = h; P.w = w; P.Ho = Ho; P.Wo = Wo;
  P.sc1h = P.sc1w = 1.0f / (float)up;                 // F.interpolate(scale_factor=up): the given factor is the scale
  P.sc2h = (float)P.S1h / (float)Hb;                  // size-based resizes: input_size / output_size in float
  P.sc2w = (float)P.S1w / (float)Wb;
  P.sc3h = (float)h / (float)Ho;
  P.sc3w = (float)w / (float)Wo;
  P.thr = thr;
  const int ident3 = (h == Ho && w == Wo) ? 1 : 0;     // scale exactly 1: src == dst, lambda == 0
  int ty = RS_TY, tx = RS_TX;
  long long nL = 0, nS = 0, n2 = 0;
  for (;;) {                                          // shrink the output tile until every staged region fits
    const int e2h = ident3 ? ty : h_extent(P.sc3h, ty, h), e2w = ident3 ? tx : h_extent(P.sc3w, tx, w);
    const int eSh = h_extent(P.sc2h, e2h, P.S1h), eSw = h_extent(P.sc2w, e2w, P.S1w);
    const int eLh = h_extent(P.sc1h, eSh, H), eLw = h_extent(P.sc1w, eSw, W);
    nL = up > 1 ? (long long)eLh * eLw : 0;
    nS = (long long)eSh * eSw;
    n2 = ident3 ? 0 : (long long)e2h * e2w;
    if (nL <= RS_MAX_REGION && nS <= RS_MAX_REGION && n2 <= RS_MAX_REGION && e2h <= RS_MAX_ROWS && eSh <= RS_MAX_ROWS &&
        eLh <= RS_MAX_ROWS && ty <= RS_MAX_ROWS)
      break;
    if (ty <= 4 && tx <= 8)
      VKN_FAIL(VKN_E_UNSUPPORTED, "rescale_masks: resize ratios %g / %g make a tile's dependency cone exceed shared memory",
               (double)P.sc3h, (double)P.sc2h);
    if (tx > 2 * ty || ty <= 4) tx /= 2; else ty /= 2;
  }
  P.offS = (int)((nL + 3) & ~3LL);
  P.off2 = P.offS + (int)((nS + 3) & ~3LL);
  P.offT = P.off2 + (int)((n2 + 3) & ~3LL);
  P.ty = ty;
  P.tx = tx;
  const size_t smem = ((size_t)P.offT + 3 * RS_MAX_ROWS * 4 + 4) * sizeof(float);   // sized to the cone: small regions keep occupancy high
  dim3 grid(ceil_div(Wo, tx), ceil_div(Ho, ty), K);
  if (grid.y > 65535 || grid.z > 65535) VKN_FAIL(VKN_E_UNSUPPORTED, "rescale_masks: too many masks / rows for one launch");
  VKN_LAUNCH_MARK("vkn_rescale_masks_kernel", stream);
  if (dtype == VKN_BF16) {
    static unsigned long long a = 0;
    if (first_use_on_device(a)) {
      VKN_CUDA_OK(cudaFuncSetAttribute(vkn_rescale_masks_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((3 * RS_MAX_REGION + 3 * RS_MAX_ROWS * 4 + 16) * sizeof(float))));
    }
    VKN_CUDA_OK(launch_chain(vkn_rescale_masks_kernel<__nv_bfloat16>, grid, dim3(RS_NT), smem, stream, (const __nv_bfloat16 *)masks,
                             probs, bits, P, ident3));
  } else if (dtype == VKN_F32) {
    static unsigned long long a = 0;
    if (first_use_on_device(a)) {
      VKN_CUDA_OK(cudaFuncSetAttribute(vkn_rescale_masks_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((3 * RS_MAX_REGION + 3 * RS_MAX_ROWS * 4 + 16) * sizeof(float))));
    }
    VKN_CUDA_OK(launch_chain(vkn_rescale_masks_kernel<float>, grid, dim3(RS_NT), smem, stream, (const float *)masks, probs, bits, P, ident3));
  } else {
    VKN_FAIL(VKN_E_INVALID, "rescale_masks: bad dtype code");
  }
  return VKN_OK;
}

}  // namespace vkn
