// Training-side cost matrix of MaskHungarianAssigner: for one image, the N predicted masks against the M
// ground-truth masks,
//
//   cost[n,m] = w_cls  * FocalLossCost(cls_logits)[n, label_m]                                   (mmdet match_cost.py, v2.18)
//             + w_mask * -( sum_p pm[n,p] t[m,p] + sum_p (1 - pm[n,p]) (1 - t[m,p]) ) / HW       (MaskCost,  mask_hungarian_assigner.py:93-110)
//             + w_dice * -( 2 sum_p pd[n,p] t[m,p] ) / ( sum_p pd[n,p]^2 + eps + sum_p t[m,p]^2 + eps )   (DiceCost, :43-75)
//
// with pm = clamp(sigmoid(logit), 0.01, 1), pd = clamp(sigmoid(logit), 0.001, 1) (pred_act=True, act_mode='sigmoid': what every
// shipped config sets).  The reference materialises both activations, 1 - pm, 1 - t and runs three einsums over the full masks;
// here ONE pass reads the logits and the targets once: both [N x HW] . [HW x M] contractions share the staged tiles, the
// negative term follows from row / column sums (sum (1-pm)(1-t) = HW - sum pm - sum t + sum pm t), pixel chunks are reduced in
// a fixed order (deterministic), and a small second kernel assembles the three costs.  The Hungarian solve that consumes the
// matrix stays the reference's (scipy, on the host).
#include "common.cuh"

namespace vkn {

constexpr int MC_NT = 256;
constexpr int MC_PX = 64;            // pixels per staged block
constexpr int MC_LD = MC_PX + 4;     // row stride (floats): 16-byte aligned rows, conflict-free float4 reads at row stride 1
constexpr int MC_NB = 128;           // predictions per CTA (4 per thread: n = tn + 32 i)
constexpr int MC_MB = 32;            // targets per CTA     (4 per thread: m = tm + 8 j)

__global__ void __launch_bounds__(MC_NT, 2) vkn_match_cost_partial_kernel(const float *__restrict__ logits, const float *__restrict__ gt,
                                                                      int N, int M, int HW, int blocks_per_chunk,
                                                                      float *__restrict__ p_dm, float *__restrict__ p_row,
                                                                      float *__restrict__ p_col) {
  extern __shared__ __align__(16) float mc_smem[];
  float *pd = mc_smem;                       // [128][MC_LD]  clamp(sigmoid, 0.001, 1)
  float *pm = pd + MC_NB * MC_LD;            // [128][MC_LD]  clamp(sigmoid, 0.01, 1)
  float *tt = pm + MC_NB * MC_LD;            // [32][MC_LD]   targets
  const int tid = threadIdx.x, tn = tid >> 3, tm = tid & 7;
  const int chunk = blockIdx.x, m0 = blockIdx.y * MC_MB, n0 = blockIdx.z * MC_NB;
  const int nchunks = gridDim.x;
  float acc_d[4][4], acc_m[4][4], rs_d[4], rs_m[4], cs_2[4], cs_1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    rs_d[i] = rs_m[i] = cs_2[i] = cs_1[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc_d[i][j] = acc_m[i][j] = 0.f;
  }
  const int blk0 = chunk * blocks_per_chunk;
  for (int b = 0; b < blocks_per_chunk; ++b) {
    const int p0 = (blk0 + b) * MC_PX;
    if (p0 >= HW) break;
    __syncthreads();
    // stage: activations of 128 predictions x 64 pixels, 32 targets x 64 pixels (zero beyond N / M / HW: no contribution)
    for (int idx = tid; idx < MC_NB * MC_PX; idx += MC_NT) {
      const int r = idx >> 6, px = idx & 63, n = n0 + r, p = p0 + px;
      float d = 0.f, m = 0.f;
      if (n < N && p < HW) {
        const float s = sigmoid_fast(__ldg(logits + (size_t)n * HW + p));      // ex2.approx + rcp.approx: ~1e-7 relative
        d = fminf(fmaxf(s, 0.001f), 1.0f);
        m = fminf(fmaxf(s, 0.01f), 1.0f);
      }
      pd[r * MC_LD + px] = d;
      pm[r * MC_LD + px] = m;
    }
    for (int idx = tid; idx < MC_MB * MC_PX; idx += MC_NT) {
      const int r = idx >> 6, px = idx & 63, m = m0 + r, p = p0 + px;
      tt[r * MC_LD + px] = (m < M && p < HW) ? __ldg(gt + (size_t)m * HW + p) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int px = 0; px < MC_PX; px += 4) {
      float4 a[4], c[4], t[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = *reinterpret_cast<const float4 *>(pd + (tn + 32 * i) * MC_LD + px);
        c[i] = *reinterpret_cast<const float4 *>(pm + (tn + 32 * i) * MC_LD + px);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) t[j] = *reinterpret_cast<const float4 *>(tt + (tm + 8 * j) * MC_LD + px);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc_d[i][j] = fmaf(a[i].x, t[j].x, acc_d[i][j]);
          acc_d[i][j] = fmaf(a[i].y, t[j].y, acc_d[i][j]);
          acc_d[i][j] = fmaf(a[i].z, t[j].z, acc_d[i][j]);
          acc_d[i][j] = fmaf(a[i].w, t[j].w, acc_d[i][j]);
          acc_m[i][j] = fmaf(c[i].x, t[j].x, acc_m[i][j]);
          acc_m[i][j] = fmaf(c[i].y, t[j].y, acc_m[i][j]);
          acc_m[i][j] = fmaf(c[i].z, t[j].z, acc_m[i][j]);
          acc_m[i][j] = fmaf(c[i].w, t[j].w, acc_m[i][j]);
        }
      if (tm == 0) {           // row sums of this thread's 4 predictions (one thread per row group)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          rs_d[i] += (a[i].x * a[i].x + a[i].y * a[i].y) + (a[i].z * a[i].z + a[i].w * a[i].w);
          rs_m[i] += (c[i].x + c[i].y) + (c[i].z + c[i].w);
        }
      }
      if (tn == 0) {           // column sums of this thread's 4 targets
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          cs_2[j] += (t[j].x * t[j].x + t[j].y * t[j].y) + (t[j].z * t[j].z + t[j].w * t[j].w);
          cs_1[j] += (t[j].x + t[j].y) + (t[j].z + t[j].w);
        }
      }
    }
  }
  // partials of this pixel chunk (every CTA writes all of its entries, also the all-zero ones past HW)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + tn + 32 * i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + tm + 8 * j;
      if (m < M) *reinterpret_cast<float2 *>(p_dm + (((size_t)chunk * N + n) * M + m) * 2) = make_float2(acc_d[i][j], acc_m[i][j]);
    }
    if (tm == 0 && blockIdx.y == 0) *reinterpret_cast<float2 *>(p_row + ((size_t)chunk * N + n) * 2) = make_float2(rs_d[i], rs_m[i]);
  }
  if (tn == 0 && blockIdx.z == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + tm + 8 * j;
      if (m < M) *reinterpret_cast<float2 *>(p_col + ((size_t)chunk * M + m) * 2) = make_float2(cs_2[j], cs_1[j]);
    }
  }
  (void)nchunks;
}

struct McParams {
  float w_cls, w_mask, w_dice, dice_eps, alpha, gamma, focal_eps;
};

__global__ void __launch_bounds__(256) vkn_match_cost_final_kernel(const float *__restrict__ p_dm, const float *__restrict__ p_row,
                                                                   const float *__restrict__ p_col, int nchunks,
                                                                   const float *__restrict__ cls_logits,
                                                                   const long long *__restrict__ gt_labels, int N, int M, int HW,
                                                                   int ncls, McParams P, float *__restrict__ cost) {
  // one warp per output: lane l sums the chunks l, l + 32, ... (fixed order), then a fixed shuffle tree -> deterministic
  const int idx = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (idx >= N * M) return;
  const int n = idx / M, m = idx - n * M;
  float a = 0.f, pt = 0.f, b = 0.f, spm = 0.f, c = 0.f, st = 0.f;
  for (int ch = lane; ch < nchunks; ch += 32) {
    const float2 dm = __ldg(reinterpret_cast<const float2 *>(p_dm + (((size_t)ch * N + n) * M + m) * 2));
    const float2 r = __ldg(reinterpret_cast<const float2 *>(p_row + ((size_t)ch * N + n) * 2));
    const float2 cc = __ldg(reinterpret_cast<const float2 *>(p_col + ((size_t)ch * M + m) * 2));
    a += dm.x;
    pt += dm.y;
    b += r.x;
    spm += r.y;
    c += cc.x;
    st += cc.y;
  }
  a = warp_sum(a);
  pt = warp_sum(pt);
  b = warp_sum(b);
  spm = warp_sum(spm);
  c = warp_sum(c);
  st = warp_sum(st);
  if (lane != 0) return;
  float total = 0.f;
  if (P.w_cls != 0.f && cls_logits != nullptr) {       // mmdet FocalLossCost: pos_cost[:, label] - neg_cost[:, label]
    const long long lab = gt_labels[m];
    float v = 0.f;
    if (lab >= 0 && lab < ncls) {
      const float p = 1.0f / (1.0f + expf(-__ldg(cls_logits + (size_t)n * ncls + lab)));
      const float neg = -logf(1.0f - p + P.focal_eps) * (1.0f - P.alpha) * powf(p, P.gamma);
      const float pos = -logf(p + P.focal_eps) * P.alpha * powf(1.0f - p, P.gamma);
      v = pos - neg;
    }
    total = v * P.w_cls;
  }
  if (P.w_mask != 0.f) {
    const float neg = (((float)HW - spm) - st) + pt;
    total += (-(pt + neg) / (float)HW) * P.w_mask;
  }
  if (P.w_dice != 0.f) total += (-(2.0f * a) / ((b + P.dice_eps) + (c + P.dice_eps))) * P.w_dice;
  cost[idx] = total;
}

static int mc_chunks(int N, int M, int HW, int *bpc) {
  const int nblk = ceil_div(HW, MC_PX);
  const int tiles = ceil_div(M, MC_MB) * ceil_div(N, MC_NB);
  int chunks = 296 / tiles;                 // two CTAs (78 KB of shared memory each) per SM
  if (chunks < 1) chunks = 1;
  if (chunks > nblk) chunks = nblk;
  *bpc = ceil_div(nblk, chunks);
  return ceil_div(nblk, *bpc);
}

size_t match_cost_workspace_bytes(int N, int M, int HW) {
  int bpc;
  const int ch = mc_chunks(N, M, HW, &bpc);
  return ((size_t)ch * N * M * 2 + (size_t)ch * N * 2 + (size_t)ch * M * 2) * sizeof(float) + 256;
}

int launch_match_cost(const float *mask_logits, const float *cls_logits, const float *gt_masks, const long long *gt_labels, int N,
                      int M, int HW, int ncls, const float *params7, float *cost, void *workspace, size_t workspace_bytes,
                      cudaStream_t stream) {
  if (N < 1 || M < 1 || HW < 1) VKN_FAIL(VKN_E_INVALID, "vkn_match_cost: N, M, HW must be positive (the caller handles empty sets)");
  if (!mask_logits || !gt_masks || !cost || !params7) VKN_FAIL(VKN_E_INVALID, "vkn_match_cost: null argument");
  McParams P = {params7[0], params7[1], params7[2], params7[3], params7[4], params7[5], params7[6]};
  if (P.w_cls != 0.f && cls_logits != nullptr && (!gt_labels || ncls < 1)) VKN_FAIL(VKN_E_INVALID, "vkn_match_cost: cls cost needs labels");
  if (workspace_bytes < match_cost_workspace_bytes(N, M, HW) || !workspace) VKN_FAIL(VKN_E_WORKSPACE, "vkn_match_cost: workspace too small");
  int bpc;
  const int ch = mc_chunks(N, M, HW, &bpc);
  float *p_dm = (float *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  float *p_row = p_dm + (size_t)ch * N * M * 2;
  float *p_col = p_row + (size_t)ch * N * 2;
  const size_t smem = (size_t)(2 * MC_NB + MC_MB) * MC_LD * sizeof(float);
  static unsigned long long attr_mask = 0;
  if (first_use_on_device(attr_mask))
    VKN_CUDA_OK(cudaFuncSetAttribute(vkn_match_cost_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VKN_LAUNCH_MARK("vkn_match_cost_partial_kernel", stream);
  vkn_match_cost_partial_kernel<<<dim3(ch, ceil_div(M, MC_MB), ceil_div(N, MC_NB)), MC_NT, smem, stream>>>(mask_logits, gt_masks, N, M, HW,
                                                                                                         bpc, p_dm, p_row, p_col);
  VKN_LAUNCH_MARK("vkn_match_cost_final_kernel", stream);
  vkn_match_cost_final_kernel<<<ceil_div(N * M, 8), 256, 0, stream>>>(p_dm, p_row, p_col, ch, cls_logits, gt_labels, N, M, HW, ncls, P,
                                                                       cost);
  VKN_CUDA_OK(cudaGetLastError());
  return VKN_OK;
}

}  // namespace vkn
