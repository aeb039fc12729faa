// Post-loop result assembly on the device (second half):
//   * vkn_panoptic_merge: the joint score-weighted argmax merge of thing and stuff probability maps into one panoptic id
//     map (VideoKernelIterHead.merge_stuff_thing_stuff_joint, knet/video/kernel_iter_head.py:832-895).  The reference does
//     one full-resolution multiply + argmax and then, per kept kernel, two `.sum().item()` reductions and a masked store --
//     2-3 host synchronisations per segment.  Here: one pass over the maps (argmax owner per pixel + the two areas of
//     every kernel), one single-CTA kernel that walks the kernels in score order and applies the reference's filter rules,
//     one pass that paints the id map.  Integer atomics only: the result is deterministic and bit-identical to the
//     reference's on the same probabilities.
//   * vkn_mask_boxes: the mask -> box reduction of VideoKernelUpdateHead.segm2result (knet/video/kernel_update_head.py:
//     734-744, unitrack tensor_mask2box: extent of the non-zero pixels, (-1,-1,10,10) for an empty mask).
#include "common.cuh"

namespace vkn {

constexpr int PM_NT = 256;
constexpr int PM_MAXT = 1024;      // kernels (things + stuff) per image

// owner[p] = argmax_k scores[k] * masks[k][p]  (first maximum wins, like torch.argmax); won[k] += 1 at its pixels;
// full[k] = #{p : masks[k][p] >= 0.5}.  One thread per pixel walks the T maps (coalesced across the warp).
__global__ void __launch_bounds__(PM_NT) vkn_panoptic_owner_kernel(const float *__restrict__ masks, const float *__restrict__ scores,
                                                                   int T, int HW, int *__restrict__ owner, int *__restrict__ won,
                                                                   int *__restrict__ full) {
  extern __shared__ int sm_cnt[];                      // [2][T] per-CTA counters (shared-memory atomics, then one global add each)
  pdl_wait();
  for (int i = threadIdx.x; i < 2 * T; i += PM_NT) sm_cnt[i] = 0;
  __syncthreads();
  const int p = blockIdx.x * PM_NT + threadIdx.x;
  if (p < HW) {
    float best = 0.f;
    int bk = 0;
    for (int k = 0; k < T; ++k) {
      const float m = __ldg(masks + (size_t)k * HW + p);
      const float v = __fmul_rn(__ldg(scores + k), m);   // the reference multiplies in fp32, then compares
      if (k == 0 || v > best) {
        best = v;
        bk = k;
      }
      if (m >= 0.5f) atomicAdd(&sm_cnt[T + k], 1);
    }
    owner[p] = bk;
    atomicAdd(&sm_cnt[bk], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T; i += PM_NT) {
    if (sm_cnt[i]) atomicAdd(won + i, sm_cnt[i]);
    if (sm_cnt[T + i]) atomicAdd(full + i, sm_cnt[T + i]);
  }
  pdl_trigger();
}

// One CTA: order the kernels by descending score (stable), walk them with the reference's rules (:861-893), emit
// seg_of[k] (0 = dropped), the segment table rows [id, isthing, category_id, instance_id | -1, area | -1], their scores,
// the kept thing indices and the two counts.
__global__ void __launch_bounds__(PM_NT) vkn_panoptic_segments_kernel(const float *__restrict__ scores, const int *__restrict__ labels,
                                                                      const int *__restrict__ won, const int *__restrict__ full,
                                                                      int T, int num_thing, float inst_thr, double overlap_thr,
                                                                      int *__restrict__ seg_of, int *__restrict__ table,
                                                                      float *__restrict__ seg_scores, int *__restrict__ kept,
                                                                      int *__restrict__ counts) {
  __shared__ int order[PM_MAXT];
  pdl_wait();
  for (int k = threadIdx.x; k < T; k += PM_NT) {         // rank of k = #{j : s_j > s_k or (s_j == s_k and j < k)}
    const float sk = scores[k];
    int r = 0;
    for (int j = 0; j < T; ++j) {
      const float sj = scores[j];
      r += (sj > sk) || (sj == sk && j < k);
    }
    order[r] = k;
    seg_of[k] = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int nseg = 0, nkept = 0;
    for (int i = 0; i < T; ++i) {
      const int k = order[i];
      const int cls = labels[k];
      const bool thing = cls < num_thing;
      if (thing && scores[k] < inst_thr) continue;
      const int area = won[k], orig = full[k];
      if (area <= 0 || orig <= 0) continue;
      if ((double)area / (double)orig < overlap_thr) continue;
      ++nseg;
      seg_of[k] = nseg;
      int *row = table + (size_t)(nseg - 1) * 5;
      row[0] = nseg;
      row[1] = thing ? 1 : 0;
      row[2] = thing ? cls : cls - num_thing + 1;
      row[3] = thing ? k : -1;
      row[4] = thing ? -1 : area;
      seg_scores[nseg - 1] = scores[k];
      if (thing) kept[nkept++] = k;
    }
    counts[0] = nseg;
    counts[1] = nkept;
  }
  pdl_trigger();
}

__global__ void __launch_bounds__(PM_NT) vkn_panoptic_paint_kernel(const int *__restrict__ owner, const int *__restrict__ seg_of,
                                                                   int HW, int *__restrict__ seg) {
  pdl_wait();
  const int p = blockIdx.x * PM_NT + threadIdx.x;
  if (p < HW) seg[p] = seg_of[owner[p]];
  pdl_trigger();
}

int launch_panoptic_merge(const float *masks, const float *scores, const int *labels, int T, int H, int W, int num_thing,
                          double inst_thr, double overlap_thr, int *seg, int *table, float *seg_scores, int *kept, int *counts,
                          void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (T < 1 || T > PM_MAXT) VKN_FAIL(VKN_E_UNSUPPORTED, "panoptic_merge: %d kernels (supported: 1..%d)", T, PM_MAXT);
  if (H < 1 || W < 1) VKN_FAIL(VKN_E_INVALID, "panoptic_merge: bad map size");
  const int HW = H * W;
  const size_t need = ((size_t)HW + 3 * (size_t)T) * sizeof(int);
  if (!workspace || workspace_bytes < need)
    VKN_FAIL(VKN_E_WORKSPACE, "panoptic_merge: workspace too small: %zu bytes given, %zu needed", workspace_bytes, need);
  int *owner = (int *)workspace, *won = owner + HW, *full = won + T, *seg_of = full + T;
  VKN_CUDA_OK(cudaMemsetAsync(won, 0, 2 * (size_t)T * sizeof(int), stream));
  const int nb = ceil_div(HW, PM_NT);
  VKN_LAUNCH_MARK("vkn_panoptic_owner_kernel", stream);
  vkn_panoptic_owner_kernel<<<nb, PM_NT, 2 * (size_t)T * sizeof(int), stream>>>(masks, scores, T, HW, owner, won, full);
  VKN_CUDA_OK(cudaGetLastError());
  VKN_LAUNCH_MARK("vkn_panoptic_segments_kernel", stream);
  VKN_CUDA_OK(launch_chain(vkn_panoptic_segments_kernel, dim3(1), dim3(PM_NT), 0, stream, scores, labels, (const int *)won,
                           (const int *)full, T, num_thing, (float)inst_thr, overlap_thr, seg_of, table, seg_scores, kept, counts));
  VKN_LAUNCH_MARK("vkn_panoptic_paint_kernel", stream);
  VKN_CUDA_OK(launch_chain(vkn_panoptic_paint_kernel, dim3(nb), dim3(PM_NT), 0, stream, (const int *)owner, (const int *)seg_of, HW,
                           seg));
  return VKN_OK;
}

// ---- mask -> box ------------------------------------------------------------------------------------------------
// One CTA per mask: extent of its non-zero pixels -> (x_min, y_min, x_max, y_max), or (-1, -1, 10, 10) when empty.
template <typename T>
__global__ void __launch_bounds__(PM_NT) vkn_mask_boxes_kernel(const T *__restrict__ masks, int H, int W, float *__restrict__ boxes) {
  __shared__ int red[4][PM_NT / 32];
  const T *m = masks + (size_t)blockIdx.x * H * W;
  int x0 = 1 << 30, y0 = 1 << 30, x1 = -1, y1 = -1;
  for (int p = threadIdx.x; p < H * W; p += PM_NT) {
    if (m[p] != (T)0) {
      const int y = p / W, x = p - y * W;
      x0 = min(x0, x);
      y0 = min(y0, y);
      x1 = max(x1, x);
      y1 = max(y1, y);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o));
    y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
    x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = x0;
    red[1][warp] = y0;
    red[2][warp] = x1;
    red[3][warp] = y1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < PM_NT / 32; ++w) {
      x0 = min(x0, red[0][w]);
      y0 = min(y0, red[1][w]);
      x1 = max(x1, red[2][w]);
      y1 = max(y1, red[3][w]);
    }
    float *b = boxes + (size_t)blockIdx.x * 4;
    if (x1 < 0) {
      b[0] = -1.f; b[1] = -1.f; b[2] = 10.f; b[3] = 10.f;
    } else {
      b[0] = (float)x0; b[1] = (float)y0; b[2] = (float)x1; b[3] = (float)y1;
    }
  }
}

int launch_mask_boxes(const void *masks, int elem_bytes, int K, int H, int W, float *boxes, cudaStream_t stream) {
  if (K < 0 || H < 1 || W < 1 || !boxes) VKN_FAIL(VKN_E_INVALID, "mask_boxes: bad argument");
  if (K == 0) return VKN_OK;
  VKN_LAUNCH_MARK("vkn_mask_boxes_kernel", stream);
  if (elem_bytes == 1) vkn_mask_boxes_kernel<uint8_t><<<K, PM_NT, 0, stream>>>((const uint8_t *)masks, H, W, boxes);
  else if (elem_bytes == 4) vkn_mask_boxes_kernel<float><<<K, PM_NT, 0, stream>>>((const float *)masks, H, W, boxes);
  else VKN_FAIL(VKN_E_INVALID, "mask_boxes: element size %d (1 = bool / uint8, 4 = float32)", elem_bytes);
  VKN_CUDA_OK(cudaGetLastError());
  return VKN_OK;
}

}  // namespace vkn
