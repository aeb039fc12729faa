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
  const bool valid = p < HW;                              // whole warps walk the maps: the counters are warp-aggregated
  const size_t pp = valid ? (size_t)p : 0;
  const int lane = threadIdx.x & 31;
  float best = 0.f;
  int bk = 0;
#pragma unroll 4
  for (int k = 0; k < T; ++k) {
    const float m = valid ? __ldg(masks + (size_t)k * HW + pp) : 0.f;
    const float v = __fmul_rn(__ldg(scores + k), m);     // the reference multiplies in fp32, then compares
    if (k == 0 || v > best) {
      best = v;
      bk = k;
    }
    const unsigned hot = __ballot_sync(0xffffffffu, valid && m >= 0.5f);     // one shared-memory atomic per warp and map
    if (lane == 0 && hot) atomicAdd(&sm_cnt[T + k], __popc(hot));
  }
  if (valid) owner[p] = bk;
  {
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const unsigned same = __match_any_sync(act, bk);   // neighbouring pixels mostly share their owner
      if (lane == __ffs(same) - 1) atomicAdd(&sm_cnt[bk], __popc(same));
    }
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
// The mask is walked as a flat array in 16-byte groups (scalar head / tail up to the alignment of this mask's first byte): an
// all-zero group -- most of a mask -- costs one load and one test; a group with a set pixel pays ONE division for its (row,
// column) and walks its elements from there.  (The first version loaded element by element and divided per set pixel: 237 us for
// 100 masks of 375 x 1242; this one is bound by the 46 MB it reads.)
template <typename T>
__device__ __forceinline__ void mb_visit(const T *m, int p, int W, int &x0, int &y0, int &x1, int &y1) {
  if (m[p] != (T)0) {
    const int y = p / W, x = p - y * W;
    x0 = min(x0, x);
    y0 = min(y0, y);
    x1 = max(x1, x);
    y1 = max(y1, y);
  }
}
template <typename T>
__global__ void __launch_bounds__(PM_NT) vkn_mask_boxes_kernel(const T *__restrict__ masks, int H, int W, float *__restrict__ boxes) {
  __shared__ int red[4][PM_NT / 32];
  constexpr int EPG = 16 / (int)sizeof(T);                  // elements per 16-byte group
  const int HW = H * W;
  const T *m = masks + (size_t)blockIdx.x * HW;
  int x0 = 1 << 30, y0 = 1 << 30, x1 = -1, y1 = -1;
  int head = (int)(((16 - (reinterpret_cast<uintptr_t>(m) & 15)) & 15) / sizeof(T));
  if (head > HW) head = HW;
  const int ngroups = (HW - head) / EPG;
  const int tail0 = head + ngroups * EPG;
  for (int p = threadIdx.x; p < head; p += PM_NT) mb_visit(m, p, W, x0, y0, x1, y1);
  for (int p = tail0 + threadIdx.x; p < HW; p += PM_NT) mb_visit(m, p, W, x0, y0, x1, y1);
  const uint4 *g4 = reinterpret_cast<const uint4 *>(m + head);
#pragma unroll 4
  for (int g = threadIdx.x; g < ngroups; g += PM_NT) {
    const uint4 v = __ldg(g4 + g);
    if ((v.x | v.y | v.z | v.w) == 0u) continue;            // (-0.0f would pass this test and is then rejected element-wise)
    const int p0 = head + g * EPG;
    int y = p0 / W, x = p0 - y * W;
#pragma unroll
    for (int e = 0; e < EPG; ++e) {
      if (m[p0 + e] != (T)0) {
        x0 = min(x0, x);
        y0 = min(y0, y);
        x1 = max(x1, x);
        y1 = max(y1, y);
      }
      if (++x == W) {
        x = 0;
        ++y;
      }
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


// ---- tracking association --------------------------------------------------------------------
// QuasiDenseEmbedTracker.match (knet/video/qdtrack/trackers/quasi_dense_embed_tracker.py:137-207) up to the memory update:
// sort the detections by score, drop duplicates by box IoU (:145-154), bi-directional softmax of the embedding
// similarities against the memory (:166-170), same-category mask (:182-184), the greedy assignment with column
// knock-out (:186-198) and the new-track ids (:199-204).  The reference does this with ~4 host synchronisations per
// detection; here it is ONE single-CTA launch (tens of detections x a few hundred memory entries).
constexpr int TM_NT = 256;
constexpr int TM_MAXN = 256;       // detections per frame
constexpr int TM_SMEM_SCORES = 8192;   // floats of the score matrix staged in shared memory for the serial greedy walk

__device__ __forceinline__ float tm_iou(const float *a, const float *b) {      // mmdet bbox_overlaps(mode='iou', eps=1e-6)
  const float ix = fmaxf(fminf(a[2], b[2]) - fmaxf(a[0], b[0]), 0.f), iy = fmaxf(fminf(a[3], b[3]) - fmaxf(a[1], b[1]), 0.f);
  const float ov = ix * iy;
  const float uni = fmaxf((a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - ov, 1e-6f);
  return ov / uni;
}

__global__ void __launch_bounds__(TM_NT) vkn_track_match_kernel(
    const float *__restrict__ bboxes, const long long *__restrict__ labels, const float *__restrict__ embeds, int n, int D,
    const long long *__restrict__ memo_labels, const float *__restrict__ memo_embeds, const long long *__restrict__ memo_ids, int m,
    float obj_score_thr, float match_score_thr, float init_score_thr, float nms_conf_thr, float nms_backdrop_iou_thr,
    float nms_class_iou_thr, int with_cats, long long num_tracklets, int *__restrict__ sel, long long *__restrict__ ids,
    int *__restrict__ counts, float *__restrict__ scores /* workspace [n, m] */) {
  __shared__ int order[TM_MAXN], keep[TM_MAXN], nkeep_s;
  __shared__ float sc_s[TM_SMEM_SCORES];          // the score matrix of the greedy walk when it fits (else it stays in global memory)
  __shared__ long long mid_s[32 * 32];            // memo_ids of the single-warp walk
  __shared__ float det_s[TM_MAXN];                // detection scores in walk order
  __shared__ float red_v[TM_NT / 32];
  __shared__ int red_i[TM_NT / 32], pick_s;
  __shared__ float conf_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 1. order by descending score (stable)
  for (int i = tid; i < n; i += TM_NT) {
    const float si = bboxes[i * 5 + 4];
    int r = 0;
    for (int j = 0; j < n; ++j) {
      const float sj = bboxes[j * 5 + 4];
      r += (sj > si) || (sj == si && j < i);
    }
    order[r] = i;
  }
  __syncthreads();
  // 2. duplicate removal: detection i (sorted) is dropped when it overlaps ANY better-scored detection too much
  for (int i = tid; i < n; i += TM_NT) {
    const float *bi = bboxes + order[i] * 5;
    const float thr = bi[4] < obj_score_thr ? nms_backdrop_iou_thr : nms_class_iou_thr;
    int ok = 1;
    for (int j = 0; j < i; ++j)
      if (tm_iou(bi, bboxes + order[j] * 5) > thr) {
        ok = 0;
        break;
      }
    keep[i] = ok;
  }
  __syncthreads();
  if (tid == 0) {
    int k = 0;
    for (int i = 0; i < n; ++i)
      if (keep[i]) sel[k++] = order[i];
    nkeep_s = k;
  }
  __syncthreads();
  const int nk = nkeep_s;
  for (int i = tid; i < nk; i += TM_NT) ids[i] = -1;
  if (nk > 0 && m > 0) {
    // 3. similarities, bi-directional softmax, category mask
    // a warp per detection, four memory entries at a time (their loads are in flight together: the loop is latency-bound):
    // coalesced loads, lane-strided partial sums, fixed shuffle tree
    for (int i = warp; i < nk; i += TM_NT / 32) {
      const float *a = embeds + (size_t)sel[i] * D;
      for (int j0 = 0; j0 < m; j0 += 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const float *b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) b[u] = memo_embeds + (size_t)min(j0 + u, m - 1) * D;
#pragma unroll 4
        for (int d = lane; d < D; d += 32) {
          const float av = __ldg(a + d);
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] = fmaf(av, __ldg(b[u] + d), acc[u]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
        if (lane == 0) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (j0 + u < m) scores[i * m + j0 + u] = acc[u];
        }
      }
    }
    __syncthreads();
    // row softmax -> keep exp and sums in place via two passes; column softmax needs the raw feats: use a second buffer
    float *feats = scores + (size_t)nk * m;                 // workspace holds 2 x [n, m]
    for (int e = tid; e < nk * m; e += TM_NT) feats[e] = scores[e];
    __syncthreads();
    // d2t: softmax over the memory entries -- a warp per detection (lanes over the entries, shuffle reductions)
    for (int i = warp; i < nk; i += TM_NT / 32) {
      float mx = -3.0e38f, sm = 0.f;
      for (int j = lane; j < m; j += 32) mx = fmaxf(mx, feats[i * m + j]);
      mx = warp_max(mx);
      for (int j = lane; j < m; j += 32) sm += expf(feats[i * m + j] - mx);
      sm = warp_sum(sm);
      for (int j = lane; j < m; j += 32) scores[i * m + j] = expf(feats[i * m + j] - mx) / sm;
    }
    __syncthreads();
    // t2d: softmax over the detections; average of the two; category mask -- a warp per memory entry (lanes over detections)
    for (int j = warp; j < m; j += TM_NT / 32) {
      float mx = -3.0e38f, sm = 0.f;
      for (int i = lane; i < nk; i += 32) mx = fmaxf(mx, feats[i * m + j]);
      mx = warp_max(mx);
      for (int i = lane; i < nk; i += 32) sm += expf(feats[i * m + j] - mx);
      sm = warp_sum(sm);
      const long long ml = with_cats ? memo_labels[j] : 0;
      for (int i = lane; i < nk; i += 32) {
        float v = (scores[i * m + j] + expf(feats[i * m + j] - mx) / sm) / 2.f;
        if (with_cats && labels[sel[i]] != ml) v *= 0.f;
        scores[i * m + j] = v;
      }
    }
    __syncthreads();
    // 4. greedy assignment in score order, a matched memory column is knocked out (its score reads as 0, as the reference
    //    writes it) for every other detection.  Inherently serial over the detections: ONE warp walks them with shuffles only
    //    (lane l owns the columns l, l + 32, ...; its knocked-out columns are bits of a register) -- no block barrier per step.
    // (the walk is a chain of dependent loads: everything it touches is staged in shared memory first)
    const bool in_smem = nk * m <= TM_SMEM_SCORES;
    if (m <= 32 * 32) {
      if (in_smem)
        for (int e = tid; e < nk * m; e += TM_NT) sc_s[e] = scores[e];
      for (int j = tid; j < m; j += TM_NT) mid_s[j] = memo_ids[j];
      for (int i = tid; i < nk; i += TM_NT) det_s[i] = bboxes[sel[i] * 5 + 4];
      __syncthreads();
    }
    const float *sc = in_smem ? sc_s : scores;
    if (warp == 0 && m <= 32 * 32) {
      uint32_t knocked = 0u;                                  // bit c: column lane + 32 c is taken
      for (int i = 0; i < nk; ++i) {
        float bv = -3.0e38f;
        int bj = 0x7fffffff;
        for (int j = lane, cidx = 0; j < m; j += 32, ++cidx) {
          const float v = ((knocked >> cidx) & 1u) ? 0.f : sc[i * m + j];
          if (v > bv) {
            bv = v;
            bj = j;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
          if (ov > bv || (ov == bv && oj < bj)) {
            bv = ov;
            bj = oj;
          }
        }
        int knock = -1;                                       // every lane evaluates the same decision
        if (bv > match_score_thr) {
          const long long id = mid_s[bj];
          if (id > -1) {
            if (det_s[i] > obj_score_thr) {
              if (lane == 0) ids[i] = id;
              knock = bj;
            } else if (bv > nms_conf_thr) {
              if (lane == 0) ids[i] = -2;
            }
          }
        }
        if (knock >= 0 && (knock & 31) == lane) knocked |= 1u << (knock >> 5);
      }
    } else if (m > 32 * 32) {
    for (int i = 0; i < nk; ++i) {
      float bv = -3.0e38f;
      int bj = 0x7fffffff;
      for (int j = tid; j < m; j += TM_NT) {
        const float v = scores[i * m + j];
        if (v > bv) {
          bv = v;
          bj = j;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (ov > bv || (ov == bv && oj < bj)) {
          bv = ov;
          bj = oj;
        }
      }
      if (lane == 0) {
        red_v[warp] = bv;
        red_i[warp] = bj;
      }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < TM_NT / 32; ++w)
          if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bj)) {
            bv = red_v[w];
            bj = red_i[w];
          }
        int knock = -1;
        if (bv > match_score_thr) {
          const long long id = memo_ids[bj];
          if (id > -1) {
            if (bboxes[sel[i] * 5 + 4] > obj_score_thr) {
              ids[i] = id;
              knock = bj;
            } else if (bv > nms_conf_thr) {
              ids[i] = -2;
            }
          }
        }
        pick_s = knock;
        conf_s = bv;
      }
      __syncthreads();
      const int knock = pick_s;
      if (knock >= 0)
        for (int r = tid; r < nk; r += TM_NT)
          if (r != i) scores[r * m + knock] = 0.f;
      __syncthreads();
    }
    }
    __syncthreads();
  }
  // 5. new tracks: unmatched detections above init_score_thr, numbered in score order
  if (tid == 0) {
    long long next = num_tracklets;
    for (int i = 0; i < nk; ++i)
      if (ids[i] == -1 && bboxes[sel[i] * 5 + 4] > init_score_thr) ids[i] = next++;
    counts[0] = nk;
    counts[1] = (int)(next - num_tracklets);
  }
}

int launch_track_match(const float *bboxes, const long long *labels, const float *embeds, int n, int D, const long long *memo_labels,
                       const float *memo_embeds, const long long *memo_ids, int m, const float *thr6, int with_cats,
                       long long num_tracklets, int *sel, long long *ids, int *counts, void *workspace, size_t workspace_bytes,
                       cudaStream_t stream) {
  if (n < 0 || n > TM_MAXN || m < 0 || D < 1) VKN_FAIL(VKN_E_UNSUPPORTED, "track_match: n = %d detections (supported: 0..%d)", n, TM_MAXN);
  size_t need = (size_t)2 * n * m * sizeof(float);
  if (need < 8) need = 8;
  if (!workspace || workspace_bytes < need) VKN_FAIL(VKN_E_WORKSPACE, "track_match: workspace too small: %zu given, %zu needed", workspace_bytes, need);
  VKN_LAUNCH_MARK("vkn_track_match_kernel", stream);
  vkn_track_match_kernel<<<1, TM_NT, 0, stream>>>(bboxes, labels, embeds, n, D, memo_labels, memo_embeds, memo_ids, m, thr6[0], thr6[1],
                                                  thr6[2], thr6[3], thr6[4], thr6[5], with_cats, num_tracklets, sel, ids, counts,
                                                  (float *)workspace);
  VKN_CUDA_OK(cudaGetLastError());
  return VKN_OK;
}

}  // namespace vkn
