// tcgen05 attention among the N kernels of a frame (mmcv MultiheadAttention / torch F.multi_head_attention_forward core,
// knet/det/kernel_update_head.py:204-206): softmax((q / sqrt(d)) k^T) v per (frame, head), d = 32, N <= 128.
//
// The fp32 SIMT kernel is bound by its shared-memory loads (68 % LSU, 58 us per 64 frames).  Here one CTA owns a frame and
// walks its heads two at a time (a pair of heads = 64 channels = one 128-byte swizzle row):
//   stage    q (pre-scaled), k, v of the pair: fp32 rows from global -> two fp16 planes (hi + lo = 22 significant bits;
//            all operands are O(1), far inside the fp16 range) written straight into the UMMA K-major SWIZZLE_128B layout;
//   S = QK^T tcgen05.mma kind::f16, M = 128 queries, N = keys (multiple of 16), K = 32: hi.hi + hi.lo + lo.hi -> TMEM fp32;
//   softmax  thread = query row: max / exp / sum over the valid keys straight from TMEM, P = exp(s - max) as two fp16
//            planes into shared memory (the A operand of the next product);
//   O = PV   M = 128, N = 64 (both heads' channels of the pair: V is used as it lies, MN-major; the other head's half of
//            the accumulator is ignored), K = keys: hi.hi + hi.lo + lo.hi;
//   out      O / rowsum -> the three bf16 planes the out-projection GEMM consumes (and / or fp32 rows).
// Every product is exact in the fp32 accumulators; what is dropped (lo.lo) is below 2^-22 relative.
#include <cuda_fp16.h>

#include "tc.cuh"

namespace vkn {

constexpr int AT_THREADS = 288;          // warp 0: MMA issue; warps 1-4 / 5-8: softmax + epilogue of the pair's first / second head; all: staging
constexpr uint32_t AT_TILE = 128u * 128u;             // one [128 rows x 64 fp16] plane tile: 16 KB
// Q, K, V: hi | lo (2 tiles each); P of head 0 / head 1 of the pair: [2 key chunks of 64][hi | lo] (4 tiles each)
constexpr uint32_t AT_Q = 0, AT_K = 2 * AT_TILE, AT_V = 4 * AT_TILE, AT_P = 6 * AT_TILE;
constexpr uint32_t AT_SMEM = 14 * AT_TILE;            // 224 KB

__device__ __forceinline__ uint32_t at_swz(int row, int chunk16) {       // byte offset inside a [rows x 128 B] K-major SW128 tile
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk16 ^ (row & 7)) << 4));
}
// 8 fp32 -> hi / lo fp16 words (4 x half2 each)
__device__ __forceinline__ void at_split8(const float (&v)[8], uint4 &hi, uint4 &lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(v[2 * i] - back.x, v[2 * i + 1] - back.y);
    h[i] = *reinterpret_cast<const uint32_t *>(&hh);
    l[i] = *reinterpret_cast<const uint32_t *>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(AT_THREADS, 1)
vkn_attention_tc_kernel(const float *__restrict__ q, int ldq, const float *__restrict__ k, int ldk, const float *__restrict__ v,
                        int ldv, float *__restrict__ out, int ldo, __nv_bfloat16 *__restrict__ planes, long long plane_stride,
                        int ldp, int N, int C, float scale, uint32_t idesc_qk, uint32_t idesc_pv) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = (uint64_t *)(smem + AT_SMEM);           // [0,1] S ready (head 0 / 1 of the pair), [2,3] O ready
  uint32_t *tmem_slot = (uint32_t *)(bars + 4);
  const uint32_t smem0 = smem_u32(smem), bar0 = smem_u32(bars);
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31, tid = threadIdx.x;
  const int b = blockIdx.x;
  const size_t row0 = (size_t)b * N;
  const int Nk = (N + 15) & ~15;                            // keys per MMA (multiple of 16; rows >= N are zero)
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bar0 + 8 * i, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;                           // S of head hh: columns [128 hh, +128); O: [256 + 64 hh, +64)
  pdl_wait();
  uint32_t ph = 0;
  const int npairs = C / 64;
  for (int pr = 0; pr < npairs; ++pr) {
    // ---- stage q, k, v of this head pair: rows [0, 128) x 64 channels, zero beyond the frame's N rows.
    //      (item = one row x 8 channels of q, k or v; several items per thread are in flight before the first is converted:
    //       the loop is bound by the L2 latency of these loads, not by their bytes)
    constexpr int AT_BATCH = 4;
    for (int it0 = tid; it0 < 3 * 128 * 8; it0 += AT_THREADS * AT_BATCH) {
      float4 ld[AT_BATCH][2];
#pragma unroll
      for (int u = 0; u < AT_BATCH; ++u) {
        const int it = it0 + u * AT_THREADS;
        const int which = it >> 10, rem = it & 1023, r = rem >> 3, g = rem & 7;
        ld[u][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        ld[u][1] = ld[u][0];
        if (it < 3 * 128 * 8 && r < N) {
          const float *src = (which == 0 ? q + (row0 + r) * ldq : (which == 1 ? k + (row0 + r) * ldk : v + (row0 + r) * ldv)) + pr * 64 + g * 8;
          ld[u][0] = __ldcg(reinterpret_cast<const float4 *>(src));
          ld[u][1] = __ldcg(reinterpret_cast<const float4 *>(src + 4));
        }
      }
#pragma unroll
      for (int u = 0; u < AT_BATCH; ++u) {
        const int it = it0 + u * AT_THREADS;
        if (it >= 3 * 128 * 8) break;
        const int which = it >> 10, rem = it & 1023, r = rem >> 3, g = rem & 7;
        const float sc = which == 0 ? scale : 1.0f;          // torch scales q before the product
        const float val[8] = {ld[u][0].x * sc, ld[u][0].y * sc, ld[u][0].z * sc, ld[u][0].w * sc,
                              ld[u][1].x * sc, ld[u][1].y * sc, ld[u][1].z * sc, ld[u][1].w * sc};
        uint4 hi, lo;
        at_split8(val, hi, lo);
        const uint32_t base = smem0 + (uint32_t)which * 2u * AT_TILE + at_swz(r, g);
        sts_v4(base, hi.x, hi.y, hi.z, hi.w);
        sts_v4(base + AT_TILE, lo.x, lo.y, lo.z, lo.w);
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      // ---- MMA issue: S_h = Q_h K_h^T for both heads, then O_h = P_h V as soon as each head's P is complete ----
      tc_fence_after();
      const uint64_t lo16 = (uint64_t)(AT_TILE >> 4);
      if (elect_one()) {
        const uint64_t qd = umma_desc_sw128(smem0 + AT_Q, 0, 1024), kd = umma_desc_sw128(smem0 + AT_K, 0, 1024);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {                      // lo.hi, hi.lo, hi.hi (small terms first)
            const uint64_t qa = qd + (c == 0 ? lo16 : 0), kb = kd + (c == 1 ? lo16 : 0);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma_bf16(tm + (uint32_t)(128 * hh), qa + (uint64_t)((2 * hh + ks) * 2), kb + (uint64_t)((2 * hh + ks) * 2), idesc_qk,
                        (c > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(bar0 + 8 * hh);
        }
      }
      __syncwarp();
      for (int hh = 0; hh < 2; ++hh) {
        named_bar_sync(2 + hh, 160);                         // P of head hh complete (its four row quarters)
        tc_fence_after();
        if (elect_one()) {
          const uint64_t pd = umma_desc_sw128(smem0 + AT_P + (uint32_t)hh * 4u * AT_TILE, 0, 1024);
          const uint64_t vd = umma_desc_sw128(smem0 + AT_V, 8192, 1024);
          const int nks = Nk >> 4;
          for (int c = 0; c < 3; ++c) {                      // lo.hi, hi.lo, hi.hi
            const uint64_t pa = pd + (c == 0 ? lo16 : 0), vb = vd + (c == 1 ? lo16 : 0);
            for (int ks = 0; ks < nks; ++ks)
              umma_bf16(tm + 256u + (uint32_t)(64 * hh), pa + (uint64_t)((ks >> 2) * (2 * (int)(AT_TILE >> 4)) + (ks & 3) * 2),
                        vb + (uint64_t)(ks * (2048 >> 4)), idesc_pv, (c > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(bar0 + 8 * (2 + hh));
        }
        __syncwarp();
      }
    } else {
      // ---- softmax over the valid keys, thread = query row; warps 1-4: first head of the pair, warps 5-8: second ----
      const int hh = (warp - 1) >> 2, qd_ = warp & 3, r = qd_ * 32 + lane;
      mbar_wait(bar0 + 8 * hh, ph);
      tc_fence_after();
      const uint32_t ts = tm + (uint32_t)(128 * hh) + ((uint32_t)(qd_ * 32) << 16);
      float mx = -3.0e38f;
      for (int c0 = 0; c0 < Nk; c0 += 32) {
        uint32_t sr[32];
        tmem_ld32(ts + c0, sr);
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (c0 + e < N) mx = fmaxf(mx, __uint_as_float(sr[e]));
      }
      float sum = 0.f;
      for (int c0 = 0; c0 < 128; c0 += 32) {
        float p[32];
        if (c0 < Nk) {
          uint32_t sr[32];
          tmem_ld32(ts + c0, sr);
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            p[e] = (c0 + e < N) ? __expf(__uint_as_float(sr[e]) - mx) : 0.f;
            sum += p[e];
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) p[e] = 0.f;
        }
        // P planes of this head: [2 key chunks of 64][hi | lo][128 rows x 128 B]
        const uint32_t pb = smem0 + AT_P + (uint32_t)hh * 4u * AT_TILE + (uint32_t)(c0 >> 6) * 2u * AT_TILE;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v8[e] = p[g * 8 + e];
          uint4 hi, lo;
          at_split8(v8, hi, lo);
          const uint32_t a = pb + at_swz(r, ((c0 & 63) >> 3) + g);
          sts_v4(a, hi.x, hi.y, hi.z, hi.w);
          sts_v4(a + AT_TILE, lo.x, lo.y, lo.z, lo.w);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      const float inv = 1.0f / sum;
      named_bar_sync(2 + hh, 160);                           // with warp 0: it may issue this head's PV
      // ---- epilogue: O[:, this head's 32 channels] / rowsum ----
      mbar_wait(bar0 + 8 * (2 + hh), ph);
      tc_fence_after();
      uint32_t orr[32];
      tmem_ld32(tm + 256u + (uint32_t)(64 * hh) + ((uint32_t)(qd_ * 32) << 16) + (uint32_t)(hh * 32), orr);
      if (r < N) {
        float o[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) o[e] = __uint_as_float(orr[e]) * inv;
        const int col = pr * 64 + hh * 32;
        if (out != nullptr) {
          float *op = out + (row0 + r) * ldo + col;
#pragma unroll
          for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4 *>(op + e) = make_float4(o[e], o[e + 1], o[e + 2], o[e + 3]);
        }
        if (planes != nullptr) {
          uint32_t w[3][16];
#pragma unroll
          for (int e = 0; e < 32; e += 2) split3_pair(o[e], o[e + 1], w[0][e >> 1], w[1][e >> 1], w[2][e >> 1]);
#pragma unroll
          for (int pl = 0; pl < 3; ++pl) {
            __nv_bfloat16 *pp = planes + (size_t)pl * plane_stride + (row0 + r) * ldp + col;
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              *reinterpret_cast<uint4 *>(pp + 2 * e) = make_uint4(w[pl][e], w[pl][e + 1], w[pl][e + 2], w[pl][e + 3]);
          }
        }
      }
      tc_fence_before();
    }
    ph ^= 1u;
    __syncthreads();                                         // the pair's Q / K / V / P tiles and its TMEM columns are free
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tm, 512);
  }
}

bool attention_tc_supported(int N, int C, int heads, const float *q, int ldq, const float *k, int ldk, const float *v, int ldv,
                            const float *out, int ldo, const void *planes, long long plane_stride) {
  if (heads < 1 || C % heads != 0 || C / heads != 32 || C % 64 != 0) return false;
  if (N < 1 || N > 128) return false;
  auto al = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al(q) || !al(k) || !al(v) || ldq % 4 || ldk % 4 || ldv % 4) return false;
  if (out && (!al(out) || ldo % 4)) return false;
  if (planes && (!al(planes) || plane_stride % 8)) return false;
  return true;
}

int launch_attention_tc(const float *q, int ldq, const float *k, int ldk, const float *v, int ldv, float *out, int ldo, int B,
                        int N, int C, int heads, cudaStream_t stream, void *planes, long long plane_stride) {
  if (!attention_tc_supported(N, C, heads, q, ldq, k, ldk, v, ldv, out, ldo, planes, plane_stride))
    VKN_FAIL(VKN_E_UNSUPPORTED, "tcgen05 attention: needs head_dim 32, N <= 128 and 16-byte aligned rows");
  const int Nk = (N + 15) & ~15;
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) {
    VKN_CUDA_OK(cudaFuncSetAttribute(vkn_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(AT_SMEM + 2048)));
  }
  VKN_LAUNCH_MARK("vkn_attention_tc_kernel", stream);
  VKN_CUDA_OK(launch_chain(vkn_attention_tc_kernel, dim3(B), dim3(AT_THREADS), (size_t)AT_SMEM + 2048, stream, q, ldq, k, ldk, v, ldv,
                           out, ldo, (__nv_bfloat16 *)planes, plane_stride, C, N, C, 1.0f / sqrtf(32.0f),
                           make_idesc_f16(128, Nk, 0, 0), make_idesc_f16(128, 64, 0, 1)));
  return VKN_OK;
}

}  // namespace vkn
