// tcgen05 / TMA engines for the two big contractions of a stage (bf16 storage, sm_100a only).
//
//   pooling   D[n, c] = sum_p  M[n, p] * x[c, p]          M = 1[mask > thr] in {0,1}  (exact in bf16)
//             A = M tile   [128 kernels x 64 px]  K-major, written by the producer warps straight into
//                          the 128B-swizzled UMMA layout (threshold fused: no sigmoid / compare / float temporaries)
//             B = x tile   [C channels  x 64 px]  K-major, TMA (SWIZZLE_128B) from NCHW memory
//             accumulators [128 x C] fp32 in TMEM; split over pixel ranges, partials reduced in fixed order.
//
//   mask conv D[p, n] = sum_c  x[c, p] * a[n, c]          a = mask_kernel . ft_w (fp32), split into three
//             bf16 planes hi/mid/lo so that every product is exact and the fp32 TMEM accumulation
//             carries fp32-level accuracy (the next stage thresholds these logits at 0).
//             A = x tile   [128 px x 64 ch]  MN-major (pixels contiguous), TMA (SWIZZLE_128B) from NCHW
//             B = a planes [Npad x 64 ch]    K-major, TMA (SWIZZLE_128B)
//             accumulators [128 px x Npad] fp32 in TMEM; epilogue adds the folded bias, rounds to bf16 and
//             stores pixel-contiguous (lane = pixel -> coalesced).
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer (one elected lane issues tcgen05.mma),
// warps 2-5 operand producers / epilogue (TMEM -> registers -> global).  mbarrier pipelines between them.
// Reference math: knet/det/kernel_update_head.py:190-195 and :247-260.
#include "tc.cuh"

namespace vkn {

static int npad_of(int N) { return ceil_div(N, 16) * 16; }

bool tc_supported(const VknShape &s) {
  const int HW = s.H * s.W;
  if (s.x_dtype != VKN_BF16) return false;
  if (HW % 8 != 0 || HW < PX_BLK) return false;          // TMA: 16-byte global strides; one full box at least
  if (s.C % 64 != 0 || s.C > 256) return false;
  if (npad_of(s.N) > 176) return false;                  // mask-conv B operand must fit the smem pipeline
  return true;
}

// ---- pooling ------------------------------------------------------------------------------------
struct PoolPlan {
  int nblocks, bpc, nchunks, mtiles;
};
static PoolPlan pool_plan(const VknShape &s) {
  PoolPlan p;
  p.nblocks = ceil_div(s.H * s.W, PX_BLK);
  p.mtiles = ceil_div(s.N, 128);
  const int ctas_per_frame = p.nblocks * p.mtiles;
  int sms = 148;
  if (const char *e = getenv("VKN_POOL_CTAS")) sms = atoi(e) > 0 ? atoi(e) : 148;   // tuning knob: CTAs per launch
  int target = sms / (s.B > sms ? sms : s.B);            // CTAs available per frame for one wave
  if (target < 1) target = 1;
  p.bpc = ceil_div(ctas_per_frame, target);
  if (p.bpc < 1) p.bpc = 1;
  p.nchunks = ceil_div(p.nblocks, p.bpc);
  return p;
}
int pool_tc_chunks(const VknShape &s) { return tc_supported(s) ? pool_plan(s).nchunks : 0; }

constexpr int POOL_BITS_AHEAD = 4;     // bit-mask words loaded this many pixel blocks ahead of their use
constexpr int POOL_STAGES_MAX = 3;     // (x 32 KB + raw logits 16 KB + A tile 16 KB) x 3 = 192 KB at C = 256

__global__ void __launch_bounds__(TC_THREADS, 1)
vkn_pool_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_m,
                   float *__restrict__ partials, float *__restrict__ cnt_partials, int B, int N, int C, int HW,
                   int nblocks, int bpc, float thr, uint32_t idesc, int POOL_STAGES, const uint32_t *__restrict__ bits_in,
                   int wpr) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t x_bytes = (uint32_t)C * 128u;          // C rows x 64 px x 2 B
  const uint32_t m_bytes = 128u * 128u;                 // 128 kernels x 64 px x 2 B: raw logits, and the {0,1} A tile
  const uint32_t stage_bytes = x_bytes + 2u * m_bytes;  // x | raw mask logits (TMA) | thresholded A tile
  uint64_t *bars = (uint64_t *)(smem + POOL_STAGES * stage_bytes);
  const uint32_t bar0 = smem_u32(bars);
  // barrier slots: full[s] = bar0 + 8 s (TMA: x + raw logits), empty[s] = bar0 + 8 (STAGES + s) (MMA retired),
  // tmem_full = bar0 + 16 STAGES, afull[s] = bar0 + 8 (2 STAGES + 1 + s) (A tile written by the 4 producer warps)
  uint32_t *tmem_slot = (uint32_t *)(bars + 3 * POOL_STAGES + 1);
  const uint32_t smem0 = smem_u32(smem);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int chunk = blockIdx.x, mtile = blockIdx.y, b = blockIdx.z;
  const int blk_beg = chunk * bpc;
  const int nk = min(nblocks, blk_beg + bpc) - blk_beg;
  const uint32_t ncols = (uint32_t)C;                    // 64 / 128 / 256: a power of two >= 32

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tmap_x);
      prefetch_tmap(&tmap_m);
      for (int s = 0; s < POOL_STAGES; ++s) {
        mbar_init(bar0 + 8 * s, 1);                      // TMA expect_tx arrive (x tile + raw mask tile)
        mbar_init(bar0 + 8 * (POOL_STAGES + s), 1);      // tcgen05.commit
        mbar_init(bar0 + 8 * (2 * POOL_STAGES + 1 + s), 4);   // one arrive per producer warp
      }
      mbar_init(bar0 + 16 * POOL_STAGES, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), ncols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // setup above overlapped the previous kernel's tail; the mask logits come from it

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nk; ++i) {
        const int s = i % POOL_STAGES;
        const uint32_t ph = (uint32_t)(i / POOL_STAGES) & 1u;
        mbar_wait(bar0 + 8 * (POOL_STAGES + s), ph ^ 1u);
        mbar_expect_tx(bar0 + 8 * s, bits_in ? x_bytes : x_bytes + m_bytes);
        tma_load_3d(smem0 + s * stage_bytes, &tmap_x, bar0 + 8 * s, (blk_beg + i) * PX_BLK, 0, b);
        if (!bits_in)
          tma_load_3d(smem0 + s * stage_bytes + x_bytes, &tmap_m, bar0 + 8 * s, (blk_beg + i) * PX_BLK, mtile * 128, b);
      }
    }
  } else if (warp == 1) {
    // warp-uniform issue loop, one elected lane issues (see elect_one() in tc.cuh)
    const uint64_t bdesc0 = umma_desc_sw128(smem0, 0, 1024);
    const uint64_t adesc0 = bdesc0 + (uint64_t)((x_bytes + m_bytes) >> 4);
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < nk; ++i) {
      mbar_wait(bar0 + 8 * s, ph);                              // x tile landed
      mbar_wait(bar0 + 8 * (2 * POOL_STAGES + 1 + s), ph);      // A tile written
      tc_fence_after();
      if (elect_one()) {
        const uint64_t so = (uint64_t)((uint32_t)s * (stage_bytes >> 4));
#pragma unroll
        for (int k = 0; k < PX_BLK / 16; ++k)
          umma_bf16(tmem_base, adesc0 + so + (uint64_t)(k * 2), bdesc0 + so + (uint64_t)(k * 2), idesc, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(bar0 + 8 * (POOL_STAGES + s));        // frees the stage when these MMAs retire
        if (i == nk - 1) umma_commit(bar0 + 16 * POOL_STAGES);   // accumulators complete
      }
      __syncwarp();
      if (++s == POOL_STAGES) {
        s = 0;
        ph ^= 1u;
      }
    }
  } else {
    // ---- producers: threshold the mask logits into the swizzled A tile -------------------------
    const int pt = threadIdx.x - 64;                      // 0..127
    const int j = pt & 7;                                 // 16-byte chunk (8 pixels) within the 64-px row
    const int rbase = pt >> 3;                            // rows rbase + 16 i
    float cnt[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cnt[i] = 0.f;
    float cnt_row = 0.f;
    if (bits_in != nullptr) {
      // Bit-mask input (fused loop, inner stages): the previous stage's mask conv wrote 1 bit per (kernel, pixel) -- the
      // hard mask itself -- instead of bf16 logits: 16x fewer mask bytes on both sides.  Thread = kernel row: one 8-byte
      // load per 64-pixel block (prefetched one block ahead), expanded to {0, 1} bf16 in the swizzled A tile.
      const int r = pt, n = mtile * 128 + r;
      const uint2 *rowp = reinterpret_cast<const uint2 *>(bits_in + ((size_t)b * N + (n < N ? n : 0)) * wpr) + blk_beg;
      // 8-byte loads run POOL_BITS_AHEAD blocks ahead of their use (a block lasts ~0.7 us, an L2/DRAM load ~1 us)
      uint2 q[POOL_BITS_AHEAD];
#pragma unroll
      for (int a = 0; a < POOL_BITS_AHEAD; ++a) q[a] = (n < N && a < nk) ? __ldg(rowp + a) : make_uint2(0u, 0u);
      for (int it = 0; it < nk; ++it) {
        const uint2 cur = q[0];
#pragma unroll
        for (int a = 0; a + 1 < POOL_BITS_AHEAD; ++a) q[a] = q[a + 1];
        q[POOL_BITS_AHEAD - 1] = (n < N && it + POOL_BITS_AHEAD < nk) ? __ldg(rowp + it + POOL_BITS_AHEAD) : make_uint2(0u, 0u);
        const int s = it % POOL_STAGES;
        const uint32_t ph = (uint32_t)(it / POOL_STAGES) & 1u;
        mbar_wait(bar0 + 8 * s, ph);                         // x landed => the MMAs that last read this stage retired
        const uint32_t mt0 = smem0 + s * stage_bytes + x_bytes + m_bytes + (uint32_t)r * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t byte = ((j < 4 ? cur.x : cur.y) >> ((j & 3) * 8)) & 0xffu;
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[e] = ((byte >> (2 * e)) & 1u ? 0x3f80u : 0u) | ((byte >> (2 * e + 1)) & 1u ? 0x3f800000u : 0u);   // bf16 1.0
          sts_v4(mt0 + (uint32_t)((j ^ (r & 7)) << 4), o[0], o[1], o[2], o[3]);
        }
        cnt_row += (float)(__popc(cur.x) + __popc(cur.y));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + 8 * (2 * POOL_STAGES + 1 + s));
      }
    } else
    // The raw logits arrive by TMA in the same 128B-swizzled [row][64 px] layout the A tile uses, so a thread
    // reads and writes the SAME offset: no global-load latency on this path, the ring prefetches POOL_STAGES deep.
    for (int it = 0; it < nk; ++it) {
      const int s = it % POOL_STAGES;
      const uint32_t ph = (uint32_t)(it / POOL_STAGES) & 1u;
      mbar_wait(bar0 + 8 * s, ph);
      const uint8_t *rawt = smem + s * stage_bytes + x_bytes;
      uint8_t *mt = smem + s * stage_bytes + x_bytes + m_bytes;
      const int p = (blk_beg + it) * PX_BLK + j * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rbase + 16 * i;
        const int n = mtile * 128 + r;
        const bool live = (n < N && p < HW);
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);   // Swizzle<3,4,3>
        const uint4 rw = *reinterpret_cast<const uint4 *>(rawt + off);
        const uint32_t w[4] = {rw.x, rw.y, rw.z, rw.w};
        uint32_t o[4];
        float c = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
          const bool blo = live && (lo > thr), bhi = live && (hi > thr);
          o[e] = (blo ? 0x3f80u : 0u) | (bhi ? 0x3f800000u : 0u);   // bf16 1.0 = 0x3f80
          c += (blo ? 1.f : 0.f) + (bhi ? 1.f : 0.f);
        }
        cnt[i] += c;
        *reinterpret_cast<uint4 *>(mt + off) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      fence_proxy_async();                                 // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + 8 * (2 * POOL_STAGES + 1 + s));
    }
    // ---- epilogue: accumulators -> partial sums -----------------------------------------------
    mbar_wait(bar0 + 16 * POOL_STAGES, 0);
    tc_fence_after();
    pdl_trigger();
    const int q = warp & 3;                                // TMEM lane quarter this warp may read
    // accumulator rows are TMEM lanes: thread = kernel row.  Staged through shared memory (the pipeline
    // buffers are free once tmem_full fired) so that global stores are full 128-byte lines.
    float *stg = reinterpret_cast<float *>(smem) + (size_t)q * 32 * 36;      // [32 rows][36] per warp
    const int nbase = mtile * 128 + q * 32;
    for (int c0 = 0; c0 < C; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
      for (int e = 0; e < 32; e += 4)
        *reinterpret_cast<float4 *>(stg + lane * 36 + e) = make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]),
                                                                       __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + (lane >> 3), seg = lane & 7;
        const int n = nbase + row;
        if (n < N) {
          const float4 v = *reinterpret_cast<const float4 *>(stg + row * 36 + seg * 4);
          *reinterpret_cast<float4 *>(partials + (((size_t)chunk * B + b) * N + n) * C + c0 + seg * 4) = v;
        }
      }
      __syncwarp();
    }
    if (bits_in != nullptr) {                            // thread = row: its count is complete
      const int nn = mtile * 128 + pt;
      if (nn < N) cnt_partials[((size_t)chunk * B + b) * N + nn] = cnt_row;
    } else
    // pixel counts: the 8 chunk-threads of a row are adjacent lanes
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float c = cnt[i];
      c += __shfl_xor_sync(0xffffffffu, c, 1);
      c += __shfl_xor_sync(0xffffffffu, c, 2);
      c += __shfl_xor_sync(0xffffffffu, c, 4);
      const int nn = mtile * 128 + rbase + 16 * i;
      if (j == 0 && nn < N) cnt_partials[((size_t)chunk * B + b) * N + nn] = c;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ncols);
  }
}

int launch_pool_tc(const VknShape &s, const void *x, const void *mask, float *partials, float *cnt_partials,
                   int *nchunks, cudaStream_t stream, const uint32_t *mask_bits) {
  if (!tc_supported(s)) VKN_FAIL(VKN_E_UNSUPPORTED, "tcgen05 pooling: shape/dtype not supported");
  const int HW = s.H * s.W;
  const PoolPlan p = pool_plan(s);
  *nchunks = p.nchunks;
  CUtensorMap tmap;
  const uint64_t dims[3] = {(uint64_t)HW, (uint64_t)s.C, (uint64_t)s.B};
  const uint32_t box[3] = {(uint32_t)PX_BLK, (uint32_t)s.C, 1u};
  VKN_TRY(make_tmap_bf16(&tmap, x, 3, dims, box));
  // pipeline depth = blocks per CTA (<= 4): a CTA that reduces 2 pixel blocks needs 2 stages, and the smaller
  // footprint lets CTAs of other streams' kernels co-reside on the SM
  const int pool_stages = p.bpc < POOL_STAGES_MAX ? p.bpc : POOL_STAGES_MAX;
  size_t smem = (size_t)pool_stages * ((size_t)s.C * 128 + 2 * 128 * 128) + 1024 + 256;
  CUtensorMap tmap_m;
  {
    const uint64_t mdims[3] = {(uint64_t)HW, (uint64_t)s.N, (uint64_t)s.B};
    const uint32_t mbox[3] = {(uint32_t)PX_BLK, 128u, 1u};
    VKN_TRY(make_tmap_bf16(&tmap_m, mask_bits ? x : mask, 3, mdims, mbox));     // unused in bit-mask mode
  }
  if (smem < 4 * 32 * 36 * 4 + 2048) smem = 4 * 32 * 36 * 4 + 2048;      // epilogue staging
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) {
    VKN_CUDA_OK(cudaFuncSetAttribute(vkn_pool_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  dim3 grid(p.nchunks, p.mtiles, s.B);
  VKN_LAUNCH_MARK("vkn_pool_tc_kernel", stream);
  VKN_CUDA_OK(launch_chain(vkn_pool_tc_kernel, grid, dim3(TC_THREADS), smem, stream, tmap, tmap_m,
                           partials, cnt_partials, s.B, s.N, s.C, HW, p.nblocks, p.bpc, s.mask_thr_logit,
                           make_idesc_bf16(128, s.C, 0, 0), pool_stages, mask_bits, maskgemm_tc_bits_wpr(s)));
  return VKN_OK;
}

// ---- mask conv -------------------------------------------------------------------------------------
// The B operand (hi/mid/lo bf16 planes [3][B][Npad][C] of the folded kernels) is written by the fold
// Linear's epilogue (smallops.cu, EPI_SPLIT3).  Padding rows n >= N are never written: they only feed
// accumulator columns n >= N, which the epilogue discards.
__global__ void __launch_bounds__(TC_THREADS, 1)
vkn_maskgemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_a,
                       const float *__restrict__ a_ext, int lda, __nv_bfloat16 *__restrict__ out, int B, int N,
                       int Npad, int C, int HW, int stages, uint32_t idesc, uint32_t x_lbo, uint32_t x_sbo, int F) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t x_bytes = (uint32_t)CH_BLK * MASK_TILE_P * 2u;         // 64 ch x 128 px x 2 B = 16 KB
  const uint32_t a_plane = (uint32_t)Npad * 128u;                       // Npad rows x 64 ch x 2 B
  const uint32_t stage_bytes = x_bytes + 3u * a_plane;                  // (multiple of 1024: Npad % 16 == 0 -> a_plane % 2048 == 0)
  uint64_t *bars = (uint64_t *)(smem + (size_t)stages * stage_bytes);
  const uint32_t bar0 = smem_u32(bars);
  // full[s] = bar0 + 8 s, empty[s] = bar0 + 8 (stages + s), tmem_full = bar0 + 16 stages
  uint32_t *tmem_slot = (uint32_t *)(bars + 2 * stages + 1);
  float *bias_s = (float *)(tmem_slot + 2);                              // [Npad]
  const uint32_t smem0 = smem_u32(smem);
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int p0 = blockIdx.x * MASK_TILE_P, b = blockIdx.z;     // b = frame; its kernels are those of set kb
  const int kb = b / F;
  const int nk = C / CH_BLK;
  const uint32_t ncols = Npad <= 128 ? 128u : 256u;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tmap_x);
      prefetch_tmap(&tmap_a);
      for (int s = 0; s < stages; ++s) {
        mbar_init(bar0 + 8 * s, 1);
        mbar_init(bar0 + 8 * (stages + s), 1);
      }
      mbar_init(bar0 + 16 * stages, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), ncols);
  }
  pdl_wait();     // a_ext / the planes come from the previous kernel
  for (int n = threadIdx.x; n < ((Npad + 31) & ~31); n += TC_THREADS)
    bias_s[n] = (n < N) ? a_ext[((size_t)kb * N + n) * lda + C] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nk; ++i) {
        const int s = i % stages;
        const uint32_t ph = (uint32_t)(i / stages) & 1u;
        mbar_wait(bar0 + 8 * (stages + s), ph ^ 1u);
        mbar_expect_tx(bar0 + 8 * s, stage_bytes);
        const uint32_t xs = smem0 + s * stage_bytes;
        // x tile: two 64-pixel column groups, each [64 ch rows x 128 B]
        tma_load_3d(xs, &tmap_x, bar0 + 8 * s, p0, i * CH_BLK, b);
        tma_load_3d(xs + x_bytes / 2, &tmap_x, bar0 + 8 * s, p0 + 64, i * CH_BLK, b);
#pragma unroll
        for (int t = 0; t < 3; ++t)
          tma_load_2d(xs + x_bytes + t * a_plane, &tmap_a, bar0 + 8 * s, i * CH_BLK, (t * B + kb) * Npad);
      }
    }
  } else if (warp == 1) {
    // warp-uniform issue loop, one elected lane issues (see elect_one() in tc.cuh)
    // A (x): MN-major; one UMMA_K step = 16 channel rows of 128 B = 2048 B
    // B (a plane t): K-major; UMMA_K step = 32 B inside the 128-B swizzle row
    const uint64_t adesc0 = umma_desc_sw128(smem0, x_lbo, x_sbo);
    const uint64_t bdesc0 = umma_desc_sw128(smem0 + x_bytes, 0, 1024);
    const uint32_t a_plane16 = a_plane >> 4;
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < nk; ++i) {
      mbar_wait(bar0 + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t so = (uint64_t)((uint32_t)s * (stage_bytes >> 4));
#pragma unroll
        for (int t = 0; t < 3; ++t) {
#pragma unroll
          for (int k = 0; k < CH_BLK / 16; ++k)
            umma_bf16(tmem_base, adesc0 + so + (uint64_t)(k * (2048 >> 4)), bdesc0 + so + (uint64_t)((uint32_t)t * a_plane16 + k * 2),
                      idesc, (i > 0 || t > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(bar0 + 8 * (stages + s));
        if (i == nk - 1) umma_commit(bar0 + 16 * stages);
      }
      __syncwarp();
      if (++s == stages) {
        s = 0;
        ph ^= 1u;
      }
    }
  } else {
    // ---- epilogue: TMEM lane = pixel, column = kernel ---------------------------------------------
    mbar_wait(bar0 + 16 * stages, 0);
    tc_fence_after();
    pdl_trigger();
    const int q = warp & 3;
    // TMEM lane = pixel, column = kernel.  Each 32x32 chunk is transposed through shared memory (free
    // after tmem_full) so that a lane stores 8 consecutive pixels of one kernel row: 16-byte stores.
    __nv_bfloat16 *stg = reinterpret_cast<__nv_bfloat16 *>(smem) + (size_t)q * 32 * 40;   // [32 kernels][40] per warp
    const int pw = p0 + q * 32;                               // first pixel of this warp
    __nv_bfloat16 *ob = out + (size_t)b * N * HW;
    for (int n0 = 0; n0 < Npad; n0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)n0, r);
#pragma unroll
      for (int e = 0; e < 32; ++e)
        stg[e * 40 + lane] = __float2bfloat16_rn(__uint_as_float(r[e]) + bias_s[n0 + e]);
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int nl = it * 8 + (lane >> 2), seg = lane & 3;
        const int n = n0 + nl, p = pw + seg * 8;
        if (n < N && p < HW) {                                  // HW % 8 == 0: an 8-pixel segment never straddles the end
          const uint4 v = *reinterpret_cast<const uint4 *>(stg + nl * 40 + seg * 8);
          *reinterpret_cast<uint4 *>(ob + (size_t)n * HW + p) = v;
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ncols);
  }
}

// ---- mask conv, persistent form for frame batches ------------------------------------------------------
// When a launch holds several pixel tiles per SM (frame batches / clips), re-streaming the kernel planes for
// every tile (172 KB at N=100) costs ~3x the x traffic.  Here a CTA owns ONE frame's planes -- all three
// bf16 planes x all K chunks stay resident in shared memory -- and walks that frame's pixel tiles:
// only x streams (3-stage TMA ring), accumulators are double-buffered in TMEM so the tensor core works on
// tile i+1 while the epilogue warps drain tile i.  Requires Npad <= 112 (planes + ring must fit 227 KB).
// A resident plane chunk holds round_up(N, 8) rows (whole 8-row swizzle atoms), not Npad: the MMA (N = Npad) then
// reads up to 8 rows past it -- whatever follows in shared memory -- which only feeds accumulator columns >= N that
// the epilogue never stores.  The 12 KB this saves at N = 100 is what makes the third ring stage fit.
constexpr int MP_XS_MAX = 4;  // x ring depth (16 KB stages): as many as fit next to the planes and the 16 KB store staging
// Debug accounting (vkn_debug_timestamps): cycles each warp role spends blocked on a barrier, per CTA.
//   slot 0 CTA cycles | 1 TMA: ring slot free | 2 MMA: x landed | 3 MMA: accumulator free | 4 epilogue: accumulator
//   ready | 5 epilogue: staging box free + barriers | 6 tiles | 7 marker
#define MP_WAIT(acc, stmt)                 \
  do {                                     \
    if (dbg != nullptr) {                  \
      const long long t0__ = clock64();    \
      stmt;                                \
      acc += clock64() - t0__;             \
    } else {                               \
      stmt;                                \
    }                                      \
  } while (0)
constexpr int MP_ACC = 2;     // TMEM accumulator buffers (128 columns each)
constexpr int MP_THREADS = 256;   // warps 0-5 as in the one-tile kernel, warps 6-7: x converters of the fp16 mode

__global__ void __launch_bounds__(MP_THREADS, 1)
vkn_maskgemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_a,
                               const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_pf,
                               const float *__restrict__ a_ext, int lda, __nv_bfloat16 *__restrict__ out, int B, int N,
                               int Npad, int C, int HW, uint32_t idesc, uint32_t x_lbo, uint32_t x_sbo, int F, int MP_XS,
                               int total_tiles, int pf_dist, uint32_t *__restrict__ bits_out, int wpr, float thr,
                               int stg_bytes, unsigned long long *dbg, int Ng, int rows8, int nplanes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nk = C / CH_BLK;
  // Kernel groups (gridDim.y): more kernels than one resident plane set holds (N > 112: the real VPS counts 117 / 166) are
  // split into groups of Ng (64 or 96, a multiple of the 32-row store box); the CTAs of a group walk ALL pixel tiles with
  // their group's planes resident, so x is read once per group -- the second read of a tile comes from L2, the groups run
  // side by side.  Npad stays the row stride of the plane buffer; the MMA N and the epilogue extent are Ng.
  const int n_off = (int)blockIdx.y * Ng;
  const int Ncur = min(N - n_off, Ng);                                  // kernels this CTA computes
  const uint32_t x_bytes = (uint32_t)CH_BLK * MASK_TILE_P * 2u;         // 16 KB
  const uint32_t a_plane = (uint32_t)rows8 * 128u;                      // one plane, one 64-channel chunk (8-row atoms)
  // nplanes = 3: bf16 hi/mid/lo planes x bf16 x (round 1).  nplanes = 2: fp16 hi/lo planes (22 bits) x fp16 x -- a third fewer
  // MMAs and resident plane bytes.  kind::f16 does not take fp16 x bf16, so warps 6-7 convert every x stage bf16 -> fp16 in
  // place (exact: 8 significant bits into 11, |x| far inside the fp16 range) between the TMA ring and the MMA warp.
  const uint32_t planes_bytes = (uint32_t)nk * (uint32_t)nplanes * a_plane;
  uint8_t *xring = smem + planes_bytes;                                 // (planes_bytes is a multiple of 1024)
  uint8_t *stg_base = xring + MP_XS * x_bytes;                          // epilogue staging: 2 x [32 kernels][128 px] bf16
  uint64_t *bars = (uint64_t *)(stg_base + stg_bytes);                  // (no staging in bit-mask mode: one more ring stage)
  const uint32_t bar0 = smem_u32(bars);
  // barriers: 0 planes_full | 1 planes_free | X_FULL + s | X_EMPTY + s | ACC_FULL + a | ACC_EMPTY + a | X_CONV + s
  const int PL_FREE = 1, X_FULL = 2, X_EMPTY = 2 + MP_XS, ACC_FULL = 2 + 2 * MP_XS, ACC_EMPTY = 2 + 2 * MP_XS + MP_ACC;
  const int X_CONV = 2 + 2 * MP_XS + 2 * MP_ACC;
  uint32_t *tmem_slot = (uint32_t *)(bars + 2 + 3 * MP_XS + 2 * MP_ACC);
  float *bias_s = (float *)(((uintptr_t)(tmem_slot + 2) + 15) & ~(uintptr_t)15);   // [4 epilogue warps][2 segments][128], 16-byte aligned
  const uint32_t smem0 = smem_u32(smem), xring0 = smem_u32(xring);
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int ntiles = (HW + MASK_TILE_P - 1) / MASK_TILE_P;
  // Balanced schedule: the launch's tiles (frame-major) are cut into gridDim.x equal contiguous ranges, so every SM works
  // whatever the frame count (148 / frames CTAs per frame left 13-35 % of the SMs idle at 64 / 96 frames).  A range that
  // crosses a frame boundary is a second SEGMENT: the CTA swaps the resident planes (and the folded bias) once.
  const int g_lo = (int)((long long)blockIdx.x * total_tiles / gridDim.x);
  const int g_hi = (int)((long long)(blockIdx.x + 1) * total_tiles / gridDim.x);
  const long long t_start = dbg ? clock64() : 0;
  long long w0 = 0, w1 = 0;
  long long w2 = 0, w3 = 0, w4 = 0;
  if (dbg) dbg += (size_t)blockIdx.x * 16;                              // <= 148 CTAs x 16 slots

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tmap_x);
      prefetch_tmap(&tmap_a);
      prefetch_tmap(&tmap_o);
      prefetch_tmap(&tmap_pf);
      mbar_init(bar0, 1);
      mbar_init(bar0 + 8 * PL_FREE, 1);
      for (int s = 0; s < MP_XS; ++s) {
        mbar_init(bar0 + 8 * (X_FULL + s), 1);
        mbar_init(bar0 + 8 * (X_EMPTY + s), 1);
        mbar_init(bar0 + 8 * (X_CONV + s), 2);                                     // one arrive per converter warp
      }
      for (int a = 0; a < MP_ACC; ++a) {
        mbar_init(bar0 + 8 * (ACC_FULL + a), 1);
        mbar_init(bar0 + 8 * (ACC_EMPTY + a), 4);                               // one arrive per epilogue warp
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 128u * MP_ACC);
  }
  pdl_wait();     // a_ext / the planes come from the previous kernel
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int seg = 0;
      for (int g = g_lo; g < g_hi; ++seg) {
        const int b = g / ntiles, t0 = g - b * ntiles, kb = b / F;
        const int t1 = min(ntiles, t0 + (g_hi - g));
        if (seg > 0) MP_WAIT(w0, mbar_wait(bar0 + 8 * PL_FREE, (uint32_t)(seg - 1) & 1u));   // MMAs of the previous frame retired
        mbar_expect_tx(bar0, planes_bytes);
        for (int c = 0; c < nk; ++c)
          for (int t = 0; t < nplanes; ++t)
            tma_load_2d(smem0 + (uint32_t)(c * nplanes + t) * a_plane, &tmap_a, bar0, c * CH_BLK, (t * B + kb) * Npad + n_off);
        for (int tile = t0; tile < t1; ++tile) {
          const int p0 = tile * MASK_TILE_P;
          // The ring holds 48-64 KB per SM -- less than DRAM latency x bandwidth needs: the tile pf_dist ahead is pulled
          // into L2 by TMA prefetch (no shared memory), so the ring only has to cover the L2 -> SM hop.
          const int gp = g + (tile - t0) + pf_dist;         // global index of the tile to prefetch
          const bool pf = pf_dist > 0 && gp < g_hi;
          const int bp = pf ? gp / ntiles : 0, pp = pf ? (gp - bp * ntiles) * MASK_TILE_P : 0;
          for (int c = 0; c < nk; ++c) {
            // interleaved with the loads: one chunk ahead per chunk.  The prefetch map is unswizzled with 256-byte rows
            // (the whole 128-pixel tile per channel): half the TMA row requests of the two 128-byte-row load boxes
            if (pf) tma_prefetch_3d(&tmap_pf, pp, c * CH_BLK, bp);
            MP_WAIT(w0, mbar_wait(bar0 + 8 * (X_EMPTY + s), ph ^ 1u));
            mbar_expect_tx(bar0 + 8 * (X_FULL + s), x_bytes);
            tma_load_3d(xring0 + s * x_bytes, &tmap_x, bar0 + 8 * (X_FULL + s), p0, c * CH_BLK, b);
            tma_load_3d(xring0 + s * x_bytes + x_bytes / 2, &tmap_x, bar0 + 8 * (X_FULL + s), p0 + 64, c * CH_BLK, b);
            if (++s == MP_XS) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
        g += t1 - t0;
      }
      if (dbg) dbg[1] = (unsigned long long)w0;
    }
  } else if (warp == 1) {
    // Warp-uniform issue loop (all lanes wait, one elected lane issues): descriptors are 64-bit bases plus small
    // immediates kept in uniform registers -- a single thread walking divergent code needed ~15 dependent
    // instructions per tcgen05.mma (48 MMAs per tile).
    const uint64_t adesc0 = umma_desc_sw128(xring0, x_lbo, x_sbo);
    const uint64_t bdesc0 = umma_desc_sw128(smem0, 0, 1024);
    const uint32_t a_plane16 = a_plane >> 4;
    int s = 0;
    uint32_t xph = 0, li = 0;
    int seg = 0, seg_end = g_lo;                             // seg_end: first tile (global index) of the next segment
    for (int g = g_lo; g < g_hi; ++g, ++li) {
      if (g == seg_end) {                                    // new frame: its planes must have landed
        const int b = g / ntiles;
        seg_end = min(g_hi, (b + 1) * ntiles);
        MP_WAIT(w0, mbar_wait(bar0, (uint32_t)seg & 1u));
        ++seg;
      }
      const bool seg_last = (g + 1 == seg_end);
      const uint32_t buf = li & 1u;
      MP_WAIT(w1, mbar_wait(bar0 + 8 * (ACC_EMPTY + buf), ((li >> 1) & 1u) ^ 1u));
      tc_fence_after();
      const uint32_t dt = tmem_base + buf * 128u;
      for (int c = 0; c < nk; ++c) {
        MP_WAIT(w0, mbar_wait(bar0 + 8 * ((nplanes == 2 ? X_CONV : X_FULL) + s), xph));
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = adesc0 + (uint64_t)((uint32_t)s * (x_bytes >> 4));
          const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)(c * nplanes) * a_plane16);
          for (int t = 0; t < nplanes; ++t) {
#pragma unroll
            for (int k = 0; k < CH_BLK / 16; ++k)
              umma_bf16(dt, ad + (uint64_t)(k * (2048 >> 4)), bd + (uint64_t)((uint32_t)t * a_plane16 + k * 2), idesc,
                        (c > 0 || t > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(bar0 + 8 * (X_EMPTY + s));
          if (c == nk - 1) {
            umma_commit(bar0 + 8 * (ACC_FULL + buf));
            if (seg_last) umma_commit(bar0 + 8 * PL_FREE);   // the resident planes may be replaced
          }
        }
        __syncwarp();
        if (++s == MP_XS) {
          s = 0;
          xph ^= 1u;
        }
      }
    }
    if (dbg && lane == 0) {
      dbg[2] = (unsigned long long)w0;
      dbg[3] = (unsigned long long)w1;
      dbg[6] = li;
      dbg[7] = 0x6d61736b67656d6dull;
    }
  } else if (warp >= 6) {
    // ---- x converters (fp16 mode): every landed stage bf16 -> fp16 in place, then hand it to the MMA warp ----
    if (nplanes == 2) {
      const int ct = (int)threadIdx.x - 192;                 // 0..63
      int s = 0;
      uint32_t xph = 0;
      for (int g = g_lo; g < g_hi; ++g)
        for (int c = 0; c < nk; ++c) {
          mbar_wait(bar0 + 8 * (X_FULL + s), xph);
          const uint32_t base = xring0 + (uint32_t)s * x_bytes + (uint32_t)ct * 16u;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = lds_u4(base + (uint32_t)((half * 8 + i) * 1024));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __half2 h = __floats2half2_rn(__uint_as_float(w[e] << 16), __uint_as_float(w[e] & 0xffff0000u));
                w[e] = *reinterpret_cast<const uint32_t *>(&h);
              }
              sts_v4(base + (uint32_t)((half * 8 + i) * 1024), w[0], w[1], w[2], w[3]);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar0 + 8 * (X_CONV + s));
          if (++s == MP_XS) {
            s = 0;
            xph ^= 1u;
          }
        }
    }
  } else {
    // ---- epilogue: TMEM lane = pixel, column = kernel.  32 kernels x 128 pixels at a time are rounded to bf16 into a
    //      dense [32][128 px] staging box (lane = pixel: a warp writes 64 contiguous bytes, one conflict-free wavefront)
    //      and leave as ONE TMA store of full 256-byte rows; the box is double-buffered against the store reading it.
    //      (The earlier per-warp transpose cost 47 % of the shared-memory data pipe the tensor core reads through.)
    const int q = warp & 3;
    const bool leader = threadIdx.x == 64;
    const uint32_t stg0 = smem_u32(stg_base);
    uint32_t jj = 0, bias0 = 0;
    int li = 0, seg = 0, seg_end = g_lo, b = 0, tile = 0;
    for (int g = g_lo; g < g_hi; ++g, ++li, ++tile) {
      if (g == seg_end) {                                    // new frame: this warp's private copy of the folded bias
        b = g / ntiles;
        tile = g - b * ntiles;
        seg_end = min(g_hi, (b + 1) * ntiles);
        const int kb = b / F;
        bias0 = smem_u32(bias_s) + (uint32_t)(((warp - 2) * 2 + (seg & 1)) * 128) * 4u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int n = lane + 32 * i;
          const float v = (n < Ncur) ? a_ext[((size_t)kb * N + n_off + n) * lda + C] : 0.f;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias0 + (uint32_t)n * 4u), "f"(v));
        }
        __syncwarp();
        ++seg;
      }
      const int buf = li % MP_ACC;
      MP_WAIT(w0, mbar_wait(bar0 + 8 * (ACC_FULL + buf), (uint32_t)(li / MP_ACC) & 1u));
      tc_fence_after();
      if (g + 1 == g_hi) pdl_trigger();
      if (bits_out != nullptr) {
        // Inner stage of the fused loop: the only consumer is the next stage's hard threshold, so write the thresholded
        // BIT per (kernel, pixel) -- of the bf16-rounded logit, i.e. exactly what thresholding the stored map would
        // give -- 32 pixels of a warp per ballot word.  No staging, no TMA store, 16x fewer bytes.
        const int p = tile * MASK_TILE_P + q * 32 + lane;
        for (int n0 = 0; n0 < Ng; n0 += 32) {
          float4 bq[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) bq[e] = lds_f4(bias0 + (uint32_t)(n0 + 4 * e) * 4u);
          const float *bqf = reinterpret_cast<const float *>(bq);
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128 + n0), r);
          if (n0 + 32 >= Ng) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * (ACC_EMPTY + buf));
          }
          uint32_t mine = 0;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float v = __bfloat162float(__float2bfloat16_rn(__uint_as_float(r[e]) + bqf[e]));
            const uint32_t w = __ballot_sync(0xffffffffu, p < HW && v > thr);
            if (lane == e) mine = w;
          }
          if (n0 + lane < Ncur) bits_out[((size_t)b * N + n_off + n0 + lane) * wpr + tile * 4 + q] = mine;
        }
        continue;
      }
      for (int n0 = 0; n0 < Ng; n0 += 32, ++jj) {
        float4 bq[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) bq[e] = lds_f4(bias0 + (uint32_t)(n0 + 4 * e) * 4u);
        uint32_t r[32];
        MP_WAIT(w2, tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128 + n0), r));
        if (n0 + 32 >= Ng) {               // accumulators of this tile are in registers: hand the TMEM buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar0 + 8 * (ACC_EMPTY + buf));
        }
        MP_WAIT(w1, if (leader) bulk_wait_group_read<1>();   // the store issued two chunks ago has read its box
                named_bar_sync(1, 128));
        const uint32_t sb = stg0 + (jj & 1u) * (uint32_t)(32 * MASK_TILE_P * 2) + (uint32_t)(q * 32 + lane) * 2u;
        MP_WAIT(w3, _Pragma("unroll") for (int e = 0; e < 8; ++e) {
          sts_u16(sb + (4 * e + 0) * (MASK_TILE_P * 2), bf16_bits(__uint_as_float(r[4 * e + 0]) + bq[e].x));
          sts_u16(sb + (4 * e + 1) * (MASK_TILE_P * 2), bf16_bits(__uint_as_float(r[4 * e + 1]) + bq[e].y));
          sts_u16(sb + (4 * e + 2) * (MASK_TILE_P * 2), bf16_bits(__uint_as_float(r[4 * e + 2]) + bq[e].z));
          sts_u16(sb + (4 * e + 3) * (MASK_TILE_P * 2), bf16_bits(__uint_as_float(r[4 * e + 3]) + bq[e].w));
        });
        MP_WAIT(w4, fence_proxy_async(); named_bar_sync(1, 128));
        if (leader) {
          tma_store_3d(&tmap_o, stg0 + (jj & 1u) * (uint32_t)(32 * MASK_TILE_P * 2), tile * MASK_TILE_P, n_off + n0, b);
          bulk_commit_group();
        }
      }
    }
    if (leader) bulk_wait_all();
    if (dbg && leader) {
      dbg[4] = (unsigned long long)w0;
      dbg[5] = (unsigned long long)w1;
      dbg[8] = (unsigned long long)w2;
      dbg[9] = (unsigned long long)w3;
      dbg[10] = (unsigned long long)w4;
      dbg[0] = (unsigned long long)(clock64() - t_start);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128u * MP_ACC);
  }
}


// ---- mask conv, wide form: pixels are the N operand ---------------------------------------------------------------
// D[n, p] = sum_c a[n, c] x[c, p]:  A = one plane chunk [128 kernel rows x 64 ch] K-major (resident, as above; rows >= N are
// whatever follows in shared memory and only produce accumulator ROWS >= N, never read), B = x tile [256 px x 64 ch] MN-major
// straight from NCHW, accumulators [128 x 256 px] fp32 in TMEM, double-buffered (2 x 256 = all 512 columns).
// Why: every tcgen05.mma re-reads its A and B tiles from shared memory; at N = 112 the 3-plane product needs
// 48 x (4 KB + 3.5 KB) = 360 KB of operand reads per 128 pixels and the shared-memory pipe, not HBM, bounds the kernel.
// With N = 256 the same pixels cost 24 x (4 KB + 8 KB) = 288 KB and the MMAs run at their full 128-column rate.
// Epilogue: TMEM lane = kernel row, so a thread holds 32 CONSECUTIVE pixels of its row: bias is one register, the bf16
// row segment leaves as 64 contiguous bytes (no transpose, no staging), the bit mask is packed in-thread (no ballots).
constexpr int MW_TILE = 256;                 // pixels per tile
constexpr int MW_XS = 2;                     // x ring depth (32 KB stages)

__global__ void __launch_bounds__(TC_THREADS, 1)
vkn_maskgemm_tc_wide_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_a,
                            const float *__restrict__ a_ext, int lda, __nv_bfloat16 *__restrict__ out, int B, int N, int Npad,
                            int C, int HW, uint32_t idesc, uint32_t x_lbo, uint32_t x_sbo, int F, int total_tiles, int pf_dist,
                            uint32_t *__restrict__ bits_out, int wpr, float thr) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nk = C / CH_BLK;
  const uint32_t x_bytes = (uint32_t)CH_BLK * MW_TILE * 2u;             // 32 KB: four 64-px groups of [64 ch x 128 B]
  const uint32_t a_plane = (uint32_t)((N + 7) & ~7) * 128u;
  const uint32_t planes_bytes = (uint32_t)nk * 3u * a_plane;
  uint8_t *xring = smem + planes_bytes;
  uint64_t *bars = (uint64_t *)(xring + MW_XS * x_bytes);
  const uint32_t bar0 = smem_u32(bars);
  constexpr int PL_FREE = 1, X_FULL = 2, X_EMPTY = 2 + MW_XS, ACC_FULL = 2 + 2 * MW_XS, ACC_EMPTY = 2 + 2 * MW_XS + 2;
  uint32_t *tmem_slot = (uint32_t *)(bars + 2 + 2 * MW_XS + 4);
  const uint32_t smem0 = smem_u32(smem), xring0 = smem_u32(xring);
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int ntiles = (HW + MW_TILE - 1) / MW_TILE;
  const int g_lo = (int)((long long)blockIdx.x * total_tiles / gridDim.x);
  const int g_hi = (int)((long long)(blockIdx.x + 1) * total_tiles / gridDim.x);

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tmap_x);
      prefetch_tmap(&tmap_a);
      mbar_init(bar0, 1);
      mbar_init(bar0 + 8 * PL_FREE, 1);
      for (int s = 0; s < MW_XS; ++s) {
        mbar_init(bar0 + 8 * (X_FULL + s), 1);
        mbar_init(bar0 + 8 * (X_EMPTY + s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(bar0 + 8 * (ACC_FULL + a), 1);
        mbar_init(bar0 + 8 * (ACC_EMPTY + a), 4);                       // one arrive per epilogue warp
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 512u);
  }
  pdl_wait();     // a_ext / the planes come from the previous kernel
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0, seg = 0;
      uint32_t ph = 0;
      for (int g = g_lo; g < g_hi; ++seg) {
        const int b = g / ntiles, t0 = g - b * ntiles, kb = b / F;
        const int t1 = min(ntiles, t0 + (g_hi - g));
        if (seg > 0) mbar_wait(bar0 + 8 * PL_FREE, (uint32_t)(seg - 1) & 1u);     // MMAs of the previous frame retired
        mbar_expect_tx(bar0, planes_bytes);
        for (int c = 0; c < nk; ++c)
          for (int t = 0; t < 3; ++t)
            tma_load_2d(smem0 + (uint32_t)(c * 3 + t) * a_plane, &tmap_a, bar0, c * CH_BLK, (t * B + kb) * Npad);
        for (int tile = t0; tile < t1; ++tile) {
          const int p0 = tile * MW_TILE;
          const int gp = g + (tile - t0) + pf_dist;                      // L2 prefetch of the tile pf_dist ahead
          const bool pf = pf_dist > 0 && gp < g_hi;
          const int bp = pf ? gp / ntiles : 0, pp = pf ? (gp - bp * ntiles) * MW_TILE : 0;
          for (int c = 0; c < nk; ++c) {
            if (pf) {
#pragma unroll
              for (int q = 0; q < 4; ++q) tma_prefetch_3d(&tmap_x, pp + 64 * q, c * CH_BLK, bp);
            }
            mbar_wait(bar0 + 8 * (X_EMPTY + s), ph ^ 1u);
            mbar_expect_tx(bar0 + 8 * (X_FULL + s), x_bytes);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              tma_load_3d(xring0 + s * x_bytes + q * (x_bytes / 4), &tmap_x, bar0 + 8 * (X_FULL + s), p0 + 64 * q, c * CH_BLK, b);
            if (++s == MW_XS) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
        g += t1 - t0;
      }
    }
  } else if (warp == 1) {
    // warp-uniform issue loop, one elected lane issues (see elect_one() in tc.cuh)
    const uint64_t adesc0 = umma_desc_sw128(smem0, 0, 1024);              // planes: K-major, 8-row groups 1024 B apart
    const uint64_t bdesc0 = umma_desc_sw128(xring0, x_lbo, x_sbo);        // x: MN-major, 64-px groups x_lbo apart
    const uint32_t a_plane16 = a_plane >> 4;
    int s = 0;
    uint32_t xph = 0, li = 0;
    int seg = 0, seg_end = g_lo;
    for (int g = g_lo; g < g_hi; ++g, ++li) {
      if (g == seg_end) {                                    // new frame: its planes must have landed
        const int b = g / ntiles;
        seg_end = min(g_hi, (b + 1) * ntiles);
        mbar_wait(bar0, (uint32_t)seg & 1u);
        ++seg;
      }
      const bool seg_last = (g + 1 == seg_end);
      const uint32_t buf = li & 1u;
      mbar_wait(bar0 + 8 * (ACC_EMPTY + buf), ((li >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t dt = tmem_base + buf * (uint32_t)MW_TILE;
      for (int c = 0; c < nk; ++c) {
        mbar_wait(bar0 + 8 * (X_FULL + s), xph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = adesc0 + (uint64_t)((uint32_t)(c * 3) * a_plane16);
          const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)s * (x_bytes >> 4));
#pragma unroll
          for (int t = 0; t < 3; ++t) {
#pragma unroll
            for (int k = 0; k < CH_BLK / 16; ++k)
              umma_bf16(dt, ad + (uint64_t)((uint32_t)t * a_plane16 + k * 2), bd + (uint64_t)(k * (2048 >> 4)), idesc,
                        (c > 0 || t > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(bar0 + 8 * (X_EMPTY + s));
          if (c == nk - 1) {
            umma_commit(bar0 + 8 * (ACC_FULL + buf));
            if (seg_last) umma_commit(bar0 + 8 * PL_FREE);
          }
        }
        __syncwarp();
        if (++s == MW_XS) {
          s = 0;
          xph ^= 1u;
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int n = q * 32 + lane;                             // this thread's kernel row
    float bias = 0.f;
    int li = 0, seg_end = g_lo, b = 0, tile = 0;
    for (int g = g_lo; g < g_hi; ++g, ++li, ++tile) {
      if (g == seg_end) {
        b = g / ntiles;
        tile = g - b * ntiles;
        seg_end = min(g_hi, (b + 1) * ntiles);
        bias = (n < N) ? a_ext[((size_t)(b / F) * N + n) * lda + C] : 0.f;
      }
      const int buf = li & 1;
      mbar_wait(bar0 + 8 * (ACC_FULL + buf), (uint32_t)(li >> 1) & 1u);
      tc_fence_after();
      if (g + 1 == g_hi) pdl_trigger();
      const size_t rowbase = ((size_t)b * N + (n < N ? n : 0));
#pragma unroll 1
      for (int c0 = 0; c0 < MW_TILE; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * MW_TILE + c0), r);
        if (c0 + 32 >= MW_TILE) {          // accumulators of this tile are in registers: hand the TMEM buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar0 + 8 * (ACC_EMPTY + buf));
        }
        const int p0 = tile * MW_TILE + c0;
        if (n >= N) continue;
        const int np = max(0, min(32, HW - p0));             // HW % 8 == 0: np is a multiple of 8
        if (bits_out != nullptr) {                           // every word of the row is written (pixels >= HW as 0)
          uint32_t word = 0;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float v = __bfloat162float(__float2bfloat16_rn(__uint_as_float(r[e]) + bias));
            word |= (e < np && v > thr) ? (1u << e) : 0u;
          }
          bits_out[rowbase * wpr + (size_t)tile * (MW_TILE / 32) + (c0 >> 5)] = word;
        } else if (np > 0) {
          uint32_t w[16];
#pragma unroll
          for (int e = 0; e < 32; e += 2)
            w[e >> 1] = bf16_bits(__uint_as_float(r[e]) + bias) | (bf16_bits(__uint_as_float(r[e + 1]) + bias) << 16);
          __nv_bfloat16 *op = out + rowbase * HW + p0;
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            if (2 * e < np) *reinterpret_cast<uint4 *>(op + 2 * e) = make_uint4(w[e], w[e + 1], w[e + 2], w[e + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

int maskgemm_tc_npad(const VknShape &s) { return npad_of(s.N); }

// words of 32 pixel bits per (frame, kernel) row of the bit-mask hand-off: whole 128-pixel tiles
int maskgemm_tc_bits_wpr(const VknShape &s) { return ceil_div(s.H * s.W, MW_TILE) * (MW_TILE / 32); }
// true when launch_maskgemm_tc takes the persistent kernel (the one that can emit the bit mask)
bool maskgemm_tc_persistent(const VknShape &s) {
  if (!tc_supported(s)) return false;
  const int F = s.frames_per_set > 1 ? s.frames_per_set : 1;
  // (N > 112 kernels: two kernel groups of 64 / 96, see the kernel; VKN_MASK_GROUPS=0 restricts it to one group)
  bool groups_ok = true;
  if (const char *e = getenv("VKN_MASK_GROUPS")) groups_ok = e[0] != '0';
  const bool fits = npad_of(s.N) <= 112 || (groups_ok && npad_of(s.N) <= 192);
  bool persist = fits && ceil_div(s.H * s.W, MASK_TILE_P) * s.B * F >= 2 * 148;
  if (const char *e = getenv("VKN_MASK_PERSIST")) persist = (e[0] == '1') && fits;
  return persist;
}

// fp16 mode of the persistent kernel (two fp16 planes of the kernels, x converted in the ring): default on; VKN_MASK_F16=0
// keeps the three bf16 planes of round 1
bool maskgemm_tc_planes_f16(const VknShape &s) {
  if (!maskgemm_tc_persistent(s)) return false;
  if (const char *e = getenv("VKN_MASK_WIDE")) {
    if (e[0] == '1') return false;                            // the pixels-as-N variant reads three bf16 planes
  }
  if (const char *e = getenv("VKN_MASK_F16")) return e[0] != '0';
  return true;
}

int launch_maskgemm_tc(const VknShape &s, const void *x, const float *a_ext, int lda, const void *a_split_ws, void *out,
                       cudaStream_t stream, uint32_t *bits_out, bool planes_f16) {
  if (!tc_supported(s)) VKN_FAIL(VKN_E_UNSUPPORTED, "tcgen05 mask conv: shape/dtype not supported");
  const int HW = s.H * s.W, Npad = npad_of(s.N);
  const int F = s.frames_per_set > 1 ? s.frames_per_set : 1;
  CUtensorMap tmx, tma;
  {
    const uint64_t dims[3] = {(uint64_t)HW, (uint64_t)s.C, (uint64_t)s.B * F};
    const uint32_t box[3] = {64u, (uint32_t)CH_BLK, 1u};
    VKN_TRY(make_tmap_bf16(&tmx, x, 3, dims, box));
  }
  {
    const uint64_t dims[2] = {(uint64_t)s.C, (uint64_t)3 * s.B * Npad};
    const uint32_t box[2] = {(uint32_t)CH_BLK, (uint32_t)Npad};
    VKN_TRY(make_tmap_bf16(&tma, a_split_ws, 2, dims, box));
  }
  const size_t stage_bytes = (size_t)CH_BLK * MASK_TILE_P * 2 + (size_t)3 * Npad * 128;
  int stages = (int)((220 * 1024) / stage_bytes);
  if (stages > 2) stages = 2;      // 2 stages hide the TMA latency of the 4-chunk K loop; smaller footprint -> co-residency
  if (stages > s.C / CH_BLK) stages = s.C / CH_BLK;
  if (stages < 1) VKN_FAIL(VKN_E_UNSUPPORTED, "tcgen05 mask conv: N %d too large for shared memory", s.N);
  const size_t smem = (size_t)stages * stage_bytes + 1024 + (2 * stages + 1) * 8 + 16 + (size_t)(Npad + 32) * 4 + 64;
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) {
    VKN_CUDA_OK(cudaFuncSetAttribute(vkn_maskgemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  // x operand (A, MN-major, SWIZZLE_128B): 64-px groups are 8192 B apart (LBO), 8-channel-row groups 1024 B (SBO)
  const uint32_t x_lbo = (uint32_t)CH_BLK * 128u, x_sbo = 1024u;
  const int ntiles = ceil_div(HW, MASK_TILE_P), frames = s.B * F;
  const bool persist = maskgemm_tc_persistent(s);                // several tiles per SM: keep the planes resident
  if (bits_out && !persist) VKN_FAIL(VKN_E_INVALID, "tcgen05 mask conv: the bit-mask output needs the persistent kernel");
  if (!bits_out && !out) VKN_FAIL(VKN_E_INVALID, "tcgen05 mask conv: no output");
  const int rows8w = (s.N + 7) & ~7;
  const size_t wsmem = (size_t)(s.C / CH_BLK) * 3 * rows8w * 128 + MW_XS * (size_t)CH_BLK * MW_TILE * 2 + (2 + 2 * MW_XS + 4) * 8 + 16 +
                       1024 + 64;
  // opt-in (VKN_MASK_WIDE=1): measured equal to the 128-pixel-tile kernel (163 us per 64 frames either way) -- both are
  // bound by the bytes of x a CTA can keep in flight next to 160 KB of resident planes, not by the operand pipe
  bool wide = false;
  if (const char *e = getenv("VKN_MASK_WIDE")) wide = persist && wsmem <= 227 * 1024 && e[0] == '1';
  if (wide) {
    const int nt = ceil_div(HW, MW_TILE), total_tiles = nt * frames;
    int pf_dist = 1;
    if (const char *e = getenv("VKN_MASK_PF")) pf_dist = atoi(e);
    const int grid_x = total_tiles < 148 ? total_tiles : 148;
    {     // resident planes: whole 8-row atoms only
      const uint64_t dims[2] = {(uint64_t)s.C, (uint64_t)3 * s.B * Npad};
      const uint32_t box[2] = {(uint32_t)CH_BLK, (uint32_t)rows8w};
      VKN_TRY(make_tmap_bf16(&tma, a_split_ws, 2, dims, box));
    }
    static unsigned long long wattr = 0;
    if (first_use_on_device(wattr)) {
      VKN_CUDA_OK(cudaFuncSetAttribute(vkn_maskgemm_tc_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    VKN_LAUNCH_MARK("vkn_maskgemm_tc_wide_kernel", stream);
    VKN_CUDA_OK(launch_chain(vkn_maskgemm_tc_wide_kernel, dim3(grid_x), dim3(TC_THREADS), wsmem, stream, tmx, tma, a_ext, lda,
                             (__nv_bfloat16 *)out, s.B, s.N, Npad, s.C, HW, make_idesc_bf16(128, MW_TILE, 0, 1), x_lbo, x_sbo, F,
                             total_tiles, pf_dist, bits_out, maskgemm_tc_bits_wpr(s), s.mask_thr_logit));
    return VKN_OK;
  }
  if (persist) {
    const int total_tiles = ntiles * frames;
    int pf_dist = 1;                                         // L2 prefetch distance in tiles (VKN_MASK_PF; measured 0/1/2/4: 1 is best)
    if (const char *e = getenv("VKN_MASK_PF")) pf_dist = atoi(e);
    const int ngroups = Npad <= 112 ? 1 : 2;
    const int Ng = ngroups == 1 ? Npad : (Npad <= 128 ? 64 : 96);      // kernels per group: MMA N, multiple of the 32-row store box
    int sms = 148;                                           // VKN_MASK_CTAS: leave SMs to the kernels of other streams
    if (const char *e = getenv("VKN_MASK_CTAS")) sms = atoi(e) > 0 && atoi(e) <= 148 ? atoi(e) : 148;
    const int per_grp = sms / ngroups;
    const int grid_x = total_tiles < per_grp ? total_tiles : per_grp;
    const int rows8 = ngroups == 1 ? ((s.N + 7) & ~7) : Ng;
    {     // resident planes: whole 8-row atoms only (see the kernel comment)
      const uint64_t dims[2] = {(uint64_t)s.C, (uint64_t)3 * s.B * Npad};
      const uint32_t box[2] = {(uint32_t)CH_BLK, (uint32_t)rows8};
      VKN_TRY(make_tmap_bf16(&tma, a_split_ws, 2, dims, box));
    }
    const int stg_bytes = bits_out ? 0 : 2 * 32 * MASK_TILE_P * 2;      // the bit-mask epilogue needs no staging box
    const int nplanes = planes_f16 ? 2 : 3;
    auto psmem_of = [&](int xs) {
      return (size_t)(s.C / CH_BLK) * nplanes * rows8 * 128 + xs * (size_t)CH_BLK * MASK_TILE_P * 2 + (size_t)stg_bytes +
             (2 + 3 * xs + 2 * MP_ACC) * 8 + 16 + 4 * 2 * 128 * 4 + 1024 + 64;
    };
    int xs_depth = planes_f16 ? MP_XS_MAX + 2 : MP_XS_MAX;     // the converter adds a hop: a deeper ring where it fits
    while (xs_depth > 2 && psmem_of(xs_depth) > 227 * 1024) --xs_depth;
    const size_t psmem = psmem_of(xs_depth);
    static unsigned long long pattr = 0;
    if (first_use_on_device(pattr)) {
      VKN_CUDA_OK(cudaFuncSetAttribute(vkn_maskgemm_tc_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    if (psmem > 227 * 1024) VKN_FAIL(VKN_E_UNSUPPORTED, "persistent mask conv: shared memory %zu exceeds 227 KB", psmem);
    VKN_LAUNCH_MARK("vkn_maskgemm_tc_persist_kernel", stream);
    CUtensorMap tmo;
    {
      const uint64_t dims[3] = {(uint64_t)HW, (uint64_t)s.N, (uint64_t)frames};
      const uint32_t box[3] = {(uint32_t)MASK_TILE_P, 32u, 1u};
      VKN_TRY(make_tmap_bf16_plain(&tmo, out ? out : x, dims, box));     // unused (never stored through) in bit-mask mode
    }
    CUtensorMap tmpf;
    {
      const uint64_t dims[3] = {(uint64_t)HW, (uint64_t)s.C, (uint64_t)frames};
      const uint32_t box[3] = {(uint32_t)MASK_TILE_P, (uint32_t)CH_BLK, 1u};
      VKN_TRY(make_tmap_bf16_plain(&tmpf, x, dims, box));
    }
    VKN_CUDA_OK(launch_chain(vkn_maskgemm_tc_persist_kernel, dim3(grid_x, ngroups), dim3(MP_THREADS), psmem, stream, tmx, tma,
                             tmo, tmpf, a_ext, lda, (__nv_bfloat16 *)out, s.B, s.N, Npad, s.C, HW,
                             planes_f16 ? make_idesc_f16(128, Ng, 1, 0) : make_idesc_bf16(128, Ng, 1, 0),
                             x_lbo, x_sbo, F, xs_depth, total_tiles, pf_dist, bits_out, maskgemm_tc_bits_wpr(s), s.mask_thr_logit,
                             stg_bytes, debug_ts_slot(), Ng, rows8, nplanes));
    return VKN_OK;
  }
  dim3 grid(ceil_div(HW, MASK_TILE_P), 1, s.B * F);
  VKN_LAUNCH_MARK("vkn_maskgemm_tc_kernel", stream);
  VKN_CUDA_OK(launch_chain(vkn_maskgemm_tc_kernel, grid, dim3(TC_THREADS), smem, stream, tmx, tma, a_ext, lda,
                           (__nv_bfloat16 *)out, s.B, s.N, Npad, s.C, HW, stages, make_idesc_bf16(128, Npad, 1, 0), x_lbo,
                           x_sbo, F));
  return VKN_OK;
}

}  // namespace vkn
