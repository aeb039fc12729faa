// tcgen05 / TMA / mbarrier PTX wrappers and the host-side tensor-map builder shared by the tensor-core engines
// (gemm_tc.cu: pooling + mask conv; rowgemm_tc.cu: row GEMMs).  sm_100a only.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace vkn {

constexpr int TC_THREADS = 192;
constexpr int PX_BLK = 64;            // pixels per pipeline stage (pooling): 128 B of bf16 = one swizzle row
constexpr int MASK_TILE_P = 128;      // pixels per CTA (mask conv) = UMMA M
constexpr int CH_BLK = 64;            // channels per pipeline stage (mask conv)
constexpr uint32_t SPIN_LIMIT = 1u << 28;

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar),
               "r"(bytes)
               : "memory");
}
// Bounded wait: a protocol bug traps (reported as a launch failure) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < SPIN_LIMIT; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
// One lane of a converged warp (deterministic: the same lane for the same mask).  The MMA / TMA issue loops run
// warp-uniform with only the issuing instructions under elect_one(): the compiler then keeps descriptors, barrier
// addresses and loop state in uniform registers, which is what makes the tcgen05.mma issue loop a handful of
// instructions per MMA instead of ~15 (measured: profiles/r1_ncu_summary.md, "issue-bound MMA warp").
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// warp index as a value the compiler knows to be warp-uniform
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// TMA prefetch of a tile into L2 (no shared memory, no barrier): lets a CTA keep more DRAM requests in flight than
// its shared-memory ring holds -- the ring then only covers the L2 -> SM latency.
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// Explicit shared-state-space accesses (32-bit addresses).  Through generic pointers the compiler must assume a
// shared-memory store may alias the next bias load and serialises load -> convert -> store chains (measured: 54 cycles
// per element in the mask-conv epilogue); these keep loads batched ahead of the stores.
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("{\n\t.reg .b16 h;\n\tcvt.u16.u32 h, %1;\n\tst.shared.u16 [%0], h;\n\t}" ::"r"(addr), "r"(v));
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v)); }
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d));
}
__device__ __forceinline__ uint32_t bf16_bits(float v) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v)); }

// 256-bit global stores (sm_100): one full 32-byte sector per lane and instruction.  Row-per-thread epilogues write
// 32 different rows per instruction, so the L2 transaction count -- not bytes -- bounds them: v8 halves it against v4.
__device__ __forceinline__ void stg_v8(void *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                       uint32_t g, uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e),
               "r"(f), "r"(g), "r"(h)
               : "memory");
}

// TMA store of a shared-memory box (bulk-group completion)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the reverse: thread i of the warp writes its 32 values to row (lane base + i), 32 consecutive columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, SWIZZLE_128B, sm_100 version bit (cute::UMMA::SmemDescriptor layout:
// start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type=2 [61,64)).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}

// cute::UMMA::InstrDescriptor for kind::f16: D fp32, A/B bf16.
static uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// same with fp16 A and B operands (format code 0).  kind::f16 takes fp16 or bf16 operands but NOT one of each (mixed
// formats raise an illegal-instruction fault on B200: tools/probes/mma_probe.cu)
static uint32_t make_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// ---- host: tensor maps ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode(EncodeTiledFn *fn) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VKN_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (!p || qres != cudaDriverEntryPointSuccess) VKN_FAIL(VKN_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cached = (EncodeTiledFn)p;
  }
  *fn = cached;
  return VKN_OK;
}

// bf16 tensor, rank 2 or 3, dims/box given innermost first; 128B swizzle (inner box = 64 elements).
static int make_tmap_bf16(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint32_t *box) {
  EncodeTiledFn enc;
  VKN_TRY(get_encode(&enc));
  cuuint64_t gdim[3];
  cuuint64_t gstride[2];
  cuuint32_t bx[3], estr[3];
  uint64_t stride = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstride[i - 1] = stride;
    stride *= dims[i];
  }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstride, bx,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) VKN_FAIL(VKN_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return VKN_OK;
}


// bf16 tensor, rank 3, contiguous, NO swizzle (TMA stores from a dense shared-memory box); dims/box innermost first.
static int make_tmap_bf16_plain(CUtensorMap *map, const void *base, const uint64_t *dims, const uint32_t *box) {
  EncodeTiledFn enc;
  VKN_TRY(get_encode(&enc));
  cuuint64_t gdim[3] = {dims[0], dims[1], dims[2]};
  cuuint64_t gstride[2] = {dims[0] * 2, dims[0] * dims[1] * 2};
  cuuint32_t bx[3] = {box[0], box[1], box[2]}, estr[3] = {1, 1, 1};
  if (reinterpret_cast<uintptr_t>(base) & 15) VKN_FAIL(VKN_E_INVALID, "TMA: global base address must be 16-byte aligned");
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), gdim, gstride, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) VKN_FAIL(VKN_E_CUDA, "cuTensorMapEncodeTiled (plain) failed with CUresult %d", (int)r);
  return VKN_OK;
}

// rank-3 tensor map with explicit byte strides, element size and swizzle mode (TMA stores of the row-GEMM epilogue)
static int make_tmap_store(CUtensorMap *map, const void *base, bool f32, const uint64_t *dims, const uint64_t *stride_bytes,
                           const uint32_t *box, bool swizzle128) {
  EncodeTiledFn enc;
  VKN_TRY(get_encode(&enc));
  cuuint64_t gdim[3] = {dims[0], dims[1], dims[2]};
  cuuint64_t gstride[2] = {stride_bytes[0], stride_bytes[1]};
  cuuint32_t bx[3] = {box[0], box[1], box[2]}, estr[3] = {1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride_bytes[0] & 15) || (stride_bytes[1] & 15))
    VKN_FAIL(VKN_E_INVALID, "TMA store: base / strides must be 16-byte aligned");
  CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), gdim,
                   gstride, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) VKN_FAIL(VKN_E_CUDA, "cuTensorMapEncodeTiled (store) failed with CUresult %d", (int)r);
  return VKN_OK;
}

// bf16 tensor with explicit byte strides (innermost dimension contiguous); dims/box innermost first; 128B swizzle.
static int make_tmap_bf16_strided(CUtensorMap *map, const void *base, int rank, const uint64_t *dims,
                                  const uint64_t *stride_bytes /* [rank-1] */, const uint32_t *box) {
  EncodeTiledFn enc;
  VKN_TRY(get_encode(&enc));
  cuuint64_t gdim[3], gstride[2];
  cuuint32_t bx[3], estr[3];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstride[i - 1] = stride_bytes[i - 1];
      if (stride_bytes[i - 1] % 16 != 0) VKN_FAIL(VKN_E_INVALID, "TMA: global stride %llu is not a multiple of 16 bytes", (unsigned long long)stride_bytes[i - 1]);
    }
  }
  if (reinterpret_cast<uintptr_t>(base) & 15) VKN_FAIL(VKN_E_INVALID, "TMA: global base address must be 16-byte aligned");
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstride, bx,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) VKN_FAIL(VKN_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return VKN_OK;
}

}  // namespace vkn
