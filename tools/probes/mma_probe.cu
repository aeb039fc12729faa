// Hardware probe (not part of the product): which operand forms of tcgen05.mma kind::f16 behave as the PTX text says
// on sm_100a.  T1: A = fp16, B = bf16 (mixed formats in one instruction).  T2: scale-input-d (D = A.B + D * 2^-s).
// T3: A operand from tensor memory.  Prints max abs error against a double reference for each.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mma_probe mma_probe.cu -I../../video-k-net_b200/csrc -I../../include
#include <cuda_fp16.h>
#include <vector>
#include <cmath>
#include <cstdlib>
#include "tc.cuh"

namespace vkn { void set_error(const char *, ...) {} void launch_mark(const char *, cudaStream_t) {} bool pdl_enabled() { return false; } }
using namespace vkn;

__device__ __forceinline__ void umma_f16_scaled12(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, 12;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// A [128][64] 16-bit, B [64][64] 16-bit (row-major, K contiguous).  out[test][128][64] fp32.
__global__ void __launch_bounds__(128, 1) probe_kernel(const uint16_t *A, const uint16_t *B, float *out, uint32_t idesc_mixed, int test) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = smem, *sB = smem + 16384;
  uint64_t *bar = (uint64_t *)(smem + 16384 + 8192);
  uint32_t *slot = (uint32_t *)(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // K-major SWIZZLE_128B: row r, 16-byte chunk c -> (r/8)*1024 + (r%8)*128 + ((c ^ (r%8)) << 4)
  for (int i = tid; i < 128 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *(uint4 *)(sA + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = *(const uint4 *)(A + r * 64 + c * 8);
  }
  for (int i = tid; i < 64 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *(uint4 *)(sB + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = *(const uint4 *)(B + r * 64 + c * 8);
  }
  fence_proxy_async();
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  const uint64_t ad = umma_desc_sw128(smem_u32(sA), 0, 1024), bd = umma_desc_sw128(smem_u32(sB), 0, 1024);
  uint32_t ph = 0;
  // ---- T1: mixed formats, D0 = A.B^T (cols 0..63)
  if (tid == 0 && (test & 1)) {
    for (int k = 0; k < 4; ++k) umma_bf16(tm, ad + k * 2, bd + k * 2, idesc_mixed, k > 0);
  }
  if (tid == 0) umma_commit(smem_u32(bar));
  mbar_wait(smem_u32(bar), ph); ph ^= 1;
  tc_fence_after();
  {
    uint32_t r[32];
    for (int c0 = 0; c0 < 64; c0 += 32) {
      tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
      for (int e = 0; e < 32; ++e) out[(0 * 128 + tid) * 64 + c0 + e] = __uint_as_float(r[e]);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // ---- T2: D0 = A.B^T (k = 0 only, scaled input) : first MMA of the second pass uses D * 2^-12
  if (tid == 0 && (test & 2)) {
    umma_f16_scaled12(tm, ad, bd, idesc_mixed);                 // D = A0.B0 + D * 2^-12
  }
  if (tid == 0) umma_commit(smem_u32(bar));
  mbar_wait(smem_u32(bar), ph); ph ^= 1;
  tc_fence_after();
  {
    uint32_t r[32];
    for (int c0 = 0; c0 < 64; c0 += 32) {
      tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
      for (int e = 0; e < 32; ++e) out[(1 * 128 + tid) * 64 + c0 + e] = __uint_as_float(r[e]);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // ---- T3: A from TMEM: row tid's 64 K values packed 2 per column into columns 128..159
  {
    for (int c0 = 0; c0 < 32; c0 += 8) {
      uint32_t r[8];
      for (int e = 0; e < 8; ++e) r[e] = *(const uint32_t *)(A + tid * 64 + (c0 + e) * 2);
      tmem_st8(tm + ((uint32_t)(warp * 32) << 16) + 128 + c0, r);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0 && (test & 4)) {
    for (int k = 0; k < 4; ++k) umma_f16_ts(tm + 64, tm + 128 + k * 8, bd + k * 2, idesc_mixed, k > 0);
  }
  if (tid == 0) umma_commit(smem_u32(bar));
  mbar_wait(smem_u32(bar), ph); ph ^= 1;
  tc_fence_after();
  {
    uint32_t r[32];
    for (int c0 = 0; c0 < 64; c0 += 32) {
      tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + 64 + c0, r);
      for (int e = 0; e < 32; ++e) out[(2 * 128 + tid) * 64 + c0 + e] = __uint_as_float(r[e]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tm, 256);
  }
}

int main(int argc, char **argv) {
  const int test = argc > 1 ? atoi(argv[1]) : 7;
  const int afmt = argc > 2 ? atoi(argv[2]) : 0;
  std::vector<uint16_t> hA(128 * 64), hB(64 * 64);
  std::vector<double> fA(128 * 64), fB(64 * 64);
  srand(1);
  for (int i = 0; i < 128 * 64; ++i) {
    float v = (float)(rand() % 20001 - 10000) / 3000.f;
    if (i % 7 == 0) v *= 1e-3f;
    if (i % 11 == 0) v *= 3e-6f;             // fp16 subnormal range
    __half h = __float2half_rn(v);
    hA[i] = *(uint16_t *)&h;
    fA[i] = (double)__half2float(h);
  }
  for (int i = 0; i < 64 * 64; ++i) {
    float v = (float)(rand() % 20001 - 10000) / 20000.f;
    if (i % 5 == 0) v *= 1e-6f;              // below the fp16 range: only a true bf16 read gets this right
    __nv_bfloat16 b = __float2bfloat16_rn(v);
    hB[i] = *(uint16_t *)&b;
    fB[i] = (double)__bfloat162float(b);
  }
  uint16_t *dA, *dB;
  float *dO;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dO, 3 * 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0, 3 * 128 * 64 * 4);
  // idesc: D fp32 (1<<4), A fp16 (0<<7), B bf16 (1<<10), N = 64, M = 128
  const uint32_t idesc = (1u << 4) | ((uint32_t)afmt << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  probe_kernel<<<1, 128, 40000>>>(dA, dB, dO, idesc, test);
  cudaError_t e = cudaDeviceSynchronize();
  printf("test mask %d afmt %d: %s\n", test, afmt, cudaGetErrorString(e));
  std::vector<float> o(3 * 128 * 64);
  cudaMemcpy(o.data(), dO, o.size() * 4, cudaMemcpyDeviceToHost);
  double e1 = 0, e2 = 0, e3 = 0, mag = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < 64; ++n) {
      double full = 0, k0 = 0;
      for (int k = 0; k < 64; ++k) {
        full += fA[r * 64 + k] * fB[n * 64 + k];
        if (k < 16) k0 += fA[r * 64 + k] * fB[n * 64 + k];
      }
      mag = fmax(mag, fabs(full));
      e1 = fmax(e1, fabs(o[(0 * 128 + r) * 64 + n] - full));
      e2 = fmax(e2, fabs(o[(1 * 128 + r) * 64 + n] - (k0 + full / 4096.0)));
      e3 = fmax(e3, fabs(o[(2 * 128 + r) * 64 + n] - full));
    }
  printf("max |D| %.4f\nT1 mixed fp16 x bf16     max abs err %.3e\nT2 scale-input-d 2^-12   max abs err %.3e\nT3 A from tensor memory  max abs err %.3e\n", mag, e1, e2, e3);
  return 0;
}
