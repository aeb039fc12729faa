// Shared device/host helpers for libvknet (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "vknet.h"

namespace vkn {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
#define VKN_FAIL(code, ...)          \
  do {                               \
    ::vkn::set_error(__VA_ARGS__);   \
    return (code);                   \
  } while (0)
#define VKN_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) VKN_FAIL(VKN_E_CUDA, "%s -> %s", #expr, cudaGetErrorString(e__)); \
  } while (0)
#define VKN_TRY(expr)           \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != VKN_OK) return rc__; \
  } while (0)

// ---- launch accounting / live per-kernel timing (vkn_profile_begin/end) ---------------------------
// Every kernel launch goes through VKN_LAUNCH_MARK(name): it bumps the launch counter and, while a
// profile is open, records a CUDA event on the launching stream so that vkn_profile_end can report the
// device time of each launch (events bracket launches on the stream the kernels run on).
void launch_mark(const char *name, cudaStream_t stream);
#define VKN_LAUNCH_MARK(name, stream) ::vkn::launch_mark(name, stream)

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// Every kernel of the stage chain is launched with programmaticStreamSerialization: its prologue
// (barrier init, TMEM alloc, weight-tile prefetch -- data no kernel of the chain writes) overlaps the
// previous kernel's tail; `pdl_wait()` must precede the first access to anything a previous kernel wrote
// and every global store.  VKN_PDL=0 in the environment disables the attribute.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                       cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// cudaFuncSetAttribute is per DEVICE: a process-wide "done" flag breaks the second GPU of a process.  `mask` is a static
// bitmask owned by the call site (one bit per device ordinal); returns true when the attribute still has to be set on the
// current device and marks it.
static inline bool first_use_on_device(unsigned long long &mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- element access ---------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ void store_as(T *p, float v);
template <>
__device__ __forceinline__ void store_as<float>(float *p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_as<__nv_bfloat16>(__nv_bfloat16 *p, float v) {
  *p = __float2bfloat16_rn(v);
}

// 8 consecutive elements -> 8 floats (16-byte aligned for bf16, 32-byte for f32)
__device__ __forceinline__ void load8(const float *p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4 *>(p);
  const float4 b = *reinterpret_cast<const float4 *>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16 *p, float (&v)[8]) {
  const uint4 r = *reinterpret_cast<const uint4 *>(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
// 4 consecutive elements
__device__ __forceinline__ void load4(const float *p, float (&v)[4]) {
  const float4 a = *reinterpret_cast<const float4 *>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void load4(const __nv_bfloat16 *p, float (&v)[4]) {
  const uint2 r = *reinterpret_cast<const uint2 *>(p);
  v[0] = __uint_as_float(r.x << 16);
  v[1] = __uint_as_float(r.x & 0xffff0000u);
  v[2] = __uint_as_float(r.y << 16);
  v[3] = __uint_as_float(r.y & 0xffff0000u);
}
__device__ __forceinline__ void store4(float *p, const float (&v)[4]) {
  *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16 *p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t *>(&a);
  r.y = *reinterpret_cast<uint32_t *>(&b);
  *reinterpret_cast<uint2 *>(p) = r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// ex2.approx-based exponential (max rel. error ~2 ulp) + IEEE divide: error well below fp32 round-off of the gate sum
__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + __expf(-v)); }
// same with the hardware reciprocal (rcp.approx: 1 ulp) instead of the IEEE divide sequence: 4 instructions per element
__device__ __forceinline__ float sigmoid_fast(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
// fp32 pair -> three packed bf16x2 words (hi, mid, lo): v == hi + mid + lo to 24 bits.  cvt.rn.bf16x2.f32 converts both
// halves in ONE instruction (the scalar cvt runs on the quarter-rate conversion pipe and bounded the epilogues)
// fp32 pair -> two packed fp16x2 words: v == hi + lo to 22 bits (operands well inside the fp16 range)
__device__ __forceinline__ void split2h_pair(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  __half2 h = __floats2half2_rn(x0, x1);
  hi = *reinterpret_cast<uint32_t *>(&h);
  const float2 back = __half22float2(h);
  h = __floats2half2_rn(x0 - back.x, x1 - back.y);
  lo = *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ void split3_pair(float x0, float x1, uint32_t &hi, uint32_t &mid, uint32_t &lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  hi = *reinterpret_cast<uint32_t *>(&h);
  x0 -= __uint_as_float(hi << 16);
  x1 -= __uint_as_float(hi & 0xffff0000u);
  h = __floats2bfloat162_rn(x0, x1);
  mid = *reinterpret_cast<uint32_t *>(&h);
  x0 -= __uint_as_float(mid << 16);
  x1 -= __uint_as_float(mid & 0xffff0000u);
  h = __floats2bfloat162_rn(x0, x1);
  lo = *reinterpret_cast<uint32_t *>(&h);
}

// ---- the fused "rows" operators (smallops.cu) --------------------------------------------------
enum Pro : int {
  PRO_COPY = 0,     // v = S(a0)                      S(.) = sum of nsum slices (+ pbias[k] + pres[r][k])
  PRO_LN = 1,       // v = LN0(S(a0))
  PRO_LN_RELU = 2,  // v = relu(LN0(S(a0)))
  PRO_MUL = 3,      // v = a0 * a1
  PRO_GATE = 4,     // v = sigmoid(LN0(a0)) * LN1(a1) + sigmoid(LN2(a2)) * LN3(a3)   (kernel_updator.py:74-88)
  PRO_PLANES = 5    // a0 points at bf16 hi/mid/lo planes [3][M][lda] written by a producer's EPI_SPLIT3 epilogue
                    // (plane stride = sum_stride elements): copied with cp.async, no register pass (tensor-core path only)
};
enum Epi : int { EPI_BIAS = 1, EPI_RELU = 2, EPI_RES = 4, EPI_ROWSCALE = 8, EPI_SPLIT3 = 16, EPI_NOOUT = 32,
                 // tcgen05 row GEMM only: y = LayerNorm(acc + bias + res) * ln_g + ln_b (then EPI_RELU) in the epilogue;
                 // the tile must span the whole row (N <= 256)
                 EPI_LN = 64,
                 // tcgen05 row GEMM only, applied after bias / residual / LayerNorm and before ReLU, in this order:
                 // y = sigmoid(y);  y *= mul[row][col];  y += add2[row][col]   (the KernelUpdator gate arithmetic in the epilogue)
                 EPI_SIGMOID = 128, EPI_MUL = 256, EPI_ADD2 = 512,
                 // with EPI_SPLIT3 and padded rows (the mask conv's kernel operand): two fp16 planes (hi + lo, 22 bits) instead of
                 // three bf16 planes, for the mask conv's fp16 mode
                 EPI_SPLIT2H = 1024 };

struct RowSrc {
  const float *a[4];
  int lda[4];
  const float *ln_g[4];
  const float *ln_b[4];
  int pro;
  int nsum;                 // slices of a[0] to add (>= 1)
  long long sum_stride;     // elements between slices
  const float *pbias;       // optional [K] added before LN
  const float *pres;        // optional [M,K] residual added before LN
  int ldpres;
  int pres_mod;             // > 0: the residual row is (row % pres_mod) -- one [pres_mod,K] block broadcast over frames
};

struct LinArgs {
  RowSrc src;
  const void *w;            // [N,K] row-major, weight dtype
  int ldw;
  const float *bias;        // [N] or null
  const float *rowscale;    // [M] or null: bias is multiplied by rowscale[row] (EPI_ROWSCALE)
  const float *ln_g, *ln_b; // [N] LayerNorm affine of the fused epilogue (EPI_LN)
  const float *res;         // [M,N] residual added in the epilogue (EPI_RES)
  int ldres;
  const float *mul;         // [M,N] elementwise factor (EPI_MUL)
  int ldmul;
  const float *add2;        // [M,N] addend applied after the factor (EPI_ADD2)
  int ldadd2;
  float *out;               // [M,N] (slice z of a split-K launch writes out + z * out_split_stride)
  int ldo;
  long long out_split_stride;
  int ksplit;               // number of K slices (gridDim.z = nprob * ksplit)
  float *side;              // optional: the prologue result [M,K], written by the n-block-0 CTAs
  int ldside;
  int M, N, K;
  int epi;
  // EPI_SPLIT3: columns < split_C are also written as bf16 hi/mid/lo planes [3][B][Npad][split_C]
  // (row p = b * split_N + n); the operand layout of the tcgen05 mask-conv engine.
  __nv_bfloat16 *split_planes;
  int split_B, split_N, split_Npad, split_C;
  unsigned long long *dbg;   // optional: per-CTA phase timestamps (vkn_debug_timestamps), null in production
};

// launches (all enqueue on `stream`, never synchronise)
int launch_linear(const LinArgs *probs, int nprob, int w_dtype, cudaStream_t stream);
// static kernels [N,C] (+ bias [N]) -> a_ext rows [w | b | pad] and, when planes != null, bf16 hi/mid/lo planes [3][1][Npad][C]
int launch_pack_kernels(const float *w, const float *b, int N, int C, float *a_ext, int lda, void *planes, int Npad,
                        cudaStream_t stream);
unsigned long long *debug_ts_slot();   // api.cu: next launch's timestamp block or null
int launch_rowop(const RowSrc &src, float *out, int ldo, int M, int K, cudaStream_t stream);
// row transform once per row -> fp32 rows (optional) and / or bf16 hi/mid/lo planes [3][M][ldp] (optional)
int launch_rowprep(const RowSrc &src, float *out, int ldo, void *planes, int ldp, long long plane_stride, int M, int K,
                   cudaStream_t stream);
int launch_attention(const float *q, int ldq, const float *k, int ldk, const float *v, int ldv, float *out,
                     int ldo, int B, int N, int C, int heads, cudaStream_t stream, void *planes = nullptr,
                     long long plane_stride = 0);
// tcgen05 attention (attention_tc.cu): head_dim 32, N <= 128; same contract as launch_attention
bool attention_tc_supported(int N, int C, int heads, const float *q, int ldq, const float *k, int ldk, const float *v, int ldv,
                            const float *out, int ldo, const void *planes, long long plane_stride);
int launch_attention_tc(const float *q, int ldq, const float *k, int ldk, const float *v, int ldv, float *out, int ldo, int B,
                        int N, int C, int heads, cudaStream_t stream, void *planes, long long plane_stride);
// tcgen05 row GEMM (rowgemm_tc.cu): same contract as launch_linear for PRO_PLANES inputs and bf16 weights
bool linear_tc_supported(const LinArgs &a);
int launch_linear_tc(const LinArgs *probs, int nprob, cudaStream_t stream);
// chain form (rowgemm_tc.cu): a sequence of row GEMMs / row transforms over the same rows in ONE launch, a CTA per 128-row tile
struct ChainBuild;
ChainBuild *chain_begin(int M);
bool chain_empty(const ChainBuild *b);
int chain_add_gemm(ChainBuild *b, const LinArgs *probs, int nprob);
int chain_add_rowprep(ChainBuild *b, const RowSrc &src, float *out, int ldo, void *planes, int ldp, long long plane_stride, int K);
int chain_launch(ChainBuild *b, cudaStream_t stream);
// single-frame row engine (framechain.cu): the row operators of a stage as two cluster launches
bool frame_chain_supported(const VknShape &s, const VknHeadW &w);
size_t frame_chain_pack_bytes(const VknShape &s, const VknHeadW &w);
int launch_frame_chain_pack(const VknShape &s, const VknHeadW &w, void *out, size_t bytes, cudaStream_t stream);
int launch_frame_chain(const VknShape &s, const VknHeadW &w, const float *xp0, const float *cnt, const float *x_feat_in,
                       const float *pf, float *x_feat_out, float *obj0_ws, float *qkv_ws, float *obj_out, float *cls_out, float *a_ext, int lda, void *a_split, int Npad,
                       cudaStream_t stream);
// SIMT fp32 engines for the two big contractions (gemm_simt.cu)
int launch_pool_simt(const VknShape &s, const void *x, const void *mask, float *partials, float *cnt_partials,
                     int *nchunks, cudaStream_t stream);
int pool_simt_chunks(const VknShape &s);
int launch_pool_reduce(const VknShape &s, const float *partials, const float *cnt_partials, int nchunks,
                       float *xp0, float *cnt, cudaStream_t stream, void *planes = nullptr);
int launch_maskgemm_simt(const VknShape &s, const void *x, const float *a_ext, int lda, void *out,
                         cudaStream_t stream);
// tcgen05 / TMA engines (gemm_tc.cu)
bool tc_supported(const VknShape &s);
int launch_pool_tc(const VknShape &s, const void *x, const void *mask, float *partials, float *cnt_partials,
                   int *nchunks, cudaStream_t stream, const uint32_t *mask_bits = nullptr);
int pool_tc_chunks(const VknShape &s);
int launch_maskgemm_tc(const VknShape &s, const void *x, const float *a_ext, int lda, const void *a_split_ws,
                       void *out, cudaStream_t stream, uint32_t *bits_out = nullptr, bool planes_f16 = false);
// true when the persistent mask conv will take the kernels as two fp16 planes (the producer must then write EPI_SPLIT2H)
bool maskgemm_tc_planes_f16(const VknShape &s);
// bit-mask hand-off between the stages of the fused loop (1 bit per kernel and pixel instead of bf16 logits)
int maskgemm_tc_bits_wpr(const VknShape &s);
bool maskgemm_tc_persistent(const VknShape &s);
int maskgemm_tc_npad(const VknShape &s);
// post-loop mask path (postproc.cu)
int launch_rescale_masks(const void *masks, int dtype, int K, int H, int W, int up, int Hb, int Wb, int h, int w, int Ho,
                         int Wo, float thr, float *probs, uint8_t *bits, cudaStream_t stream);

// training-side cost matrix (matchcost.cu)
size_t match_cost_workspace_bytes(int N, int M, int HW);
int launch_match_cost(const float *mask_logits, const float *cls_logits, const float *gt_masks, const long long *gt_labels, int N,
                      int M, int HW, int ncls, const float *params7, float *cost, void *workspace, size_t workspace_bytes,
                      cudaStream_t stream);

// post-loop result assembly (panoptic.cu)
int launch_panoptic_merge(const float *masks, const float *scores, const int *labels, int T, int H, int W, int num_thing,
                          double inst_thr, double overlap_thr, int *seg, int *table, float *seg_scores, int *kept, int *counts,
                          void *workspace, size_t workspace_bytes, cudaStream_t stream);
int launch_mask_boxes(const void *masks, int elem_bytes, int K, int H, int W, float *boxes, cudaStream_t stream);
int launch_track_match(const float *bboxes, const long long *labels, const float *embeds, int n, int D, const long long *memo_labels,
                       const float *memo_embeds, const long long *memo_ids, int m, const float *thr6, int with_cats,
                       long long num_tracklets, int *sel, long long *ids, int *counts, void *workspace, size_t workspace_bytes,
                       cudaStream_t stream);

}  // namespace vkn
