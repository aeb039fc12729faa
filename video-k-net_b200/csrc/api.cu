// C-ABI entry points of libvknet.so and the host-side orchestration of one KernelUpdateHead stage.
// Everything here only enqueues kernels on the caller's stream: no allocation, no synchronisation.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace vkn {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- launch accounting / profiling ------------------------------------------------------------------
constexpr int PROF_MAX = 512;
static unsigned long long g_launches = 0;
static struct {
  bool on = false;
  int n = 0;
  cudaStream_t st = nullptr;
  cudaEvent_t ev[PROF_MAX + 1];
  const char *names[PROF_MAX];
  bool created = false;
} g_prof;

// ---- debug: per-CTA phase timestamps of the row-operator launches (tools/linear_timeline.py) ---------------
constexpr size_t DBG_STRIDE = 4096 * 8;       // u64 per launch: up to 4096 CTAs x 8 slots
static unsigned long long *g_ts = nullptr;
static size_t g_ts_cap = 0, g_ts_launch = 0;
unsigned long long *debug_ts_slot() {
  if (!g_ts || (g_ts_launch + 1) * DBG_STRIDE > g_ts_cap) return nullptr;
  return g_ts + (g_ts_launch++) * DBG_STRIDE;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("VKN_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

void launch_mark(const char *name, cudaStream_t stream) {
  ++g_launches;
  if (g_prof.on && g_prof.n < PROF_MAX) {
    cudaEventRecord(g_prof.ev[g_prof.n], stream);
    g_prof.names[g_prof.n] = name;
    g_prof.st = stream;
    ++g_prof.n;
  }
}

constexpr int FFN_KSPLIT = 8;   // max K-slices of the second FFN Linear (K = ffn_dim); fewer when there are many rows
constexpr int A_EXT_PAD = 8;    // a_ext row = C folded channels + bias column, padded to C + 8

// ---- workspace ----------------------------------------------------------------------------------
struct Layout {
  // pooling
  float *pool_part, *cnt_part, *xp0, *cnt, *xp;
  // rows
  float *dyn, *inp, *igp, *ugp, *fc, *o, *qkv, *att, *y, *o2, *h, *zpart, *pre_c[2], *pre_m[2], *mk, *a_ext;
  float *link_fc, *link_cur, *obj_tmp;
  void *row_planes;  // bf16 hi/mid/lo planes [3][P][C] of a materialised row transform (gate output, pre-head rows)
  void *a_split;     // tcgen05 engine: bf16 hi/mid/lo planes of a_ext
  void *pl[4];       // tcgen05 row engine: bf16 hi/mid/lo planes [3][P][C] handed from one row operator to the next
  void *obj_pl[2];   // planes of a stage's obj_feat (= the next stage's proposal_feat), ping-pong across stages
  // iter loop
  void *mask_pp[2];
  uint32_t *mask_bits;   // bit-mask hand-off between the stages of the fused loop: [frames][N][words per row]
  float *obj_pp[2], *cls_tmp;
  size_t total;
};

// the same shape seen by the two big contractions: one entry of B per FRAME
static VknShape frames_shape(const VknShape &s) {
  VknShape f = s;
  if (s.frames_per_set > 1) f.B = s.B * s.frames_per_set;
  f.frames_per_set = 1;
  return f;
}

static int pool_chunks_max(const VknShape &s) {
  int a = pool_simt_chunks(s);
  int b = pool_tc_chunks(s);
  return a > b ? a : b;
}

static void carve(const VknShape &s, char *base, Layout &L) {
  size_t off = 0;
  auto take = [&](size_t bytes) -> void * {
    off = align_up(off, 256);
    void *p = base ? (void *)(base + off) : nullptr;
    off += bytes;
    return p;
  };
  const size_t P = (size_t)s.B * s.N, C = s.C, F = s.ffn_dim, HW = (size_t)s.H * s.W;
  const size_t f = sizeof(float);
  const size_t fps = s.frames_per_set > 1 ? s.frames_per_set : 1;   // frames per kernel set
  const int nch = pool_chunks_max(frames_shape(s));
  L.pool_part = (float *)take((size_t)nch * P * fps * C * f);
  L.cnt_part = (float *)take((size_t)nch * P * fps * f);
  L.xp0 = (float *)take(P * C * f);
  L.cnt = (float *)take(P * f);
  L.xp = (float *)take(P * C * f);
  L.dyn = (float *)take(P * 2 * C * f);
  L.inp = (float *)take(P * 2 * C * f);
  L.igp = (float *)take(P * C * f);
  L.ugp = (float *)take(P * C * f);
  L.fc = (float *)take(P * C * f);
  L.o = (float *)take(P * C * f);
  L.qkv = (float *)take(P * 3 * C * f);
  L.att = (float *)take(P * C * f);
  L.y = (float *)take(P * C * f);
  L.o2 = (float *)take(P * C * f);
  L.h = (float *)take(P * F * 6);          // fp32 hidden [P,F] or its three bf16 planes [3][P,F]
  L.zpart = (float *)take((size_t)FFN_KSPLIT * P * C * f);
  for (int i = 0; i < 2; ++i) {
    L.pre_c[i] = (float *)take(P * C * f);
    L.pre_m[i] = (float *)take(P * C * f);
  }
  L.mk = (float *)take(P * C * f);
  L.a_ext = (float *)take(P * (C + A_EXT_PAD) * f);
  L.link_fc = (float *)take(P * C * f);
  L.link_cur = (float *)take(P * C * f);
  L.obj_tmp = (float *)take(P * C * f);
  L.row_planes = take((size_t)3 * P * C * 2);
  for (int i = 0; i < 4; ++i) L.pl[i] = take((size_t)3 * P * C * 2);
  for (int i = 0; i < 2; ++i) L.obj_pl[i] = take((size_t)3 * P * C * 2);
  const size_t npad = (size_t)ceil_div(s.N, 128) * 128;
  L.a_split = take((size_t)3 * s.B * npad * C * 2);
  const size_t esz = s.x_dtype == VKN_BF16 ? 2 : 4;
  for (int i = 0; i < 2; ++i) L.mask_pp[i] = take((size_t)s.B * fps * s.N * HW * esz);
  L.mask_bits = (uint32_t *)take((size_t)s.B * fps * s.N * (size_t)(ceil_div((int)HW, 256) * 8) * 4);
  for (int i = 0; i < 2; ++i) L.obj_pp[i] = (float *)take(P * C * f);
  L.cls_tmp = (float *)take(P * (size_t)s.num_classes * f);
  L.total = align_up(off, 256);
}

static int check_shape(const VknShape *s) {
  if (!s) VKN_FAIL(VKN_E_INVALID, "null shape");
  if (s->B < 1 || s->N < 1 || s->H < 1 || s->W < 1) VKN_FAIL(VKN_E_INVALID, "B/N/H/W must be positive");
  if (s->C < 64 || s->C > 256 || s->C % 64 != 0)
    VKN_FAIL(VKN_E_UNSUPPORTED, "C = %d: channels must be a multiple of 64 in [64, 256]", s->C);
  if (s->num_heads < 1 || s->C % s->num_heads != 0 || s->C / s->num_heads > 32)
    VKN_FAIL(VKN_E_UNSUPPORTED, "num_heads %d does not give head_dim <= 32 for C %d", s->num_heads, s->C);
  if (s->with_ffn && (s->ffn_dim < 32 || s->ffn_dim % 32 != 0))
    VKN_FAIL(VKN_E_UNSUPPORTED, "ffn_dim %d must be a positive multiple of 32", s->ffn_dim);
  if (s->num_classes < 1) VKN_FAIL(VKN_E_INVALID, "num_classes must be positive");
  if ((s->x_dtype != VKN_F32 && s->x_dtype != VKN_BF16) || (s->w_dtype != VKN_F32 && s->w_dtype != VKN_BF16))
    VKN_FAIL(VKN_E_INVALID, "bad dtype code");
  if (s->engine < VKN_ENGINE_AUTO || s->engine > VKN_ENGINE_TC) VKN_FAIL(VKN_E_INVALID, "bad engine code");
  if (s->N > 1024) VKN_FAIL(VKN_E_UNSUPPORTED, "N = %d kernels per frame exceeds 1024", s->N);
  if (s->frames_per_set < 0 || s->frames_per_set > 1024) VKN_FAIL(VKN_E_INVALID, "frames_per_set out of range");
  return VKN_OK;
}

struct Ctx {
  VknShape s;
  Layout L;
  cudaStream_t st;
  int P;
  bool use_tc;
  bool rows_tc;     // row operators on the tcgen05 row engine (planes handed between launches)
  bool fuse_ln;     // row engine: LayerNorms fused into the GEMM epilogues (always in chain form; VKN_RG_FUSE_LN=1 forces it for separate launches)
  bool chain_on;    // row engine: the row operators between two attentions run as ONE chain kernel (VKN_CHAIN=0: one launch per operator)
  ChainBuild *chain;  // the chain being assembled (null: every operator launches immediately)
};

// Row-engine operators either launch at once or join the chain under construction; emit_flush launches the chain.
static int emit_gemm(Ctx &c, const LinArgs *p, int n) {
  return c.chain ? chain_add_gemm(c.chain, p, n) : launch_linear_tc(p, n, c.st);
}
static int emit_rowprep(Ctx &c, const RowSrc &src, float *out, int ldo, void *planes, int ldp, long long plane_stride, int M, int K) {
  return c.chain ? chain_add_rowprep(c.chain, src, out, ldo, planes, ldp, plane_stride, K)
                 : launch_rowprep(src, out, ldo, planes, ldp, plane_stride, M, K, c.st);
}
static int emit_flush(Ctx &c) { return c.chain ? chain_launch(c.chain, c.st) : VKN_OK; }

static int make_ctx(const VknShape *s, void *ws, size_t ws_bytes, void *stream, Ctx &c) {
  VKN_TRY(check_shape(s));
  c.s = *s;
  if (!ws) VKN_FAIL(VKN_E_WORKSPACE, "null workspace");
  if (reinterpret_cast<uintptr_t>(ws) & 255) VKN_FAIL(VKN_E_WORKSPACE, "workspace must be 256-byte aligned");
  carve(c.s, (char *)ws, c.L);
  if (c.L.total > ws_bytes)
    VKN_FAIL(VKN_E_WORKSPACE, "workspace too small: %zu bytes given, %zu needed", ws_bytes, c.L.total);
  c.st = (cudaStream_t)stream;
  c.P = s->B * s->N;
  const bool can = tc_supported(frames_shape(c.s));
  if (s->engine == VKN_ENGINE_TC && !can)
    VKN_FAIL(VKN_E_UNSUPPORTED, "tcgen05 engine requested but the shape/dtype does not qualify");
  c.use_tc = (s->engine != VKN_ENGINE_SIMT) && can;
  // Row engine: below a few hundred rows the fused-prologue warp-MMA chain (fewer launches) wins; above, the
  // tcgen05 row GEMM (measured crossover: profiles/, VKN_ROWS_TC_MIN overrides; 0 disables).
  int rows_min = 400;
  if (const char *e = getenv("VKN_ROWS_TC_MIN")) rows_min = atoi(e);
  c.rows_tc = c.use_tc && s->w_dtype == VKN_BF16 && rows_min > 0 && c.P >= rows_min;
  {
    const char *e = getenv("VKN_RG_FUSE_LN");
    c.fuse_ln = e && e[0] == '1';
  }
  {
    const char *e = getenv("VKN_CHAIN");
    c.chain_on = c.rows_tc && !(e && e[0] == '0');
  }
  c.chain = nullptr;
  return VKN_OK;
}

// ---- RowSrc builders ------------------------------------------------------------------------------
static RowSrc src_copy(const float *a, int lda) {
  RowSrc r;
  memset(&r, 0, sizeof(r));
  r.a[0] = a;
  r.lda[0] = lda;
  r.pro = PRO_COPY;
  r.nsum = 1;
  return r;
}
static RowSrc src_ln(const float *a, int lda, const float *g, const float *b, bool relu) {
  RowSrc r = src_copy(a, lda);
  r.pro = relu ? PRO_LN_RELU : PRO_LN;
  r.ln_g[0] = g;
  r.ln_b[0] = b;
  return r;
}
static LinArgs lin(const RowSrc &src, const void *w, int ldw, const float *bias, float *out, int ldo, int M,
                   int N, int K, int epi) {
  LinArgs a;
  memset(&a, 0, sizeof(a));
  a.src = src;
  a.w = w;
  a.ldw = ldw;
  a.bias = bias;
  a.out = out;
  a.ldo = ldo;
  a.M = M;
  a.N = N;
  a.K = K;
  a.epi = epi | (bias ? EPI_BIAS : 0);
  a.ksplit = 1;
  return a;
}
static const void *wrow(const Ctx &c, const void *w, size_t row, size_t ld) {
  const size_t esz = c.s.w_dtype == VKN_BF16 ? 2 : 4;
  return (const char *)w + row * ld * esz;
}

// ---- building blocks --------------------------------------------------------------------------------
// a3+a4: pooled feature with the feat_transform folded out -> xp [P,C]
static int k_pool(Ctx &c, const VknHeadW &w, const void *x, const void *mask, float *xp) {
  int nch = 0;
  const VknShape fs = frames_shape(c.s);     // pooling runs per frame; the reduce folds chunks AND the F frames of a set
  if (c.use_tc) VKN_TRY(launch_pool_tc(fs, x, mask, c.L.pool_part, c.L.cnt_part, &nch, c.st));
  else VKN_TRY(launch_pool_simt(fs, x, mask, c.L.pool_part, c.L.cnt_part, &nch, c.st));
  VKN_TRY(launch_pool_reduce(c.s, c.L.pool_part, c.L.cnt_part, nch, c.L.xp0, c.L.cnt, c.st));
  // x_feat = xp0 . ft_w^T + cnt (x) ft_b      (sum_p M (W x + b) = W (sum_p M x) + (sum_p M) b)
  LinArgs a = lin(src_copy(c.L.xp0, c.s.C), w.ft_w, c.s.C, w.ft_b, xp, c.s.C, c.P, c.s.C, c.s.C, EPI_ROWSCALE);
  a.rowscale = c.L.cnt;
  return launch_linear(&a, 1, c.s.w_dtype, c.st);
}

// a5: KernelUpdator up to the pre-LayerNorm fc_layer output (consumer applies relu(LN_fc_norm(.)))
static int k_update(Ctx &c, const VknUpdatorW &w, const float *xp, const RowSrc &input_src, float *fc_out) {
  const int C = c.s.C, P = c.P;
  LinArgs two[2];
  two[0] = lin(src_copy(xp, C), w.dyn_w, C, w.dyn_b, c.L.dyn, 2 * C, P, 2 * C, C, 0);   // kernel_updator.py:59
  two[1] = lin(input_src, w.inp_w, C, w.inp_b, c.L.inp, 2 * C, P, 2 * C, C, 0);         // :65-66
  VKN_TRY(launch_linear(two, 2, c.s.w_dtype, c.st));
  RowSrc g;
  memset(&g, 0, sizeof(g));
  g.pro = PRO_MUL;                                                                        // :70
  g.nsum = 1;
  g.a[0] = c.L.inp;  g.lda[0] = 2 * C;
  g.a[1] = c.L.dyn;  g.lda[1] = 2 * C;
  two[0] = lin(g, w.ig_w, C, w.ig_b, c.L.igp, C, P, C, C, 0);                             // :74
  two[1] = lin(g, w.ug_w, C, w.ug_b, c.L.ugp, C, P, C, C, 0);                             // :75
  VKN_TRY(launch_linear(two, 2, c.s.w_dtype, c.st));
  RowSrc gt;
  memset(&gt, 0, sizeof(gt));
  gt.pro = PRO_GATE;                                                                      // :76-88
  gt.nsum = 1;
  gt.a[0] = c.L.ugp;       gt.lda[0] = C;      gt.ln_g[0] = w.norm_in_g;    gt.ln_b[0] = w.norm_in_b;
  gt.a[1] = c.L.dyn + C;   gt.lda[1] = 2 * C;  gt.ln_g[1] = w.norm_out_g;   gt.ln_b[1] = w.norm_out_b;
  gt.a[2] = c.L.igp;       gt.lda[2] = C;      gt.ln_g[2] = w.inorm_in_g;   gt.ln_b[2] = w.inorm_in_b;
  gt.a[3] = c.L.inp + C;   gt.lda[3] = 2 * C;  gt.ln_g[3] = w.inorm_out_g;  gt.ln_b[3] = w.inorm_out_b;
  if (c.s.w_dtype == VKN_BF16) {
    // the gate (4 LayerNorms + 2 sigmoids per element) is evaluated once per row and handed to fc_layer as the
    // bf16 planes its tensor-core loop consumes; fused into the Linear it would be redone by all 8 column-block CTAs
    VKN_TRY(launch_rowprep(gt, nullptr, 0, c.L.row_planes, C, (long long)P * C, P, C, c.st));
    RowSrc pl = src_copy((const float *)c.L.row_planes, C);
    pl.pro = PRO_PLANES;
    pl.sum_stride = (long long)P * C;
    LinArgs f = lin(pl, w.fc_w, C, w.fc_b, fc_out, C, P, C, C, 0);                        // :90
    return launch_linear(&f, 1, c.s.w_dtype, c.st);
  }
  LinArgs f = lin(gt, w.fc_w, C, w.fc_b, fc_out, C, P, C, C, 0);                          // :90
  return launch_linear(&f, 1, c.s.w_dtype, c.st);
}

// a6: y = identity + out_proj(MHA(q, kv, kv)); consumer applies LN(norm).  `identity` must be the
// materialised q input; if null it is written to c.L.o as a side output of the q/k/v projection.
static int k_attn(Ctx &c, const VknAttnW &w, const RowSrc &qsrc, const float *identity, const RowSrc *kvsrc,
                  float *y_out) {
  const int C = c.s.C, P = c.P;
  if (kvsrc == nullptr) {
    LinArgs a = lin(qsrc, w.in_w, C, w.in_b, c.L.qkv, 3 * C, P, 3 * C, C, 0);
    if (identity == nullptr) {
      a.side = c.L.o;
      a.ldside = C;
      identity = c.L.o;
    }
    VKN_TRY(launch_linear(&a, 1, c.s.w_dtype, c.st));
  } else {
    LinArgs two[2];
    two[0] = lin(qsrc, w.in_w, C, w.in_b, c.L.qkv, 3 * C, P, C, C, 0);
    if (identity == nullptr) {
      two[0].side = c.L.o;
      two[0].ldside = C;
      identity = c.L.o;
    }
    two[1] = lin(*kvsrc, wrow(c, w.in_w, C, C), C, w.in_b + C, c.L.qkv + C, 3 * C, P, 2 * C, C, 0);
    VKN_TRY(launch_linear(two, 2, c.s.w_dtype, c.st));
  }
  VKN_TRY(launch_attention(c.L.qkv, 3 * C, c.L.qkv + C, 3 * C, c.L.qkv + 2 * C, 3 * C, c.L.att, C, c.s.B, c.s.N,
                           C, c.s.num_heads, c.st));
  LinArgs o = lin(src_copy(c.L.att, C), w.out_w, C, w.out_b, y_out, C, P, C, C, EPI_RES);
  o.res = identity;
  o.ldres = C;
  return launch_linear(&o, 1, c.s.w_dtype, c.st);
}

// a7: FFN.  in_src is the (unmaterialised) input; it is written to c.L.o2 for the residual.
// Returns the pending source: LN_ffn(o2 + b2 + sum_k zpart_k).
static int k_ffn(Ctx &c, const VknFfnW &w, const RowSrc &in_src, RowSrc *pending) {
  const int C = c.s.C, P = c.P, F = c.s.ffn_dim;
  // bf16 weights: the hidden activation is only ever the A operand of the second Linear's tensor-core loop, so the
  // first Linear emits it directly as bf16 hi/mid/lo planes (no fp32 copy, no re-split in the consumer's prologue)
  const bool planes = c.s.w_dtype == VKN_BF16;
  LinArgs a = lin(in_src, w.w1, C, w.b1, c.L.h, F, P, F, C, EPI_RELU);
  a.side = c.L.o2;
  a.ldside = C;
  if (planes) {
    a.epi |= EPI_SPLIT3 | EPI_NOOUT;
    a.split_planes = (__nv_bfloat16 *)c.L.h;
    a.split_B = 1;
    a.split_N = P;
    a.split_Npad = P;
    a.split_C = F;
  }
  VKN_TRY(launch_linear(&a, 1, c.s.w_dtype, c.st));
  RowSrc hsrc = src_copy(c.L.h, F);
  if (planes) {
    hsrc.pro = PRO_PLANES;
    hsrc.sum_stride = (long long)P * F;       // plane stride (elements)
  }
  LinArgs b = lin(hsrc, w.w2, F, nullptr, c.L.zpart, C, P, C, F, 0);
  // enough K-slices to fill ~3 CTAs on every SM (16 x 32 output tiles), at most FFN_KSPLIT, at least 256 of K each
  int ksp = 444 / (ceil_div(C, 32) * ceil_div(P, 16));
  ksp = ksp >= 8 ? 8 : (ksp >= 4 ? 4 : (ksp >= 2 ? 2 : 1));
  while (ksp > 1 && F / ksp < 256) ksp /= 2;
  if (ksp > FFN_KSPLIT) ksp = FFN_KSPLIT;
  b.ksplit = ksp;
  b.out_split_stride = (long long)P * C;
  VKN_TRY(launch_linear(&b, 1, c.s.w_dtype, c.st));
  RowSrc r = src_ln(c.L.zpart, C, w.norm_g, w.norm_b, false);
  r.nsum = ksp;
  r.sum_stride = (long long)P * C;
  r.pbias = w.b2;
  r.pres = c.L.o2;
  r.ldpres = C;
  *pending = r;
  return VKN_OK;
}

// a8: cls / mask FC stacks.  `obj_src` is the pending obj_feat; it is materialised into obj_out
// (side output of the first launch).  mk_out [P,C] = fc_mask(...) ; cls_out [P,ncls].
static int k_heads(Ctx &c, const VknHeadW &w, const RowSrc &obj_src, float *obj_out, float *cls_out,
                   float *mk_out) {
  const int C = c.s.C, P = c.P;
  if (w.num_cls_fcs < 0 || w.num_cls_fcs > VKN_MAX_FCS || w.num_mask_fcs < 0 || w.num_mask_fcs > VKN_MAX_FCS)
    VKN_FAIL(VKN_E_UNSUPPORTED, "num_cls_fcs / num_mask_fcs must be in [0, %d]", VKN_MAX_FCS);
  RowSrc cs = obj_src, ms = obj_src;
  bool need_side = obj_out != nullptr;
  if (c.s.w_dtype == VKN_BF16 && obj_src.pro != PRO_COPY && obj_src.pro != PRO_PLANES) {
    // obj_feat = LN(residual + bias + sum of FFN K-slices): done once per row -> fp32 obj_feat + bf16 planes that both
    // head branches consume, instead of inside every column-block CTA of the first head Linear
    VKN_TRY(launch_rowprep(obj_src, obj_out ? obj_out : c.L.obj_tmp, C, c.L.row_planes, C, (long long)P * C, P, C, c.st));
    RowSrc pl = src_copy((const float *)c.L.row_planes, C);
    pl.pro = PRO_PLANES;
    pl.sum_stride = (long long)P * C;
    cs = pl;
    ms = pl;
    need_side = false;
  }
  const bool with_cls = w.fc_cls_w != nullptr && cls_out != nullptr;   // KernelUpdateHeadVideo(with_cls=False)
  const int ncls_fcs = with_cls ? w.num_cls_fcs : 0;
  const int depth = ncls_fcs > w.num_mask_fcs ? ncls_fcs : w.num_mask_fcs;
  for (int i = 0; i < depth; ++i) {
    LinArgs two[2];
    int n = 0;
    if (i < ncls_fcs) {
      two[n] = lin(cs, w.cls_fc_w[i], C, nullptr, c.L.pre_c[i & 1], C, P, C, C, 0);
      cs = src_ln(c.L.pre_c[i & 1], C, w.cls_ln_g[i], w.cls_ln_b[i], true);
      ++n;
    }
    if (i < w.num_mask_fcs) {
      two[n] = lin(ms, w.mask_fc_w[i], C, nullptr, c.L.pre_m[i & 1], C, P, C, C, 0);
      ms = src_ln(c.L.pre_m[i & 1], C, w.mask_ln_g[i], w.mask_ln_b[i], true);
      ++n;
    }
    if (need_side && i == 0) {   // both start from obj_src at depth 0
      two[0].side = obj_out;
      two[0].ldside = C;
      need_side = false;
    }
    VKN_TRY(launch_linear(two, n, c.s.w_dtype, c.st));
  }
  LinArgs two[2];
  two[0] = lin(ms, w.fc_mask_w, C, w.fc_mask_b, mk_out, C, P, C, C, 0);
  if (with_cls) two[1] = lin(cs, w.fc_cls_w, C, w.fc_cls_b, cls_out, c.s.num_classes, P, c.s.num_classes, C, 0);
  if (need_side) {               // no FC layers at all: the final launch materialises obj_feat
    two[0].side = obj_out;
    two[0].ldside = C;
  }
  return launch_linear(two, with_cls ? 2 : 1, c.s.w_dtype, c.st);
}

// a9: new_mask = mk . (ft_w x + ft_b) = (mk . ft_w) x + mk . ft_b
static int k_maskgemm(Ctx &c, const VknHeadW &w, const void *x, const float *mk, void *out) {
  const int C = c.s.C, P = c.P;
  const int lda = C + A_EXT_PAD;
  LinArgs a = lin(src_copy(mk, C), w.ft_wt_ext, C, nullptr, c.L.a_ext, lda, P, C + 1, C, 0);
  if (c.use_tc) {      // the tcgen05 engine consumes bf16 hi/mid/lo planes: emit them from this epilogue
    a.epi |= EPI_SPLIT3;
    a.split_planes = (__nv_bfloat16 *)c.L.a_split;
    a.split_B = c.s.B;
    a.split_N = c.s.N;
    a.split_Npad = maskgemm_tc_npad(c.s);
    a.split_C = C;
  }
  VKN_TRY(launch_linear(&a, 1, c.s.w_dtype, c.st));
  if (c.use_tc) return launch_maskgemm_tc(c.s, x, c.L.a_ext, lda, c.L.a_split, out, c.st);
  return launch_maskgemm_simt(c.s, x, c.L.a_ext, lda, out, c.st);
}


// ---- the same stage on the tcgen05 row engine (many rows in flight) --------------------------------
// Every row transform is evaluated ONCE per row (row operator / GEMM epilogue) and handed to the next GEMM as bf16
// hi/mid/lo planes, so the GEMMs are plain  planes x W^T  on tensor cores with TMA-fed operands.
static RowSrc src_planes(const void *pl, int ld, long long plane_stride) {
  RowSrc r = src_copy((const float *)pl, ld);
  r.pro = PRO_PLANES;
  r.sum_stride = plane_stride;
  return r;
}
static void out_planes(LinArgs &a, void *pl, int rows, int ld) {
  a.epi |= EPI_SPLIT3;
  a.split_planes = (__nv_bfloat16 *)pl;
  a.split_B = 1;
  a.split_N = rows;
  a.split_Npad = rows;
  a.split_C = ld;
}

static int stage_planes(Ctx &c, const VknHeadW &w, const void *x, const float *pf, const void *pf_planes, const void *mask,
                        const float *x_feat_in, float *cls, void *new_mask, float *obj, float *x_feat_out,
                        void *obj_planes_out, const uint32_t *mask_bits_in = nullptr, uint32_t *mask_bits_out = nullptr) {
  const int C = c.s.C, P = c.P, F = c.s.ffn_dim;
  const long long PS = (long long)P * C;
  void *PLA = c.L.pl[0], *PLB = c.L.pl[1], *PLC = c.L.pl[2], *PLD = c.L.pl[3];
  if (obj_planes_out == nullptr) obj_planes_out = c.L.obj_pl[0];
  // Chain form: everything between the pooled feature and the attention, and between the attention and the mask conv,
  // is ONE launch each (a CTA per 128-row tile walks the operators by itself); LayerNorms live in the GEMM epilogues.
  const bool fuse_ln = c.fuse_ln || c.chain_on;
  if (c.chain_on) c.chain = chain_begin(P);
  // a3+a4 (+a2 folded): pooled feature -> x_feat fp32 + planes (PLB)
  if (x_feat_in == nullptr) {
    float *xf = x_feat_out ? x_feat_out : c.L.xp;
    int nch = 0;
    const VknShape fs = frames_shape(c.s);
    VKN_TRY(launch_pool_tc(fs, x, mask, c.L.pool_part, c.L.cnt_part, &nch, c.st, mask_bits_in));
    VKN_TRY(launch_pool_reduce(c.s, c.L.pool_part, c.L.cnt_part, nch, c.L.xp0, c.L.cnt, c.st, PLA));
    // the fp32 copy of x_feat is only written when the caller asked for it (VideoKernelUpdateHead returns it)
    LinArgs a = lin(src_planes(PLA, C, PS), w.ft_w, C, w.ft_b, x_feat_out ? xf : nullptr, C, P, C, C,
                    EPI_ROWSCALE | (x_feat_out ? 0 : EPI_NOOUT));
    a.rowscale = c.L.cnt;
    out_planes(a, PLB, P, C);
    VKN_TRY(emit_gemm(c, &a, 1));
  } else {
    float *copy_to = (x_feat_out && x_feat_out != x_feat_in) ? x_feat_out : nullptr;
    VKN_TRY(launch_rowprep(src_copy(x_feat_in, C), copy_to, C, PLB, C, PS, P, C, c.st));
  }
  if (pf_planes == nullptr) {
    VKN_TRY(launch_rowprep(src_copy(pf, C), nullptr, 0, PLC, C, PS, P, C, c.st));
    pf_planes = PLC;
  }
  // a5 KernelUpdator (kernel_updator.py:56-94)
  const VknUpdatorW &u = w.upd;
  LinArgs two[2];
  if (c.chain) {
    // Chain form: the gate arithmetic lives in the GEMM epilogues (no row passes, no fp32 round trips of the gate inputs):
    //   param_in = dyn_w[:C] xf + b             input_in . param_in -> planes (gate_feats)            :59-70
    //   param_out, input_out                     -> norm_out / input_norm_out in the epilogue          :86-87
    //   U = sigmoid(norm_in(update_gate(g))) * param_out;  features = sigmoid(input_norm_in(input_gate(g))) * input_out + U   :74-88
    LinArgs four[4];
    four[0] = lin(src_planes(PLB, C, PS), u.dyn_w, C, u.dyn_b, c.L.dyn, 2 * C, P, C, C, 0);                         // param_in
    four[1] = lin(src_planes(PLB, C, PS), wrow(c, u.dyn_w, C, C), C, u.dyn_b + C, c.L.dyn + C, 2 * C, P, C, C, EPI_LN);   // param_out
    four[1].ln_g = u.norm_out_g;
    four[1].ln_b = u.norm_out_b;
    four[2] = lin(src_planes(pf_planes, C, PS), wrow(c, u.inp_w, C, C), C, u.inp_b + C, c.L.inp + C, 2 * C, P, C, C, EPI_LN);   // input_out
    four[2].ln_g = u.inorm_out_g;
    four[2].ln_b = u.inorm_out_b;
    four[3] = lin(src_planes(pf_planes, C, PS), u.inp_w, C, u.inp_b, nullptr, C, P, C, C, EPI_MUL | EPI_NOOUT);     // input_in . param_in
    four[3].mul = c.L.dyn;
    four[3].ldmul = 2 * C;
    out_planes(four[3], PLA, P, C);
    VKN_TRY(emit_gemm(c, four, 4));
    two[0] = lin(src_planes(PLA, C, PS), u.ug_w, C, u.ug_b, c.L.ugp, C, P, C, C, EPI_LN | EPI_SIGMOID | EPI_MUL);    // :75, :77-86
    two[0].ln_g = u.norm_in_g;
    two[0].ln_b = u.norm_in_b;
    two[0].mul = c.L.dyn + C;
    two[0].ldmul = 2 * C;
    two[1] = lin(src_planes(PLA, C, PS), u.ig_w, C, u.ig_b, nullptr, C, P, C, C,
                 EPI_LN | EPI_SIGMOID | EPI_MUL | EPI_ADD2 | EPI_NOOUT);                                                // :74, :76-88
    two[1].ln_g = u.inorm_in_g;
    two[1].ln_b = u.inorm_in_b;
    two[1].mul = c.L.inp + C;
    two[1].ldmul = 2 * C;
    two[1].add2 = c.L.ugp;
    two[1].ldadd2 = C;
    out_planes(two[1], PLB, P, C);
    VKN_TRY(emit_gemm(c, two, 2));
    LinArgs f = lin(src_planes(PLB, C, PS), u.fc_w, C, u.fc_b, c.L.o, C, P, C, C, EPI_LN | EPI_RELU);              // :90-92
    f.ln_g = u.fc_norm_g;
    f.ln_b = u.fc_norm_b;
    out_planes(f, PLA, P, C);
    VKN_TRY(emit_gemm(c, &f, 1));
  } else {
  two[0] = lin(src_planes(PLB, C, PS), u.dyn_w, C, u.dyn_b, c.L.dyn, 2 * C, P, 2 * C, C, 0);          // :59
  two[1] = lin(src_planes(pf_planes, C, PS), u.inp_w, C, u.inp_b, c.L.inp, 2 * C, P, 2 * C, C, 0);    // :65-66
  VKN_TRY(emit_gemm(c, two, 2));
  RowSrc g;
  memset(&g, 0, sizeof(g));
  g.pro = PRO_MUL;                                                                                     // :70
  g.nsum = 1;
  g.a[0] = c.L.inp;  g.lda[0] = 2 * C;
  g.a[1] = c.L.dyn;  g.lda[1] = 2 * C;
  VKN_TRY(emit_rowprep(c, g, nullptr, 0, PLA, C, PS, P, C));
  two[0] = lin(src_planes(PLA, C, PS), u.ig_w, C, u.ig_b, c.L.igp, C, P, C, C, 0);                    // :74
  two[1] = lin(src_planes(PLA, C, PS), u.ug_w, C, u.ug_b, c.L.ugp, C, P, C, C, 0);                    // :75
  VKN_TRY(emit_gemm(c, two, 2));
  RowSrc gt;
  memset(&gt, 0, sizeof(gt));
  gt.pro = PRO_GATE;                                                                                   // :76-88
  gt.nsum = 1;
  gt.a[0] = c.L.ugp;       gt.lda[0] = C;      gt.ln_g[0] = u.norm_in_g;    gt.ln_b[0] = u.norm_in_b;
  gt.a[1] = c.L.dyn + C;   gt.lda[1] = 2 * C;  gt.ln_g[1] = u.norm_out_g;   gt.ln_b[1] = u.norm_out_b;
  gt.a[2] = c.L.igp;       gt.lda[2] = C;      gt.ln_g[2] = u.inorm_in_g;   gt.ln_b[2] = u.inorm_in_b;
  gt.a[3] = c.L.inp + C;   gt.lda[3] = 2 * C;  gt.ln_g[3] = u.inorm_out_g;  gt.ln_b[3] = u.inorm_out_b;
  VKN_TRY(emit_rowprep(c, gt, nullptr, 0, PLB, C, PS, P, C));
  if (fuse_ln) {
    // :90-92  fc_layer -> fc_norm -> ReLU in one launch (LayerNorm fused in the GEMM epilogue): o (fp32) + planes PLA
    LinArgs f = lin(src_planes(PLB, C, PS), u.fc_w, C, u.fc_b, c.L.o, C, P, C, C, EPI_LN | EPI_RELU);
    f.ln_g = u.fc_norm_g;
    f.ln_b = u.fc_norm_b;
    out_planes(f, PLA, P, C);
    VKN_TRY(emit_gemm(c, &f, 1));
  } else {
    LinArgs f = lin(src_planes(PLB, C, PS), u.fc_w, C, u.fc_b, c.L.fc, C, P, C, C, 0);                  // :90
    VKN_TRY(emit_gemm(c, &f, 1));
    VKN_TRY(emit_rowprep(c, src_ln(c.L.fc, C, u.fc_norm_g, u.fc_norm_b, true), c.L.o, C, PLA, C, PS, P, C));   // :91-92
  }
  }
  // a6 MHSA + LN (kernel_update_head.py:204-208)
  LinArgs qkv = lin(src_planes(PLA, C, PS), w.attn.in_w, C, w.attn.in_b, c.L.qkv, 3 * C, P, 3 * C, C, 0);
  VKN_TRY(emit_gemm(c, &qkv, 1));
  VKN_TRY(emit_flush(c));
  VKN_TRY(launch_attention(c.L.qkv, 3 * C, c.L.qkv + C, 3 * C, c.L.qkv + 2 * C, 3 * C, nullptr, C, c.s.B, c.s.N, C,
                           c.s.num_heads, c.st, PLB, PS));
  float *obj_dst = obj ? obj : c.L.obj_tmp;
  if (fuse_ln) {
    // out-projection + residual + attention_norm in one launch (:206-208)
    LinArgs op = lin(src_planes(PLB, C, PS), w.attn.out_w, C, w.attn.out_b, c.s.with_ffn ? c.L.o2 : obj_dst, C, P, C, C,
                     EPI_RES | EPI_LN);
    op.res = c.L.o;
    op.ldres = C;
    op.ln_g = w.attn.norm_g;
    op.ln_b = w.attn.norm_b;
    out_planes(op, c.s.with_ffn ? PLA : obj_planes_out, P, C);
    VKN_TRY(emit_gemm(c, &op, 1));
  } else {
    LinArgs op = lin(src_planes(PLB, C, PS), w.attn.out_w, C, w.attn.out_b, c.L.y, C, P, C, C, EPI_RES);
    op.res = c.L.o;
    op.ldres = C;
    VKN_TRY(emit_gemm(c, &op, 1));
    RowSrc an = src_ln(c.L.y, C, w.attn.norm_g, w.attn.norm_b, false);
    if (c.s.with_ffn) VKN_TRY(emit_rowprep(c, an, c.L.o2, C, PLA, C, PS, P, C));
    else VKN_TRY(emit_rowprep(c, an, obj_dst, C, obj_planes_out, C, PS, P, C));
  }
  if (c.s.with_ffn) {
    // a7 FFN + LN (:214-215)
    LinArgs f1 = lin(src_planes(PLA, C, PS), w.ffn.w1, C, w.ffn.b1, nullptr, F, P, F, C, EPI_RELU | EPI_NOOUT);
    out_planes(f1, c.L.h, P, F);
    VKN_TRY(emit_gemm(c, &f1, 1));
    if (c.chain) {
      // second Linear over the whole K in one tile pass: + b2 + residual, ffn_norm in the epilogue -> obj_feat + its planes
      LinArgs f2 = lin(src_planes(c.L.h, F, (long long)P * F), w.ffn.w2, F, w.ffn.b2, obj_dst, C, P, C, F, EPI_RES | EPI_LN);
      f2.res = c.L.o2;
      f2.ldres = C;
      f2.ln_g = w.ffn.norm_g;
      f2.ln_b = w.ffn.norm_b;
      out_planes(f2, obj_planes_out, P, C);
      VKN_TRY(emit_gemm(c, &f2, 1));
    } else {
    LinArgs f2 = lin(src_planes(c.L.h, F, (long long)P * F), w.ffn.w2, F, nullptr, c.L.zpart, C, P, C, F, 0);
    const int nk = ceil_div(F, 64);
    int ksp = 148 / (ceil_div(P, 128) * ceil_div(C, 256));      // K slices: fill the SMs with 128 x 256 tiles
    ksp = ksp >= 8 ? 8 : (ksp >= 4 ? 4 : (ksp >= 2 ? 2 : 1));
    while (ksp > 1 && (nk % ksp != 0 || nk / ksp < 4)) ksp /= 2;
    f2.ksplit = ksp;
    f2.out_split_stride = PS;
    VKN_TRY(emit_gemm(c, &f2, 1));
    RowSrc r = src_ln(c.L.zpart, C, w.ffn.norm_g, w.ffn.norm_b, false);
    r.nsum = ksp;
    r.sum_stride = PS;
    r.pbias = w.ffn.b2;
    r.pres = c.L.o2;
    r.ldpres = C;
    VKN_TRY(emit_rowprep(c, r, obj_dst, C, obj_planes_out, C, PS, P, C));
    }
  }
  // a8 heads (:217-227)
  if (w.num_cls_fcs < 0 || w.num_cls_fcs > VKN_MAX_FCS || w.num_mask_fcs < 0 || w.num_mask_fcs > VKN_MAX_FCS)
    VKN_FAIL(VKN_E_UNSUPPORTED, "num_cls_fcs / num_mask_fcs must be in [0, %d]", VKN_MAX_FCS);
  const bool with_cls = w.fc_cls_w != nullptr && cls != nullptr;
  const int ncls_fcs = with_cls ? w.num_cls_fcs : 0;
  const int depth = ncls_fcs > w.num_mask_fcs ? ncls_fcs : w.num_mask_fcs;
  const void *cs = obj_planes_out, *ms = obj_planes_out;
  for (int i = 0; i < depth; ++i) {
    int n = 0;
    if (!fuse_ln) {
      if (i < ncls_fcs) two[n++] = lin(src_planes(cs, C, PS), w.cls_fc_w[i], C, nullptr, c.L.pre_c[0], C, P, C, C, 0);
      if (i < w.num_mask_fcs) two[n++] = lin(src_planes(ms, C, PS), w.mask_fc_w[i], C, nullptr, c.L.pre_m[0], C, P, C, C, 0);
      VKN_TRY(emit_gemm(c, two, n));
      if (i < ncls_fcs) {
        VKN_TRY(emit_rowprep(c, src_ln(c.L.pre_c[0], C, w.cls_ln_g[i], w.cls_ln_b[i], true), nullptr, 0, PLA, C, PS, P, C));
        cs = PLA;
      }
      if (i < w.num_mask_fcs) {
        VKN_TRY(emit_rowprep(c, src_ln(c.L.pre_m[0], C, w.mask_ln_g[i], w.mask_ln_b[i], true), nullptr, 0, PLB, C, PS, P, C));
        ms = PLB;
      }
      continue;
    }
    // Linear (no bias) -> LN -> ReLU per branch, LayerNorm fused in the epilogue, planes only.  In-place plane buffers
    // are safe: a tile spans whole rows and its epilogue starts after all of its MMAs have read those rows.
    if (i < ncls_fcs) {
      two[n] = lin(src_planes(cs, C, PS), w.cls_fc_w[i], C, nullptr, nullptr, C, P, C, C, EPI_LN | EPI_RELU | EPI_NOOUT);
      two[n].ln_g = w.cls_ln_g[i];
      two[n].ln_b = w.cls_ln_b[i];
      out_planes(two[n++], PLA, P, C);
    }
    if (i < w.num_mask_fcs) {
      two[n] = lin(src_planes(ms, C, PS), w.mask_fc_w[i], C, nullptr, nullptr, C, P, C, C, EPI_LN | EPI_RELU | EPI_NOOUT);
      two[n].ln_g = w.mask_ln_g[i];
      two[n].ln_b = w.mask_ln_b[i];
      out_planes(two[n++], PLB, P, C);
    }
    VKN_TRY(emit_gemm(c, two, n));
    if (i < ncls_fcs) cs = PLA;
    if (i < w.num_mask_fcs) ms = PLB;
  }
  two[0] = lin(src_planes(ms, C, PS), w.fc_mask_w, C, w.fc_mask_b, nullptr, C, P, C, C, EPI_NOOUT);   // only its planes are consumed
  out_planes(two[0], PLD, P, C);
  if (with_cls) two[1] = lin(src_planes(cs, C, PS), w.fc_cls_w, C, w.fc_cls_b, cls, c.s.num_classes, P, c.s.num_classes, C, 0);
  VKN_TRY(emit_gemm(c, two, with_cls ? 2 : 1));
  if (new_mask == nullptr && mask_bits_out == nullptr) {
    VKN_TRY(emit_flush(c));
    c.chain = nullptr;
    return VKN_OK;
  }
  // a9 (+a2 folded): a = mk . ft_w (planes for the mask conv), bias column mk . ft_b
  const int lda = C + A_EXT_PAD;
  // the mask conv reads the folded kernels as planes and, of the fp32 rows, only the bias column
  // (fp16 mode of the persistent mask conv: the folded kernels as two fp16 planes instead of three bf16 planes)
  const bool f16 = maskgemm_tc_planes_f16(c.s);
  two[0] = lin(src_planes(PLD, C, PS), w.ft_wt_ext, C, nullptr, nullptr, lda, P, C, C, EPI_SPLIT3 | EPI_NOOUT | (f16 ? EPI_SPLIT2H : 0));
  two[0].split_planes = (__nv_bfloat16 *)c.L.a_split;
  two[0].split_B = c.s.B;
  two[0].split_N = c.s.N;
  two[0].split_Npad = maskgemm_tc_npad(c.s);
  two[0].split_C = C;
  two[1] = lin(src_planes(PLD, C, PS), wrow(c, w.ft_wt_ext, C, C), C, nullptr, c.L.a_ext + C, lda, P, 1, C, 0);
  VKN_TRY(emit_gemm(c, two, 2));
  VKN_TRY(emit_flush(c));
  c.chain = nullptr;
  return launch_maskgemm_tc(c.s, x, c.L.a_ext, lda, c.L.a_split, new_mask, c.st, mask_bits_out, f16);
}

static int stage(Ctx &c, const VknHeadW &w, const void *x, const float *pf, const void *mask,
                 const float *x_feat_in, float *cls, void *new_mask, float *obj, float *x_feat_out,
                 const void *pf_planes = nullptr, void *obj_planes_out = nullptr, const uint32_t *mask_bits_in = nullptr,
                 uint32_t *mask_bits_out = nullptr) {
  if (c.rows_tc)
    return stage_planes(c, w, x, pf, pf_planes, mask, x_feat_in, cls, new_mask, obj, x_feat_out, obj_planes_out, mask_bits_in,
                        mask_bits_out);
  if (mask_bits_in || mask_bits_out) VKN_FAIL(VKN_E_INVALID, "bit-mask hand-off is a row-engine (frame batch) path");
  const int C = c.s.C;
  if (frame_chain_supported(c.s, w)) {
    // a frame or two (the online VPS operating point): pooling, then the cluster chain (framechain.cu: every row operator of
    // the stage in two launches, activations handed between the Linears through distributed shared memory), then the mask conv
    if (x_feat_in == nullptr) {
      int nch = 0;
      const VknShape fs = frames_shape(c.s);
      if (c.use_tc) VKN_TRY(launch_pool_tc(fs, x, mask, c.L.pool_part, c.L.cnt_part, &nch, c.st));
      else VKN_TRY(launch_pool_simt(fs, x, mask, c.L.pool_part, c.L.cnt_part, &nch, c.st));
      VKN_TRY(launch_pool_reduce(c.s, c.L.pool_part, c.L.cnt_part, nch, c.L.xp0, c.L.cnt, c.st));
    } else if (x_feat_out && x_feat_out != x_feat_in) {
      VKN_CUDA_OK(cudaMemcpyAsync(x_feat_out, x_feat_in, (size_t)c.P * C * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
    }
    const int lda = C + A_EXT_PAD;
    VKN_TRY(launch_frame_chain(c.s, w, c.L.xp0, c.L.cnt, x_feat_in, pf, x_feat_in ? nullptr : x_feat_out, c.L.o, c.L.qkv, obj, cls, c.L.a_ext, lda, c.L.a_split,
                               maskgemm_tc_npad(c.s), c.st));
    if (!new_mask) return VKN_OK;
    if (c.use_tc) return launch_maskgemm_tc(c.s, x, c.L.a_ext, lda, c.L.a_split, new_mask, c.st);
    return launch_maskgemm_simt(c.s, x, c.L.a_ext, lda, new_mask, c.st);
  }
  const float *xp = x_feat_in;
  if (xp == nullptr) {
    float *dst = x_feat_out ? x_feat_out : c.L.xp;
    VKN_TRY(k_pool(c, w, x, mask, dst));
    xp = dst;
  } else if (x_feat_out && x_feat_out != x_feat_in) {
    VKN_CUDA_OK(cudaMemcpyAsync(x_feat_out, x_feat_in, (size_t)c.P * C * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
  }
  VKN_TRY(k_update(c, w.upd, xp, src_copy(pf, C), c.L.fc));
  RowSrc o = src_ln(c.L.fc, C, w.upd.fc_norm_g, w.upd.fc_norm_b, true);       // kernel_updator.py:91-92
  VKN_TRY(k_attn(c, w.attn, o, nullptr, nullptr, c.L.y));                     // kernel_update_head.py:206
  RowSrc pend = src_ln(c.L.y, C, w.attn.norm_g, w.attn.norm_b, false);
  if (c.s.with_ffn) VKN_TRY(k_ffn(c, w.ffn, pend, &pend));                    // :214-215
  VKN_TRY(k_heads(c, w, pend, obj, cls, c.L.mk));                             // :217-227
  if (new_mask) VKN_TRY(k_maskgemm(c, w, x, c.L.mk, new_mask));               // :247-260
  return VKN_OK;
}

}  // namespace vkn

using namespace vkn;

extern "C" {

int vkn_version(void) { return VKN_VERSION; }
const char *vkn_last_error(void) { return g_err; }

const char *vkn_kernel_names(void) {
  return "vkn_pool_simt_kernel\nvkn_pool_reduce_kernel\nvkn_pool_reduce_flat_kernel\nvkn_maskgemm_simt_kernel\nvkn_linear_kernel\n"
         "vkn_rowop_kernel\nvkn_attention_kernel\nvkn_attention4_kernel\nvkn_attention_tc_kernel\nvkn_pool_tc_kernel\nvkn_maskgemm_tc_kernel\nvkn_maskgemm_tc_persist_kernel\nvkn_maskgemm_tc_wide_kernel\nvkn_pack_kernels_kernel\nvkn_rowgemm_tc_kernel\nvkn_chain_tc_kernel\nvkn_frame_chain_a_kernel\nvkn_frame_chain_b_kernel\nvkn_frame_chain_pack_kernel\nvkn_match_cost_partial_kernel\nvkn_match_cost_final_kernel\nvkn_panoptic_owner_kernel\nvkn_panoptic_segments_kernel\nvkn_panoptic_paint_kernel\nvkn_mask_boxes_kernel\nvkn_track_match_kernel\nvkn_rescale_masks_kernel";
}

unsigned long long vkn_launch_count(void) { return g_launches; }

int vkn_debug_timestamps(unsigned long long *buf, size_t n_u64) {
  g_ts = buf;
  g_ts_cap = buf ? n_u64 : 0;
  g_ts_launch = 0;
  return (int)(DBG_STRIDE);
}

int vkn_profile_begin(void) {
  if (!g_prof.created) {
    for (int i = 0; i <= PROF_MAX; ++i) VKN_CUDA_OK(cudaEventCreate(&g_prof.ev[i]));
    g_prof.created = true;
  }
  g_prof.n = 0;
  g_prof.on = true;
  return VKN_OK;
}

int vkn_profile_end(const char **names, float *ms, int max_entries, int *count) {
  if (!g_prof.on) VKN_FAIL(VKN_E_INVALID, "vkn_profile_end without vkn_profile_begin");
  g_prof.on = false;
  if (!names || !ms || !count) VKN_FAIL(VKN_E_INVALID, "vkn_profile_end: null argument");
  const int n = g_prof.n;
  if (n > 0) {
    VKN_CUDA_OK(cudaEventRecord(g_prof.ev[n], g_prof.st));
    VKN_CUDA_OK(cudaEventSynchronize(g_prof.ev[n]));
  }
  *count = n < max_entries ? n : max_entries;
  for (int i = 0; i < *count; ++i) {
    names[i] = g_prof.names[i];
    VKN_CUDA_OK(cudaEventElapsedTime(&ms[i], g_prof.ev[i], g_prof.ev[i + 1]));
  }
  return VKN_OK;
}

int vkn_frame_chain_pack_bytes(const VknShape *shape, const VknHeadW *w, size_t *bytes) {
  VKN_TRY(check_shape(shape));
  if (!w || !bytes) VKN_FAIL(VKN_E_INVALID, "vkn_frame_chain_pack_bytes: null argument");
  *bytes = frame_chain_pack_bytes(*shape, *w);
  return VKN_OK;
}

int vkn_frame_chain_pack(const VknShape *shape, const VknHeadW *w, void *out, size_t bytes, void *stream) {
  VKN_TRY(check_shape(shape));
  if (!w) VKN_FAIL(VKN_E_INVALID, "vkn_frame_chain_pack: null argument");
  return launch_frame_chain_pack(*shape, *w, out, bytes, (cudaStream_t)stream);
}

int vkn_match_cost_workspace_bytes(int N, int M, int HW, size_t *bytes) {
  if (!bytes || N < 1 || M < 1 || HW < 1) VKN_FAIL(VKN_E_INVALID, "vkn_match_cost_workspace_bytes: bad argument");
  *bytes = match_cost_workspace_bytes(N, M, HW);
  return VKN_OK;
}

int vkn_match_cost(const float *mask_logits, const float *cls_logits, const float *gt_masks, const int64_t *gt_labels, int N,
                   int M, int HW, int ncls, const float *params, float *cost, void *workspace, size_t workspace_bytes,
                   void *stream) {
  return launch_match_cost(mask_logits, cls_logits, gt_masks, (const long long *)gt_labels, N, M, HW, ncls, params, cost, workspace,
                           workspace_bytes, (cudaStream_t)stream);
}

int vkn_workspace_bytes(const VknShape *shape, size_t *bytes) {
  VKN_TRY(check_shape(shape));
  if (!bytes) VKN_FAIL(VKN_E_INVALID, "null bytes pointer");
  Layout L;
  carve(*shape, nullptr, L);
  *bytes = L.total;
  return VKN_OK;
}

int vkn_mask_pool(const VknShape *s, const VknHeadW *w, const void *x, const void *mask_preds, float *x_feat,
                  void *workspace, size_t workspace_bytes, void *stream) {
  Ctx c;
  VKN_TRY(make_ctx(s, workspace, workspace_bytes, stream, c));
  if (!w || !x || !mask_preds || !x_feat) VKN_FAIL(VKN_E_INVALID, "vkn_mask_pool: null argument");
  return k_pool(c, *w, x, mask_preds, x_feat);
}

int vkn_kernel_update(const VknShape *s, const VknUpdatorW *w, const float *x_feat, const float *proposal_feat,
                      float *out, void *workspace, size_t workspace_bytes, void *stream) {
  Ctx c;
  VKN_TRY(make_ctx(s, workspace, workspace_bytes, stream, c));
  if (!w || !x_feat || !proposal_feat || !out) VKN_FAIL(VKN_E_INVALID, "vkn_kernel_update: null argument");
  VKN_TRY(k_update(c, *w, x_feat, src_copy(proposal_feat, s->C), c.L.fc));
  return launch_rowop(src_ln(c.L.fc, s->C, w->fc_norm_g, w->fc_norm_b, true), out, s->C, c.P, s->C, c.st);
}

int vkn_mhsa_ln(const VknShape *s, const VknAttnW *w, const float *q_in, const float *kv_in, float *out,
                void *workspace, size_t workspace_bytes, void *stream) {
  Ctx c;
  VKN_TRY(make_ctx(s, workspace, workspace_bytes, stream, c));
  if (!w || !q_in || !out) VKN_FAIL(VKN_E_INVALID, "vkn_mhsa_ln: null argument");
  RowSrc kv = src_copy(kv_in, s->C);
  VKN_TRY(k_attn(c, *w, src_copy(q_in, s->C), q_in, (kv_in && kv_in != q_in) ? &kv : nullptr, c.L.y));
  return launch_rowop(src_ln(c.L.y, s->C, w->norm_g, w->norm_b, false), out, s->C, c.P, s->C, c.st);
}

int vkn_ffn_ln(const VknShape *s, const VknFfnW *w, const float *in, float *out, void *workspace,
               size_t workspace_bytes, void *stream) {
  Ctx c;
  VKN_TRY(make_ctx(s, workspace, workspace_bytes, stream, c));
  if (!w || !in || !out) VKN_FAIL(VKN_E_INVALID, "vkn_ffn_ln: null argument");
  RowSrc pend;
  VKN_TRY(k_ffn(c, *w, src_copy(in, s->C), &pend));
  return launch_rowop(pend, out, s->C, c.P, s->C, c.st);
}

int vkn_heads(const VknShape *s, const VknHeadW *w, const float *obj_feat, float *cls_score, float *mask_kernel,
              void *workspace, size_t workspace_bytes, void *stream) {
  Ctx c;
  VKN_TRY(make_ctx(s, workspace, workspace_bytes, stream, c));
  if (!w || !obj_feat || !cls_score || !mask_kernel) VKN_FAIL(VKN_E_INVALID, "vkn_heads: null argument");
  return k_heads(c, *w, src_copy(obj_feat, s->C), nullptr, cls_score, mask_kernel);
}

int vkn_mask_gemm(const VknShape *s, const VknHeadW *w, const void *x, const float *mask_kernel,
                  void *new_mask_preds, void *workspace, size_t workspace_bytes, void *stream) {
  Ctx c;
  VKN_TRY(make_ctx(s, workspace, workspace_bytes, stream, c));
  if (!w || !x || !mask_kernel || !new_mask_preds) VKN_FAIL(VKN_E_INVALID, "vkn_mask_gemm: null argument");
  return k_maskgemm(c, *w, x, mask_kernel, new_mask_preds);
}

int vkn_stage_forward(const VknShape *s, const VknHeadW *w, const void *x, const float *proposal_feat,
                      const void *mask_preds, const float *x_feat_in, float *cls_score, void *new_mask_preds,
                      float *obj_feat, float *x_feat_out, void *workspace, size_t workspace_bytes, void *stream) {
  Ctx c;
  VKN_TRY(make_ctx(s, workspace, workspace_bytes, stream, c));
  if (!w || !x || !proposal_feat || !obj_feat || (!mask_preds && !x_feat_in))
    VKN_FAIL(VKN_E_INVALID, "vkn_stage_forward: null argument");
  return stage(c, *w, x, proposal_feat, mask_preds, x_feat_in, cls_score, new_mask_preds, obj_feat, x_feat_out);
}

int vkn_iter_forward(const VknShape *s, const VknHeadW *stages, int num_stages, const void *x,
                     const float *proposal_feat, const void *mask_preds, float *cls_score, void *new_mask_preds,
                     float *obj_feat, void *workspace, size_t workspace_bytes, void *stream) {
  Ctx c;
  VKN_TRY(make_ctx(s, workspace, workspace_bytes, stream, c));
  if (!stages || num_stages < 1 || !x || !proposal_feat || !mask_preds || !cls_score || !new_mask_preds || !obj_feat)
    VKN_FAIL(VKN_E_INVALID, "vkn_iter_forward: null argument");
  const float *pf = proposal_feat;
  const void *mk = mask_preds;
  // Inside the loop the only consumer of an inner stage's masks is the next stage's hard threshold (the reference's
  // simple_test keeps only the last stage's, knet/det/kernel_iter_head.py:246-253): frame batches hand over the
  // thresholded BIT per (kernel, pixel) instead of bf16 logits -- 3.5 MB -> 0.22 MB per frame and stage on both sides.
  bool use_bits = c.rows_tc && maskgemm_tc_persistent(c.s);
  if (const char *e = getenv("VKN_LOOP_BITS")) use_bits = use_bits && e[0] != '0';
  const uint32_t *bits_in = nullptr;
  for (int i = 0; i < num_stages; ++i) {
    const bool last = i == num_stages - 1;
    float *obj_o = last ? obj_feat : c.L.obj_pp[i & 1];
    void *mask_o = last ? new_mask_preds : (use_bits ? nullptr : c.L.mask_pp[i & 1]);
    uint32_t *bits_o = (!last && use_bits) ? c.L.mask_bits : nullptr;
    float *cls_o = last ? cls_score : c.L.cls_tmp;
    VKN_TRY(stage(c, stages[i], x, pf, mk, nullptr, cls_o, mask_o, obj_o, nullptr,
                  i > 0 ? c.L.obj_pl[(i - 1) & 1] : nullptr, c.L.obj_pl[i & 1], bits_in, bits_o));
    pf = obj_o;
    mk = mask_o;
    bits_in = bits_o;
  }
  return VKN_OK;
}

int vkn_init_proposals(const VknShape *s, const float *init_w, const float *init_b, const void *loc_feats,
                       const void *x_feats, void *mask_preds, float *proposal_feats, void *workspace,
                       size_t workspace_bytes, void *stream) {
  Ctx c;
  VKN_TRY(make_ctx(s, workspace, workspace_bytes, stream, c));
  if (!init_w || !loc_feats || !x_feats || !mask_preds || !proposal_feats)
    VKN_FAIL(VKN_E_INVALID, "vkn_init_proposals: null argument");
  if (s->frames_per_set > 1) VKN_FAIL(VKN_E_INVALID, "vkn_init_proposals: frames_per_set must be 1 (B counts frames)");
  const int C = s->C, N = s->N, lda = C + A_EXT_PAD;
  // (1) the static kernels as the mask-conv operand: a_ext rows [init_w | init_b], plus the bf16 planes for tcgen05
  VknShape one = c.s;          // ONE kernel set shared by all B frames
  one.B = 1;
  one.frames_per_set = s->B;
  VKN_TRY(launch_pack_kernels(init_w, init_b, N, C, c.L.a_ext, lda, c.use_tc ? c.L.a_split : nullptr,
                              maskgemm_tc_npad(one), c.st));
  // (2) mask_preds = init_kernels(loc_feats)
  if (c.use_tc) VKN_TRY(launch_maskgemm_tc(one, loc_feats, c.L.a_ext, lda, c.L.a_split, mask_preds, c.st));
  else VKN_TRY(launch_maskgemm_simt(one, loc_feats, c.L.a_ext, lda, mask_preds, c.st));
  // (3) obj = hard-mask pooling of x_feats (threshold 0.5 <=> logit 0), per frame
  VknShape fs = c.s;
  fs.mask_thr_logit = 0.f;
  int nch = 0;
  if (c.use_tc) VKN_TRY(launch_pool_tc(fs, x_feats, mask_preds, c.L.pool_part, c.L.cnt_part, &nch, c.st));
  else VKN_TRY(launch_pool_simt(fs, x_feats, mask_preds, c.L.pool_part, c.L.cnt_part, &nch, c.st));
  VKN_TRY(launch_pool_reduce(fs, c.L.pool_part, c.L.cnt_part, nch, c.L.xp0, c.L.cnt, c.st));
  // (4) proposal_feats = init_w (broadcast over frames) + obj
  RowSrc r = src_copy(c.L.xp0, C);
  r.pres = init_w;
  r.ldpres = C;
  r.pres_mod = N;
  return launch_rowop(r, proposal_feats, C, c.P, C, c.st);
}

// Link block on the tcgen05 row engine (many rows in flight: the sharded clip links all frames of a rank at once):
// same operators as below, activations handed between the row GEMMs as bf16 planes.
static int link_planes(Ctx &c, const VknLinkW &w, const float *cur, const RowSrc &kv, float *out) {
  const int C = c.s.C, P = c.P, F = c.s.ffn_dim;
  const long long PS = (long long)P * C;
  void *PLA = c.L.pl[0], *PLB = c.L.pl[1], *PLC = c.L.pl[2];
  VKN_TRY(launch_rowprep(src_copy(cur, C), nullptr, 0, PLA, C, PS, P, C, c.st));
  VKN_TRY(launch_rowprep(kv, nullptr, 0, PLB, C, PS, P, C, c.st));
  if (c.chain_on) c.chain = chain_begin(P);
  // MHA(q = cur, k = v = kv): q rows of in_proj on cur, k/v rows on kv (video/kernel_update_head.py:404-406)
  LinArgs two[2];
  two[0] = lin(src_planes(PLA, C, PS), w.attn.in_w, C, w.attn.in_b, c.L.qkv, 3 * C, P, C, C, 0);
  two[1] = lin(src_planes(PLB, C, PS), wrow(c, w.attn.in_w, C, C), C, w.attn.in_b + C, c.L.qkv + C, 3 * C, P, 2 * C, C, 0);
  VKN_TRY(emit_gemm(c, two, 2));
  VKN_TRY(emit_flush(c));
  VKN_TRY(launch_attention(c.L.qkv, 3 * C, c.L.qkv + C, 3 * C, c.L.qkv + 2 * C, 3 * C, nullptr, C, c.s.B, c.s.N, C,
                           c.s.num_heads, c.st, PLC, PS));
  if (c.chain) {
    // out-projection + residual + LayerNorm, FFN, + residual + LayerNorm: one chain launch, LayerNorms in the epilogues
    LinArgs op = lin(src_planes(PLC, C, PS), w.attn.out_w, C, w.attn.out_b, c.L.o2, C, P, C, C, EPI_RES | EPI_LN);
    op.res = cur;
    op.ldres = C;
    op.ln_g = w.attn.norm_g;
    op.ln_b = w.attn.norm_b;
    out_planes(op, PLA, P, C);
    VKN_TRY(emit_gemm(c, &op, 1));
    LinArgs f1 = lin(src_planes(PLA, C, PS), w.ffn.w1, C, w.ffn.b1, nullptr, F, P, F, C, EPI_RELU | EPI_NOOUT);
    out_planes(f1, c.L.h, P, F);
    VKN_TRY(emit_gemm(c, &f1, 1));
    LinArgs f2 = lin(src_planes(c.L.h, F, (long long)P * F), w.ffn.w2, F, w.ffn.b2, out, C, P, C, F, EPI_RES | EPI_LN);
    f2.res = c.L.o2;
    f2.ldres = C;
    f2.ln_g = w.ffn.norm_g;
    f2.ln_b = w.ffn.norm_b;
    VKN_TRY(emit_gemm(c, &f2, 1));
    VKN_TRY(emit_flush(c));
    c.chain = nullptr;
    return VKN_OK;
  }
  LinArgs op = lin(src_planes(PLC, C, PS), w.attn.out_w, C, w.attn.out_b, c.L.y, C, P, C, C, EPI_RES);
  op.res = cur;
  op.ldres = C;
  VKN_TRY(launch_linear_tc(&op, 1, c.st));
  VKN_TRY(launch_rowprep(src_ln(c.L.y, C, w.attn.norm_g, w.attn.norm_b, false), c.L.o2, C, PLA, C, PS, P, C, c.st));
  // link FFN + LN (:407-415)
  LinArgs f1 = lin(src_planes(PLA, C, PS), w.ffn.w1, C, w.ffn.b1, nullptr, F, P, F, C, EPI_RELU | EPI_NOOUT);
  out_planes(f1, c.L.h, P, F);
  VKN_TRY(launch_linear_tc(&f1, 1, c.st));
  LinArgs f2 = lin(src_planes(c.L.h, F, (long long)P * F), w.ffn.w2, F, nullptr, c.L.zpart, C, P, C, F, 0);
  const int nk = ceil_div(F, 64);
  int ksp = 148 / (ceil_div(P, 128) * ceil_div(C, 256));
  ksp = ksp >= 8 ? 8 : (ksp >= 4 ? 4 : (ksp >= 2 ? 2 : 1));
  while (ksp > 1 && (nk % ksp != 0 || nk / ksp < 4)) ksp /= 2;
  f2.ksplit = ksp;
  f2.out_split_stride = PS;
  VKN_TRY(launch_linear_tc(&f2, 1, c.st));
  RowSrc r = src_ln(c.L.zpart, C, w.ffn.norm_g, w.ffn.norm_b, false);
  r.nsum = ksp;
  r.sum_stride = PS;
  r.pbias = w.ffn.b2;
  r.pres = c.L.o2;
  r.ldpres = C;
  return launch_rowprep(r, out, C, nullptr, 0, 0, P, C, c.st);
}

int vkn_link_attend(const VknShape *s, const VknLinkW *w, const float *cur, const float *prev, const float *x_feat,
                    float *out, void *workspace, size_t workspace_bytes, void *stream) {
  Ctx c;
  VKN_TRY(make_ctx(s, workspace, workspace_bytes, stream, c));
  if (!w || !cur || !prev || !out) VKN_FAIL(VKN_E_INVALID, "vkn_link_attend: null argument");
  const int C = s->C;
  RowSrc kv = src_copy(prev, C);
  if (w->has_updator) {
    if (!x_feat) VKN_FAIL(VKN_E_INVALID, "vkn_link_attend: the updator link needs x_feat");
    VKN_TRY(k_update(c, w->upd, x_feat, src_copy(prev, C), c.L.link_fc));
    kv = src_ln(c.L.link_fc, C, w->upd.fc_norm_g, w->upd.fc_norm_b, true);
  }
  if (c.rows_tc && (reinterpret_cast<uintptr_t>(cur) & 15) == 0) return link_planes(c, *w, cur, kv, out);
  VKN_TRY(k_attn(c, w->attn, src_copy(cur, C), cur, &kv, c.L.y));
  RowSrc pend;
  VKN_TRY(k_ffn(c, w->ffn, src_ln(c.L.y, C, w->attn.norm_g, w->attn.norm_b, false), &pend));
  return launch_rowop(pend, out, C, c.P, C, c.st);
}

int vkn_rescale_masks(const void *masks, int dtype, int K, int H, int W, int up, int batch_h, int batch_w, int img_h,
                      int img_w, int ori_h, int ori_w, float mask_thr, float *probs, unsigned char *bits, void *stream) {
  if (!masks) VKN_FAIL(VKN_E_INVALID, "vkn_rescale_masks: null masks");
  return launch_rescale_masks(masks, dtype, K, H, W, up, batch_h, batch_w, img_h, img_w, ori_h, ori_w, mask_thr, probs, bits,
                              (cudaStream_t)stream);
}

int vkn_panoptic_merge(const float *masks, const float *scores, const int32_t *labels, int num_kernels, int H, int W,
                       int num_thing_classes, double instance_score_thr, double overlap_thr, int32_t *panoptic_seg,
                       int32_t *segments, float *segment_scores, int32_t *kept_things, int32_t *counts, void *workspace,
                       size_t workspace_bytes, void *stream) {
  if (!masks || !scores || !labels || !panoptic_seg || !segments || !segment_scores || !kept_things || !counts)
    VKN_FAIL(VKN_E_INVALID, "vkn_panoptic_merge: null argument");
  return launch_panoptic_merge(masks, scores, labels, num_kernels, H, W, num_thing_classes, instance_score_thr, overlap_thr,
                               panoptic_seg, segments, segment_scores, kept_things, counts, workspace, workspace_bytes,
                               (cudaStream_t)stream);
}

int vkn_mlp(const VknMlpLayer *layers, int num_layers, int w_dtype, const float *in, float *out, int rows, void *workspace,
            size_t workspace_bytes, void *stream) {
  if (!layers || num_layers < 1 || num_layers > 16 || !in || !out || rows < 1) VKN_FAIL(VKN_E_INVALID, "vkn_mlp: bad argument");
  if (w_dtype != VKN_F32 && w_dtype != VKN_BF16) VKN_FAIL(VKN_E_INVALID, "vkn_mlp: bad weight dtype");
  int maxd = 0;
  for (int i = 0; i < num_layers; ++i) {
    const VknMlpLayer &l = layers[i];
    if (!l.w || l.in_dim < 1 || l.out_dim < 1 || (i > 0 && l.in_dim != layers[i - 1].out_dim))
      VKN_FAIL(VKN_E_INVALID, "vkn_mlp: layer %d is inconsistent", i);
    if ((l.ln_g == nullptr) != (l.ln_b == nullptr)) VKN_FAIL(VKN_E_INVALID, "vkn_mlp: layer %d has half a LayerNorm", i);
    if (l.ln_g && l.out_dim > 256) VKN_FAIL(VKN_E_UNSUPPORTED, "vkn_mlp: LayerNorm over %d > 256 features", l.out_dim);
    if (l.in_dim % 2 != 0) VKN_FAIL(VKN_E_UNSUPPORTED, "vkn_mlp: odd feature count %d", l.in_dim);
    maxd = l.out_dim > maxd ? l.out_dim : maxd;
  }
  const size_t buf = align_up((size_t)rows * maxd * sizeof(float), 256);
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255) || workspace_bytes < 2 * buf)
    VKN_FAIL(VKN_E_WORKSPACE, "vkn_mlp: workspace must be 256-byte aligned and hold %zu bytes", 2 * buf);
  float *pp[2] = {(float *)workspace, (float *)((char *)workspace + buf)};
  cudaStream_t st = (cudaStream_t)stream;
  // layer i: y = W x (+ b); its LayerNorm (+ ReLU) is the PROLOGUE of layer i + 1 (one fused launch per Linear); a ReLU
  // without LayerNorm is the epilogue of the layer itself
  RowSrc src = src_copy(in, layers[0].in_dim);
  for (int i = 0; i < num_layers; ++i) {
    const VknMlpLayer &l = layers[i];
    const bool last = i == num_layers - 1;
    float *dst = (last && !l.ln_g) ? out : pp[i & 1];
    LinArgs a = lin(src, l.w, l.in_dim, l.b, dst, l.out_dim, rows, l.out_dim, l.in_dim, (l.relu && !l.ln_g) ? EPI_RELU : 0);
    VKN_TRY(launch_linear(&a, 1, w_dtype, st));
    src = l.ln_g ? src_ln(dst, l.out_dim, l.ln_g, l.ln_b, l.relu != 0) : src_copy(dst, l.out_dim);
    if (last && l.ln_g) VKN_TRY(launch_rowop(src, out, l.out_dim, rows, l.out_dim, st));
  }
  return VKN_OK;
}

int vkn_track_match(const float *bboxes, const int64_t *labels, const float *embeds, int n, int embed_dim,
                    const int64_t *memo_labels, const float *memo_embeds, const int64_t *memo_ids, int m,
                    const float *thresholds, int with_cats, int64_t num_tracklets, int32_t *selected, int64_t *ids,
                    int32_t *counts, void *workspace, size_t workspace_bytes, void *stream) {
  if (!thresholds || !selected || !ids || !counts || (n > 0 && (!bboxes || !labels || !embeds)) ||
      (m > 0 && (!memo_labels || !memo_embeds || !memo_ids)))
    VKN_FAIL(VKN_E_INVALID, "vkn_track_match: null argument");
  return launch_track_match(bboxes, (const long long *)labels, embeds, n, embed_dim, (const long long *)memo_labels, memo_embeds,
                            (const long long *)memo_ids, m, thresholds, with_cats, (long long)num_tracklets, selected,
                            (long long *)ids, counts, workspace, workspace_bytes, (cudaStream_t)stream);
}

int vkn_mask_boxes(const void *masks, int elem_bytes, int K, int H, int W, float *boxes, void *stream) {
  if (!masks && K > 0) VKN_FAIL(VKN_E_INVALID, "vkn_mask_boxes: null masks");
  return launch_mask_boxes(masks, elem_bytes, K, H, W, boxes, (cudaStream_t)stream);
}

}  // extern "C"
