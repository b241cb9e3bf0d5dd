// Host-side orchestration of the grouped ViT-Tiny backbone forward / backward and of the
// heads + loss step, plus the extern "C" boundary (include/vit2spn.h).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "mlp_tc.cuh"

namespace v2s {

static thread_local char g_err[1024] = "";
int64_t g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- optional per-kernel-class device timing (bench.py roofline; off by default) -------------
namespace prof {
constexpr int MAX_REC = 8192;
struct Rec { int cls; double work; double bytes; cudaEvent_t a, b; };
static bool enabled = false;
static Rec recs[MAX_REC];
static int n_recs = 0;
static int n_events = 0;   // events created so far (recs[i].a/b valid for i < n_events)
static const char* names[] = {"gemm_patch", "gemm_qkv", "gemm_proj", "gemm_fc1", "gemm_fc2", "gemm_dgrad",
                              "gemm_wgrad", "attn_fwd", "attn_bwd", "ln_fwd", "ln_bwd", "colsum", "misc", "heads", "gemm_mlp_fwd",
                              "gemm_mlp_bwd"};
enum { C_PATCH, C_QKV, C_PROJ, C_FC1, C_FC2, C_DGRAD, C_WGRAD, C_ATTN_F, C_ATTN_B, C_LN_F, C_LN_B, C_COLSUM, C_MISC,
       C_HEADS, C_MLP_F, C_MLP_B, C_COUNT };
struct Scope {
  int idx = -1;
  cudaStream_t st;
  Scope(int cls, double work, cudaStream_t s, double bytes = 0.0) : st(s) {
    if (!enabled || n_recs >= MAX_REC) return;
    idx = n_recs++;
    if (idx >= n_events) { cudaEventCreate(&recs[idx].a); cudaEventCreate(&recs[idx].b); n_events = idx + 1; }
    recs[idx].cls = cls; recs[idx].work = work; recs[idx].bytes = bytes;
    cudaEventRecord(recs[idx].a, st);
  }
  ~Scope() { if (idx >= 0) cudaEventRecord(recs[idx].b, st); }
};
}  // namespace prof

namespace {

inline int64_t align_up(int64_t v, int64_t a = 1024) { return (v + a - 1) / a * a; }

// ---- workspace carve-up ---------------------------------------------------------------------
struct LayerStash {   // byte offsets relative to the slot base
  int64_t mean1, rstd1, xn1, qkv, ctx, lse, x_mid, mean2, rstd2, xn2, u, h;
};
struct Plan {
  int B, at;
  int64_t es;          // activation element size
  int64_t M, MP;
  int64_t heads_bytes;
  // per-group forward scratch (groups that do not save activations)
  int64_t f_patches, f_xa, f_xb, f_xn, f_qkv, f_ctx, f_h, f_bytes;
  // per-slot stash
  int64_t s_patches, s_x[NL + 1];
  LayerStash s_layer[NL];
  int64_t s_bytes;
  // per-slot backward scratch
  int64_t b_dx, b_dxlp, b_big, b_tmp, b_bytes;
};

Plan make_plan(int B, int mode) {
  Plan p;
  memset(&p, 0, sizeof(p));
  p.B = B;
  p.at = mode == V2S_MODE_BF16 ? AT_BF16 : mode == V2S_MODE_FP16 ? AT_F16 : AT_F32;   // activation type tag
  p.es = p.at ? 2 : 4;
  p.M = (int64_t)B * NT;
  p.MP = (int64_t)B * NP;
  p.heads_bytes = align_up((int64_t)B * 8192 * 4 + 4096);
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t r = o; o += align_up(bytes); return r; };
  p.f_patches = take(p.MP * KPE * p.es);
  p.f_xa = take(p.M * D * 4);
  p.f_xb = take(p.M * D * 4);
  p.f_xn = take(p.M * D * p.es);
  p.f_qkv = take(p.M * 3 * D * p.es);
  p.f_ctx = take(p.M * D * p.es);
  p.f_h = take(p.M * DF * p.es);
  p.f_bytes = o;
  o = 0;
  p.s_patches = take(p.MP * KPE * p.es);
  for (int l = 0; l <= NL; ++l) p.s_x[l] = take(p.M * D * 4);
  for (int l = 0; l < NL; ++l) {
    LayerStash& s = p.s_layer[l];
    s.mean1 = take(p.M * 4); s.rstd1 = take(p.M * 4);
    s.xn1 = take(p.M * D * p.es);
    s.qkv = take(p.M * 3 * D * p.es);
    s.ctx = take(p.M * D * p.es);
    s.lse = take((int64_t)B * NH * NT * 4);
    s.x_mid = take(p.M * D * 4);
    s.mean2 = take(p.M * 4); s.rstd2 = take(p.M * 4);
    s.xn2 = take(p.M * D * p.es);
    s.u = take(p.M * DF * p.es);
    s.h = take(p.M * DF * p.es);
  }
  p.s_bytes = o;
  o = 0;
  p.b_dx = take(p.M * D * 4);
  p.b_dxlp = take(p.M * D * p.es);
  p.b_big = take(p.M * DF * p.es);
  p.b_tmp = take(p.M * D * p.es);
  p.b_bytes = o;
  return p;
}

int64_t plan_total(const Plan& p, int n_groups, int n_saved) {
  return p.heads_bytes + (int64_t)n_groups * p.f_bytes + (int64_t)n_saved * (p.s_bytes + p.b_bytes);
}

// base addresses inside the workspace; regions are laid out for the maximum MAXG groups so that
// forward and backward calls agree regardless of how many groups each call carries
struct Regions {
  char* heads;
  char* fwd[MAXG];
  char* stash[MAXG];
  char* bwd[MAXG];
};

int resolve_regions(const Plan& p, void* ws, int64_t ws_bytes, int n_groups, int n_saved, Regions* r) {
  if (n_groups < 0 || n_groups > MAXG || n_saved < 0 || n_saved > MAXG) {
    set_error("workspace: bad group counts");
    return 1;
  }
  // layout: heads | stash[0..n_saved) | bwd[0..n_saved) | fwd[0..n_groups)
  // (stash first so that its addresses do not depend on n_groups)
  const int64_t need = plan_total(p, n_groups, n_saved);
  if (!ws || ws_bytes < need) {
    set_error("workspace too small: have %lld bytes, need %lld", (long long)ws_bytes, (long long)need);
    return 1;
  }
  char* base = static_cast<char*>(ws);
  if ((reinterpret_cast<uintptr_t>(base) & 1023) != 0) {
    set_error("workspace must be 1024-byte aligned");
    return 1;
  }
  r->heads = base;
  char* q = base + p.heads_bytes;
  for (int i = 0; i < MAXG; ++i) { r->stash[i] = i < n_saved ? q : nullptr; if (i < n_saved) q += p.s_bytes; }
  for (int i = 0; i < MAXG; ++i) { r->bwd[i] = i < n_saved ? q : nullptr; if (i < n_saved) q += p.b_bytes; }
  for (int i = 0; i < MAXG; ++i) { r->fwd[i] = i < n_groups ? q : nullptr; if (i < n_groups) q += p.f_bytes; }
  return 0;
}

// number of stash slots the workspace was sized for is implied by the highest slot in use:
// the Python side always allocates with n_saved = 2 (two online streams) or 1; slots index the
// stash region directly, so forward and backward agree as long as the same workspace is passed.
int max_slot(const v2s_group_t* g, int n) {
  int m = -1;
  for (int i = 0; i < n; ++i) if (g[i].slot > m) m = g[i].slot;
  return m;
}

int run_gemm(const GemmDesc& d, int ta, int tb, int to, cudaStream_t s, int cls = prof::C_MISC) {
  // algorithmic HBM bytes: both operands and the output once (+ residual / auxiliary operand, second output)
  const double ea = ta ? 2.0 : 4.0, eb = tb ? 2.0 : 4.0, eo = to ? 2.0 : 4.0;
  double bytes = 0.0;
  for (int g = 0; g < d.groups; ++g) {
    bytes += (double)d.M * d.K * ea + (double)d.N * d.K * eb + (double)d.M * d.N * eo;
    if (d.epi == EPI_BIAS_RESID) bytes += (double)d.M * d.N * 4.0 + (d.ln_out[g] ? (double)d.M * d.N * 2.0 : 0.0);
    // second 16-bit tensor: gelu'(u) operand of the dgrad epilogue; the pre-GELU store only for groups that keep it
    if (d.epi == EPI_DGELU || (d.epi == EPI_BIAS_GELU && d.out[g])) bytes += (double)d.M * d.N * 2.0;
  }
  double flops = 2.0 * d.M * d.N * (double)d.K * d.groups;
  if (d.M2 > 0) {      // two-problem wgrad launch: each half of the groups has its own shape
    bytes = 0.5 * d.groups * (((double)d.M + d.N + d.M2 + d.N2) * d.K * ea + ((double)d.M * d.N + (double)d.M2 * d.N2) * eo);
    flops = d.groups * ((double)d.M * d.N + (double)d.M2 * d.N2) * (double)d.K;
  }
  prof::Scope scope(cls, flops, s, bytes);
  int handled = 0;
  V2S_TRY(launch_gemm_tc(d, ta, tb, to, s, &handled));
  if (handled) return 0;
  return launch_gemm_simt(d, ta, tb, to, s);
}

inline const void* weight_ptr(const v2s_group_t& g, int at, int64_t off) {
  return at ? static_cast<const void*>(static_cast<const bf16*>(g.params_lp) + off)
            : static_cast<const void*>(g.params + off);
}

int launch_attention_fwd(const void* const* qkv, void* const* ctx, float* const* lse, int groups, int B, int at,
                         cudaStream_t s) {
  prof::Scope scope(prof::C_ATTN_F, 4.0 * B * NH * (double)NT * NT * DH * groups, s, (double)B * NT * 4 * D * 2.0 * groups);
  if (at != AT_F32 && tc_enabled()) return launch_attn_fwd_tc(qkv, ctx, lse, groups, B, s, at == AT_F16);
  return launch_attn_fwd_simt(qkv, ctx, lse, groups, B, at, s);
}
int launch_attention_bwd(const void* const* qkv, const void* const* ctx, const float* const* lse,
                         const void* const* dctx, void* const* dqkv, int groups, int B, int at, cudaStream_t s) {
  prof::Scope scope(prof::C_ATTN_B, 10.0 * B * NH * (double)NT * NT * DH * groups, s, (double)B * NT * 8 * D * 2.0 * groups);
  if (at != AT_F32 && tc_enabled()) return launch_attn_bwd_tc(qkv, ctx, lse, dctx, dqkv, groups, B, s, at == AT_F16);
  return launch_attn_bwd_simt(qkv, ctx, lse, dctx, dqkv, groups, B, at, s);
}

int wgrad_split(int64_t rows) {
  int64_t s = rows / 1024;
  if (s < 1) s = 1;
  if (s > 16) s = 16;
  return (int)s;
}

}  // namespace

// =============================================================================================
// backbone forward
// =============================================================================================
static int backbone_forward_impl(const v2s_group_t* gs, int G, int B, int mode, void* ws, int64_t ws_bytes,
                                 cudaStream_t st) {
  if (mode < V2S_MODE_FP32 || mode > V2S_MODE_FP16) { set_error("bad compute mode %d", mode); return 1; }
  if (G < 1 || G > MAXG) { set_error("backbone_forward: 1..4 groups"); return 1; }
  if (B < 1) { set_error("backbone_forward: batch must be >= 1"); return 1; }
  const Plan p = make_plan(B, mode);
  const int at = p.at;
  const int n_saved = max_slot(gs, G) + 1;
  Regions R;
  V2S_TRY(resolve_regions(p, ws, ws_bytes, G, n_saved, &R));
  for (int g = 0; g < G; ++g) {
    if (!gs[g].params || !gs[g].x) { set_error("backbone_forward: group %d has null params/x", g); return 1; }
    if (at && !gs[g].params_lp) { set_error("backbone_forward: the 16-bit modes need params_lp (group %d)", g); return 1; }
    if (gs[g].slot >= MAXG) { set_error("backbone_forward: bad slot"); return 1; }
    if (gs[g].x_format != 0 && (gs[g].x_format != 1 || !at)) {
      set_error("backbone_forward: group %d: x_format %d (1 = 16-bit patch rows, 16-bit modes only)", g, gs[g].x_format);
      return 1;
    }
  }
  const int64_t M = p.M, MP = p.MP;

  // ---- buffer resolution ----
  auto saved = [&](int g) { return gs[g].slot >= 0; };
  auto sb = [&](int g, int64_t off) -> char* { return R.stash[gs[g].slot] + off; };
  auto fb = [&](int g, int64_t off) -> char* { return R.fwd[g] + off; };

  // ---- patch embedding: im2col (shared between groups that read the same images) ----
  void* patches[MAXG];
  {
    const float* ux[MAXG]; void* uo[MAXG]; int nu = 0;
    for (int g = 0; g < G; ++g) {
      if (gs[g].x_format == 1) { patches[g] = const_cast<void*>(gs[g].x); continue; }      // the caller's patch matrix
      patches[g] = saved(g) ? sb(g, p.s_patches) : fb(g, p.f_patches);
      int dup = -1;
      // a saved group must own its copy (it outlives the call); unsaved groups may alias
      if (!saved(g))
        for (int k = 0; k < g; ++k) if (gs[k].x == gs[g].x) { dup = k; break; }
      if (dup >= 0) { patches[g] = patches[dup]; continue; }
      ux[nu] = static_cast<const float*>(gs[g].x); uo[nu] = patches[g]; ++nu;
    }
    if (nu > 0) V2S_TRY(launch_im2col(ux, uo, nu, B, at, st));
  }
  float* x_cur[MAXG];
  {
    GemmDesc d = make_gemm_desc();
    d.M = (int)MP; d.N = D; d.K = KPE; d.groups = G;
    d.a_rs = KPE; d.a_cs = 1; d.b_rs = 1; d.b_cs = KPE;
    d.ldc = D;
    const float* pp[MAXG]; const float* tok[MAXG];
    const bool two_pass = at != AT_F32 && tc_enabled();   // tensor-core GEMM into patch rows, then assemble tokens
    d.epi = two_pass ? EPI_STORE : EPI_PATCH;
    for (int g = 0; g < G; ++g) {
      x_cur[g] = reinterpret_cast<float*>(saved(g) ? sb(g, p.s_x[0]) : fb(g, p.f_xa));
      // scratch for the [B*196,192] patch rows: the x_mid buffer of block 0 (written later)
      float* tmp = reinterpret_cast<float*>(saved(g) ? sb(g, p.s_layer[0].x_mid) : fb(g, p.f_xb));
      d.A[g] = patches[g];
      d.B[g] = weight_ptr(gs[g], at, OFF_WPE);
      d.bias[g] = gs[g].params + OFF_BPE;
      d.aux[g] = gs[g].params + OFF_POS;
      d.out[g] = two_pass ? tmp : x_cur[g];
      pp[g] = gs[g].params; tok[g] = tmp;
    }
    V2S_TRY(run_gemm(d, at, at, 0, st, prof::C_PATCH));
    if (two_pass) V2S_TRY(launch_assemble_tokens(pp, tok, x_cur, G, B, st));
    else V2S_TRY(launch_cls_rows(pp, x_cur, G, B, st));
  }

  // ---- 12 pre-LN blocks ----
  // On the tensor-core path every LayerNorm except the first is fused into the epilogue of the GEMM that
  // produces its input row (attention projection → LN2, fc2 → LN1 of the next block): V2S_NO_LNFUSE=1 disables.
  const bool ln_fuse = at != AT_F32 && tc_enabled() && !getenv("V2S_NO_LNFUSE");
  // fc1 -> GELU -> fc2 chained on chip (mlp_tc.cu): V2S_NO_MLPFUSE=1 falls back to the two separate GEMMs
  const bool mlp_fuse = ln_fuse && !getenv("V2S_NO_MLPFUSE");
  const int lp_f16 = at == AT_F16 ? 1 : 0;
  bool ln1_done = false;
  for (int l = 0; l < NL; ++l) {
    const int64_t lo = layer_off(l);
    const float *xin[MAXG], *gam[MAXG], *bet[MAXG];
    void *xn[MAXG], *qkv[MAXG], *ctx[MAXG], *u[MAXG], *h[MAXG];
    float *mean[MAXG], *rstd[MAXG], *lse[MAXG], *xmid[MAXG], *xout[MAXG];
    for (int g = 0; g < G; ++g) {
      xin[g] = x_cur[g];
      if (saved(g)) {
        const LayerStash& s = p.s_layer[l];
        xn[g] = sb(g, s.xn1); qkv[g] = sb(g, s.qkv); ctx[g] = sb(g, s.ctx);
        mean[g] = (float*)sb(g, s.mean1); rstd[g] = (float*)sb(g, s.rstd1); lse[g] = (float*)sb(g, s.lse);
        xmid[g] = (float*)sb(g, s.x_mid); u[g] = sb(g, s.u); h[g] = sb(g, s.h);
        xout[g] = (float*)sb(g, p.s_x[l + 1]);
      } else {
        xn[g] = fb(g, p.f_xn); qkv[g] = fb(g, p.f_qkv); ctx[g] = fb(g, p.f_ctx);
        mean[g] = nullptr; rstd[g] = nullptr; lse[g] = nullptr;
        xmid[g] = (float*)fb(g, p.f_xb); u[g] = nullptr; h[g] = fb(g, p.f_h);
        xout[g] = (float*)fb(g, p.f_xa);   // in place: x_in is dead once x_mid exists
      }
      gam[g] = gs[g].params + lo + L_LN1W; bet[g] = gs[g].params + lo + L_LN1B;
    }
    if (!ln1_done) {
      prof::Scope sc(prof::C_LN_F, (double)M * D * (4 + p.es) * G, st);
      V2S_TRY(launch_ln_fwd(xin, gam, bet, xn, mean, rstd, G, (int)M, at, st));
    }
    {  // fused QKV projection
      GemmDesc d = make_gemm_desc();
      d.M = (int)M; d.N = 3 * D; d.K = D; d.groups = G;
      d.a_rs = D; d.a_cs = 1; d.b_rs = 1; d.b_cs = D; d.epi = EPI_STORE; d.ldc = 3 * D;
      for (int g = 0; g < G; ++g) {
        d.A[g] = xn[g]; d.B[g] = weight_ptr(gs[g], at, lo + L_WQKV);
        d.bias[g] = gs[g].params + lo + L_BQKV; d.out[g] = qkv[g];
      }
      V2S_TRY(run_gemm(d, at, at, at, st, prof::C_QKV));
    }
    {
      const void* cq[MAXG];
      for (int g = 0; g < G; ++g) cq[g] = qkv[g];
      V2S_TRY(launch_attention_fwd(cq, ctx, lse, G, B, at, st));
    }
    {  // attention output projection + residual
      GemmDesc d = make_gemm_desc();
      d.M = (int)M; d.N = D; d.K = D; d.groups = G;
      d.a_rs = D; d.a_cs = 1; d.b_rs = 1; d.b_cs = D; d.epi = EPI_BIAS_RESID; d.ldc = D;
      for (int g = 0; g < G; ++g) {
        d.A[g] = ctx[g]; d.B[g] = weight_ptr(gs[g], at, lo + L_WO);
        d.bias[g] = gs[g].params + lo + L_BO; d.resid[g] = xin[g]; d.out[g] = xmid[g];
        if (ln_fuse) {     // LN2 of this block, fused
          d.ln_out[g] = saved(g) ? (void*)sb(g, p.s_layer[l].xn2) : (void*)fb(g, p.f_xn);
          d.ln_gamma[g] = gs[g].params + lo + L_LN2W; d.ln_beta[g] = gs[g].params + lo + L_LN2B;
          d.ln_mean[g] = saved(g) ? (float*)sb(g, p.s_layer[l].mean2) : nullptr;
          d.ln_rstd[g] = saved(g) ? (float*)sb(g, p.s_layer[l].rstd2) : nullptr;
        }
      }
      V2S_TRY(run_gemm(d, at, at, 0, st, prof::C_PROJ));
    }
    {
      const float* xm[MAXG];
      void* xn2[MAXG]; float *m2[MAXG], *r2[MAXG];
      for (int g = 0; g < G; ++g) {
        xm[g] = xmid[g];
        gam[g] = gs[g].params + lo + L_LN2W; bet[g] = gs[g].params + lo + L_LN2B;
        if (saved(g)) { xn2[g] = sb(g, p.s_layer[l].xn2); m2[g] = (float*)sb(g, p.s_layer[l].mean2); r2[g] = (float*)sb(g, p.s_layer[l].rstd2); }
        else { xn2[g] = fb(g, p.f_xn); m2[g] = nullptr; r2[g] = nullptr; }
      }
      if (!ln_fuse) {
        prof::Scope sc(prof::C_LN_F, (double)M * D * (4 + p.es) * G, st);
        V2S_TRY(launch_ln_fwd(xm, gam, bet, xn2, m2, r2, G, (int)M, at, st));
      }
      if (mlp_fuse) {
        // fc1 -> GELU -> fc2 + residual (+ LN1 of the next block) as ONE chained kernel: h never leaves the chip for
        // the EMA-target groups; the online groups still write u and h for the backward pass
        MlpDesc d = make_mlp_desc();
        d.mode = MLP_FWD; d.M = (int)M; d.groups = G; d.lp_f16 = lp_f16;
        double bytes = 0.0;
        for (int g = 0; g < G; ++g) {
          d.a[g] = xn2[g];
          d.w1[g] = weight_ptr(gs[g], at, lo + L_W1); d.w2[g] = weight_ptr(gs[g], at, lo + L_W2);
          d.b1[g] = gs[g].params + lo + L_B1; d.b2[g] = gs[g].params + lo + L_B2;
          d.u[g] = u[g]; d.h[g] = saved(g) ? h[g] : nullptr;
          d.resid[g] = xmid[g]; d.out[g] = xout[g];
          bytes += (double)M * D * (2 + 4 + 4) + 2.0 * D * DF * 2 + (saved(g) ? 2.0 * M * DF * 2 : 0.0);
          if (l + 1 < NL) {     // LN1 of the next block, fused
            const int64_t ln = layer_off(l + 1);
            d.ln_out[g] = saved(g) ? (void*)sb(g, p.s_layer[l + 1].xn1) : (void*)fb(g, p.f_xn);
            d.ln_gamma[g] = gs[g].params + ln + L_LN1W; d.ln_beta[g] = gs[g].params + ln + L_LN1B;
            d.ln_mean[g] = saved(g) ? (float*)sb(g, p.s_layer[l + 1].mean1) : nullptr;
            d.ln_rstd[g] = saved(g) ? (float*)sb(g, p.s_layer[l + 1].rstd1) : nullptr;
            bytes += (double)M * D * 2;
          }
        }
        prof::Scope scope(prof::C_MLP_F, 4.0 * M * D * (double)DF * G, st, bytes);
        V2S_TRY(launch_mlp_tc(d, st));
        ln1_done = l + 1 < NL;
        for (int g = 0; g < G; ++g) x_cur[g] = xout[g];
        continue;
      }
      GemmDesc d = make_gemm_desc();   // fc1 + GELU
      d.M = (int)M; d.N = DF; d.K = D; d.groups = G;
      d.a_rs = D; d.a_cs = 1; d.b_rs = 1; d.b_cs = D; d.epi = EPI_BIAS_GELU; d.ldc = DF;
      for (int g = 0; g < G; ++g) {
        d.A[g] = xn2[g]; d.B[g] = weight_ptr(gs[g], at, lo + L_W1);
        d.bias[g] = gs[g].params + lo + L_B1; d.out[g] = u[g]; d.out2[g] = h[g];
      }
      V2S_TRY(run_gemm(d, at, at, at, st, prof::C_FC1));
    }
    {  // fc2 + residual
      GemmDesc d = make_gemm_desc();
      d.M = (int)M; d.N = D; d.K = DF; d.groups = G;
      d.a_rs = DF; d.a_cs = 1; d.b_rs = 1; d.b_cs = DF; d.epi = EPI_BIAS_RESID; d.ldc = D;
      for (int g = 0; g < G; ++g) {
        d.A[g] = h[g]; d.B[g] = weight_ptr(gs[g], at, lo + L_W2);
        d.bias[g] = gs[g].params + lo + L_B2; d.resid[g] = xmid[g]; d.out[g] = xout[g];
        if (ln_fuse && l + 1 < NL) {     // LN1 of the next block, fused
          const int64_t ln = layer_off(l + 1);
          d.ln_out[g] = saved(g) ? (void*)sb(g, p.s_layer[l + 1].xn1) : (void*)fb(g, p.f_xn);
          d.ln_gamma[g] = gs[g].params + ln + L_LN1W; d.ln_beta[g] = gs[g].params + ln + L_LN1B;
          d.ln_mean[g] = saved(g) ? (float*)sb(g, p.s_layer[l + 1].mean1) : nullptr;
          d.ln_rstd[g] = saved(g) ? (float*)sb(g, p.s_layer[l + 1].rstd1) : nullptr;
        }
      }
      V2S_TRY(run_gemm(d, at, at, 0, st, prof::C_FC2));
      ln1_done = ln_fuse && l + 1 < NL;
    }
    for (int g = 0; g < G; ++g) x_cur[g] = xout[g];
  }

  // ---- outputs: hidden_states[-1] and its mean over the 197 tokens ----
  {
    const float* hid[MAXG]; float* feat[MAXG]; int64_t fs[MAXG]; int nf = 0;
    for (int g = 0; g < G; ++g) {
      if (gs[g].hidden)
        V2S_CUDA_OK(cudaMemcpyAsync(gs[g].hidden, x_cur[g], (size_t)M * D * 4, cudaMemcpyDeviceToDevice, st));
      if (gs[g].feat) { hid[nf] = x_cur[g]; feat[nf] = gs[g].feat; fs[nf] = gs[g].feat_stride > 0 ? gs[g].feat_stride : D; ++nf; }
    }
    if (nf) V2S_TRY(launch_pool_fwd(hid, feat, fs, nf, B, st));
  }
  return 0;
}

// =============================================================================================
// backbone backward
// =============================================================================================
// layers [layer_lo, layer_hi) in descending order; layer_hi == NL starts from the feature / hidden-state gradient,
// layer_lo == 0 finishes with the embedding gradients.  Between two calls the residual-stream gradient lives in the
// workspace, so a full backward may be issued in several ranges (gradient all-reduce overlapped per range).
static int backbone_backward_impl(const v2s_group_t* gs_in, int G_in, int B, int mode, void* ws, int64_t ws_bytes,
                                  cudaStream_t st, int layer_hi = NL, int layer_lo = 0) {
  if (layer_lo < 0 || layer_hi > NL || layer_lo >= layer_hi) { set_error("backbone_backward: bad layer range [%d, %d)", layer_lo, layer_hi); return 1; }
  // keep only the groups that saved activations and want gradients
  v2s_group_t gs[MAXG];
  int G = 0;
  for (int i = 0; i < G_in && i < MAXG; ++i)
    if (gs_in[i].slot >= 0 && gs_in[i].grads) gs[G++] = gs_in[i];
  if (G == 0) { set_error("backbone_backward: no group with slot >= 0 and grads"); return 1; }
  const Plan p = make_plan(B, mode);
  const int at = p.at;
  const int n_saved = max_slot(gs_in, G_in) + 1;
  Regions R;
  V2S_TRY(resolve_regions(p, ws, ws_bytes, 0, n_saved, &R));
  const int64_t M = p.M, MP = p.MP;
  auto sb = [&](int g, int64_t off) -> char* { return R.stash[gs[g].slot] + off; };
  auto bb = [&](int g, int64_t off) -> char* { return R.bwd[gs[g].slot] + off; };

  float* dx[MAXG]; void* dxlp[MAXG]; void* big[MAXG]; void* tmp[MAXG];
  {
    const float *df[MAXG], *dh[MAXG]; int64_t dfs[MAXG]; void* lp[MAXG];
    for (int g = 0; g < G; ++g) {
      if (layer_hi == NL && !gs[g].dfeat && !gs[g].dhidden) { set_error("backbone_backward: group needs dfeat or dhidden"); return 1; }
      dx[g] = (float*)bb(g, p.b_dx);
      dxlp[g] = at ? (void*)bb(g, p.b_dxlp) : (void*)dx[g];
      big[g] = bb(g, p.b_big); tmp[g] = bb(g, p.b_tmp);
      df[g] = gs[g].dfeat; dfs[g] = gs[g].dfeat_stride > 0 ? gs[g].dfeat_stride : D; dh[g] = gs[g].dhidden;
      lp[g] = at ? dxlp[g] : nullptr;
    }
    if (layer_hi == NL) V2S_TRY(launch_pool_bwd(df, dfs, dh, dx, lp, G, B, at, st));
  }

  const int split = wgrad_split(M);
  // dW[N_out, K_in] += dY[M, N_out]^T * X[M, K_in]   (token dimension is the reduction)
  // bias_goff >= 0: also accumulate the bias gradient (column sums of dY) — inside the tensor-core wgrad
  // kernel (extra N=16 MMA against a ones tile), or with the column-sum kernel on the SIMT path
  const bool tc = at != AT_F32 && tc_enabled();
  const bool mlp_fuse = tc && !getenv("V2S_NO_MLPFUSE") && !getenv("V2S_NO_LNFUSE");
  auto wgrad = [&](void* const* dy, int n_out, void* const* x, int k_in, int64_t goff, bool remap,
                   int64_t bias_goff = -1) -> int {
    if (bias_goff >= 0 && !tc) {
      const void* src[MAXG]; float* dst[MAXG];
      for (int g = 0; g < G; ++g) { src[g] = dy[g]; dst[g] = gs[g].grads + bias_goff; }
      prof::Scope sc(prof::C_COLSUM, (double)M * n_out * p.es * G, st);
      V2S_TRY(launch_colsum(src, dst, G, (int)M, n_out, at, st));
    }
    GemmDesc d = make_gemm_desc();
    d.M = n_out; d.N = k_in; d.K = remap ? (int)MP : (int)M; d.groups = G;
    d.a_rs = 1; d.a_cs = n_out; d.b_rs = k_in; d.b_cs = 1;
    d.a_remap = remap ? 2 : 0;
    d.epi = EPI_ACCUM; d.ldc = k_in; d.split_k = split;
    for (int g = 0; g < G; ++g) {
      d.A[g] = dy[g]; d.B[g] = x[g]; d.out[g] = gs[g].grads + goff;
      d.rowsum_out[g] = (bias_goff >= 0 && tc) ? gs[g].grads + bias_goff : nullptr;
    }
    return run_gemm(d, at, at, 0, st, prof::C_WGRAD);
  };
  // two weight gradients of one block in ONE split-K launch (tensor-core path): the groups of the second problem
  // follow those of the first (GemmDesc::M2 / N2).  Four ~20 us launches per block become two: one ramp / one tail,
  // longer K ranges per CTA and a third less reduce-add traffic.
  static const bool merge_wgrads = !getenv("V2S_NO_WGRAD_MERGE");
  auto wgrad2 = [&](void* const* dy1, int n_out1, void* const* x1, int k_in1, int64_t goff1, int64_t bias1,
                    void* const* dy2, int n_out2, void* const* x2, int k_in2, int64_t goff2, int64_t bias2) -> int {
    if (!tc || !merge_wgrads || 2 * G > MAXG) {
      V2S_TRY(wgrad(dy1, n_out1, x1, k_in1, goff1, false, bias1));
      return wgrad(dy2, n_out2, x2, k_in2, goff2, false, bias2);
    }
    GemmDesc d = make_gemm_desc();
    d.M = n_out1; d.N = k_in1; d.K = (int)M; d.groups = 2 * G;
    d.a_rs = 1; d.a_cs = n_out1; d.b_rs = k_in1; d.b_cs = 1;
    d.M2 = n_out2; d.N2 = k_in2;
    d.epi = EPI_ACCUM; d.ldc = k_in1; d.split_k = split;
    for (int g = 0; g < G; ++g) {
      d.A[g] = dy1[g]; d.B[g] = x1[g]; d.out[g] = gs[g].grads + goff1;
      d.rowsum_out[g] = bias1 >= 0 ? gs[g].grads + bias1 : nullptr;
      d.A[G + g] = dy2[g]; d.B[G + g] = x2[g]; d.out[G + g] = gs[g].grads + goff2;
      d.rowsum_out[G + g] = bias2 >= 0 ? gs[g].grads + bias2 : nullptr;
    }
    return run_gemm(d, at, at, 0, st, prof::C_WGRAD);
  };
  const bool merged = tc && merge_wgrads && 2 * G <= MAXG;
  // dX[M, K_in] = dY[M, N_out] * W[N_out, K_in]
  // late: the kernel launched just before this one is a wgrad whose inputs this GEMM shares, and nothing the wgrad
  // touches is written here (GemmDesc::late_wait)
  auto dgrad = [&](void* const* dy, int n_out, int64_t woff, int k_in, void* const* out, int epi,
                   void* const* aux, bool late = false) -> int {
    GemmDesc d = make_gemm_desc();
    d.M = (int)M; d.N = k_in; d.K = n_out; d.groups = G;
    d.a_rs = n_out; d.a_cs = 1; d.b_rs = k_in; d.b_cs = 1;
    d.epi = epi; d.ldc = k_in; d.late_wait = late ? 1 : 0;
    for (int g = 0; g < G; ++g) {
      d.A[g] = dy[g]; d.B[g] = weight_ptr(gs[g], at, woff); d.out[g] = out[g];
      d.aux[g] = aux ? aux[g] : nullptr;
    }
    return run_gemm(d, at, at, at, st, prof::C_DGRAD);
  };
  auto bias_grad = [&](void* const* dy, int n, int64_t goff) -> int {
    const void* src[MAXG]; float* dst[MAXG];
    for (int g = 0; g < G; ++g) { src[g] = dy[g]; dst[g] = gs[g].grads + goff; }
    prof::Scope sc(prof::C_COLSUM, (double)M * n * p.es * G, st);
    return launch_colsum(src, dst, G, (int)M, n, at, st);
  };

  for (int l = layer_hi - 1; l >= layer_lo; --l) {
    const int64_t lo = layer_off(l);
    const LayerStash& s = p.s_layer[l];
    void *u[MAXG], *h[MAXG], *xn2[MAXG], *ctx[MAXG], *qkv[MAXG], *xn1[MAXG];
    for (int g = 0; g < G; ++g) {
      u[g] = sb(g, s.u); h[g] = sb(g, s.h); xn2[g] = sb(g, s.xn2);
      ctx[g] = sb(g, s.ctx); qkv[g] = sb(g, s.qkv); xn1[g] = sb(g, s.xn1);
    }
    // Order: the backward chain is dgrad(W2) -> dgrad(W1) -> LN2' -> dgrad(Wo) -> attention' -> dgrad(Wqkv) -> LN1';
    // the wgrads are leaves.  Each dgrad is launched right after the wgrad that reads the same (older) gradient and
    // runs "late" (GemmDesc::late_wait): it does not drain that wgrad.  The mirrored order (wgrads late after the
    // chain kernels, dWo filling the attention kernel's partial last wave) measured the same or slightly slower.
    // ---- MLP ----
    if (!(merged && mlp_fuse)) V2S_TRY(wgrad(dxlp, D, h, DF, lo + L_W2, false));      // dW2 [192,768]
    if (l == NL - 1) V2S_TRY(bias_grad(dxlp, D, lo + L_B2));               // lower blocks: fused into LN1-bwd above
    if (mlp_fuse) {
      // du = (dx W2) * gelu'(u) and d xn2 = du W1 chained on chip (du is still written once: dW1 needs it)
      MlpDesc d = make_mlp_desc();
      d.mode = MLP_BWD; d.M = (int)M; d.groups = G; d.lp_f16 = at == AT_F16 ? 1 : 0; d.late_wait = (!merged && l != NL - 1 && !getenv("V2S_NO_LATE_WAIT")) ? 1 : 0;
      for (int g = 0; g < G; ++g) {
        d.a[g] = dxlp[g]; d.w1[g] = weight_ptr(gs[g], at, lo + L_W1); d.w2[g] = weight_ptr(gs[g], at, lo + L_W2);
        d.u[g] = u[g]; d.h[g] = big[g]; d.out[g] = tmp[g];
      }
      {
        prof::Scope scope(prof::C_MLP_B, 4.0 * M * D * (double)DF * G, st,
                          ((double)M * D * 4 + (double)M * DF * 4 + 2.0 * D * DF * 2) * G);
        V2S_TRY(launch_mlp_tc(d, st));
      }
      if (merged)      // dW2 [192,768] (the gradient dxlp is only overwritten by LN2-bwd below) with dW1 [768,192], d b1
        V2S_TRY(wgrad2(dxlp, D, h, DF, lo + L_W2, -1, big, DF, xn2, D, lo + L_W1, lo + L_B1));
      else
        V2S_TRY(wgrad(big, DF, xn2, D, lo + L_W1, false, lo + L_B1));          // dW1 [768,192] and d b1
    } else {
    V2S_TRY(dgrad(dxlp, D, lo + L_W2, DF, big, EPI_DGELU, u, l != NL - 1));   // du = (dx W2) * gelu'(u)
    V2S_TRY(wgrad(big, DF, xn2, D, lo + L_W1, false, lo + L_B1));          // dW1 [768,192] and d b1
    V2S_TRY(dgrad(big, DF, lo + L_W1, D, tmp, EPI_STORE, nullptr, true));  // d xn2
    }
    {
      const void* dy[MAXG]; const float *x[MAXG], *mu[MAXG], *rs[MAXG], *gm[MAXG];
      float *dg[MAXG], *db[MAXG], *cs[MAXG]; void* lp[MAXG];
      for (int g = 0; g < G; ++g) {
        dy[g] = tmp[g]; x[g] = (const float*)sb(g, s.x_mid); mu[g] = (const float*)sb(g, s.mean2);
        rs[g] = (const float*)sb(g, s.rstd2); gm[g] = gs[g].params + lo + L_LN2W;
        dg[g] = gs[g].grads + lo + L_LN2W; db[g] = gs[g].grads + lo + L_LN2B; lp[g] = at ? dxlp[g] : nullptr;
        cs[g] = gs[g].grads + lo + L_BO;          // d b_o = column sums of the gradient w.r.t. x_mid
      }
      { prof::Scope sc(prof::C_LN_B, (double)M * D * (12 + 2 * p.es) * G, st);
        V2S_TRY(launch_ln_bwd(dy, x, mu, rs, gm, dx, lp, dg, db, cs, G, (int)M, at, st)); }
    }
    // ---- attention ----
    if (!merged) V2S_TRY(wgrad(dxlp, D, ctx, D, lo + L_WO, false));        // dWo (d b_o: fused into LN2-bwd)
    V2S_TRY(dgrad(dxlp, D, lo + L_WO, D, tmp, EPI_STORE, nullptr, !merged));  // d ctx
    {
      const void *cq[MAXG], *cc[MAXG], *cd[MAXG]; const float* ls[MAXG];
      for (int g = 0; g < G; ++g) { cq[g] = qkv[g]; cc[g] = ctx[g]; cd[g] = tmp[g]; ls[g] = (const float*)sb(g, s.lse); }
      V2S_TRY(launch_attention_bwd(cq, cc, ls, cd, big, G, B, at, st));    // d qkv in `big` [M,576]
    }
    if (merged)        // dWo [192,192] (its gradient dxlp is only overwritten by LN1-bwd below) with dWqkv [576,192], d b_qkv
      V2S_TRY(wgrad2(dxlp, D, ctx, D, lo + L_WO, -1, big, 3 * D, xn1, D, lo + L_WQKV, lo + L_BQKV));
    else
      V2S_TRY(wgrad(big, 3 * D, xn1, D, lo + L_WQKV, false, lo + L_BQKV));   // dWqkv [576,192] and d b_qkv
    V2S_TRY(dgrad(big, 3 * D, lo + L_WQKV, D, tmp, EPI_STORE, nullptr, true));   // d xn1
    {
      const void* dy[MAXG]; const float *x[MAXG], *mu[MAXG], *rs[MAXG], *gm[MAXG];
      float *dg[MAXG], *db[MAXG], *cs[MAXG]; void* lp[MAXG];
      for (int g = 0; g < G; ++g) {
        dy[g] = tmp[g]; x[g] = (const float*)sb(g, p.s_x[l]); mu[g] = (const float*)sb(g, s.mean1);
        rs[g] = (const float*)sb(g, s.rstd1); gm[g] = gs[g].params + lo + L_LN1W;
        dg[g] = gs[g].grads + lo + L_LN1W; db[g] = gs[g].grads + lo + L_LN1B; lp[g] = at ? dxlp[g] : nullptr;
        // d b_2 of the block below = column sums of the gradient w.r.t. this block's input
        cs[g] = l > 0 ? gs[g].grads + layer_off(l - 1) + L_B2 : nullptr;
      }
      { prof::Scope sc(prof::C_LN_B, (double)M * D * (12 + 2 * p.es) * G, st);
        V2S_TRY(launch_ln_bwd(dy, x, mu, rs, gm, dx, lp, dg, db, cs, G, (int)M, at, st)); }
    }
  }
  // ---- embeddings: d pos, d cls, d patch bias, d patch weight ----
  if (layer_lo == 0) {
    const float* cdx[MAXG]; float* gr[MAXG]; void* pt[MAXG];
    for (int g = 0; g < G; ++g) {
      cdx[g] = dx[g]; gr[g] = gs[g].grads;
      pt[g] = gs[g].x_format == 1 ? const_cast<void*>(gs[g].x) : static_cast<void*>(sb(g, p.s_patches));
    }
    V2S_TRY(launch_embed_bwd(cdx, gr, G, B, st));
    if (tc) {          // compact the patch rows, then the ordinary tensor-core wgrad
      const void* src[MAXG];
      for (int g = 0; g < G; ++g) src[g] = dxlp[g];
      V2S_TRY(launch_gather_patch_rows(src, tmp, G, B, at, st));
      GemmDesc d = make_gemm_desc();
      d.M = D; d.N = KPE; d.K = (int)MP; d.groups = G;
      d.a_rs = 1; d.a_cs = D; d.b_rs = KPE; d.b_cs = 1; d.epi = EPI_ACCUM; d.ldc = KPE; d.split_k = split;
      for (int g = 0; g < G; ++g) { d.A[g] = tmp[g]; d.B[g] = pt[g]; d.out[g] = gs[g].grads + OFF_WPE; }
      V2S_TRY(run_gemm(d, at, at, 0, st, prof::C_WGRAD));
    } else {
      V2S_TRY(wgrad(dxlp, D, pt, KPE, OFF_WPE, true));
    }
  }
  return 0;
}

// =============================================================================================
// heads + loss
// =============================================================================================
namespace {
struct HeadsBufs {
  float *a1, *y1, *a1t, *y1t, *dy1, *z, *y2, *pr, *zt, *dp, *dy2, *dz, *y2b, *part;
};
int heads_bufs(void* ws, int64_t ws_bytes, int B, HeadsBufs* hb) {
  const int64_t need = (int64_t)B * 8192 * 4;
  if (!ws || ws_bytes < need) {
    set_error("heads: workspace too small (%lld < %lld)", (long long)ws_bytes, (long long)need);
    return 1;
  }
  float* w = static_cast<float*>(ws);
  const int PH = V2S_PROJ_HID, PO = V2S_PROJ_OUT;
  hb->a1 = w;  w += (int64_t)B * PH;   // relu(f W1^T + b1)
  hb->y1 = w;  w += (int64_t)B * PH;   // a1 * dropout mask
  hb->a1t = w; w += (int64_t)B * PH;
  hb->y1t = w; w += (int64_t)B * PH;
  hb->dy1 = w; w += (int64_t)B * PH;
  hb->z = w;   w += (int64_t)B * PO;
  hb->y2 = w;  w += (int64_t)B * PO;
  hb->pr = w;  w += (int64_t)B * PO;
  hb->zt = w;  w += (int64_t)B * PO;
  hb->dp = w;  w += (int64_t)B * PO;
  hb->dy2 = w; w += (int64_t)B * PO;
  hb->dz = w;  w += (int64_t)B * PO;
  hb->y2b = w; w += (int64_t)B * PO;
  hb->part = w; w += (int64_t)B * PO * 8;   // split-K partial slabs
  return 0;
}
}  // namespace

// projection_head + prediction_head on the online features, projection_head on the target
// features (ref:ssp_vit2spn_tiny.py:153-158).  Intermediates stay in the workspace heads region.
static int heads_forward_impl(const float* hp, const float* fo, const float* ft, const float* mo, const float* mt,
                              float* pred_out, float* tgt_out, int B, void* ws, int64_t ws_bytes, cudaStream_t st) {
  if (!hp || !fo || !ft) { set_error("heads_forward: null argument"); return 1; }
  HeadsBufs hb;
  V2S_TRY(heads_bufs(ws, ws_bytes, B, &hb));
  const int PH = V2S_PROJ_HID, PO = V2S_PROJ_OUT, PI = V2S_PROJ_IN;
  auto linear = [&](const float* x, int k, int64_t woff, int64_t boff, int n, int epi, float* out, float* out2,
                    const float* mask) -> int {
    GemmDesc d = make_gemm_desc();
    d.M = B; d.N = n; d.K = k; d.a_rs = k; d.a_cs = 1; d.b_rs = 1; d.b_cs = k; d.epi = epi; d.ldc = n;
    d.A[0] = x; d.B[0] = hp + woff; d.bias[0] = hp + boff; d.out[0] = out; d.out2[0] = out2; d.mask[0] = mask;
    return launch_gemm_simt(d, 0, 0, 0, st);
  };
  // K = 1024 with a [B,128] output: too few tiles for a serial K loop → 8 K-splits into partial slabs, then a
  // fixed-order reduction that also adds the bias (deterministic, unlike atomics)
  auto linear_splitk = [&](const float* x, int k, int64_t woff, int64_t boff, int n, float* out) -> int {
    GemmDesc d = make_gemm_desc();
    d.M = B; d.N = n; d.K = k; d.a_rs = k; d.a_cs = 1; d.b_rs = 1; d.b_cs = k; d.epi = EPI_STORE; d.ldc = n;
    d.split_k = 8; d.split_stride = (int64_t)B * n;
    d.A[0] = x; d.B[0] = hp + woff; d.out[0] = hb.part;
    V2S_TRY(launch_gemm_simt(d, 0, 0, 0, st));
    return launch_splitk_reduce(out, hb.part, hp + boff, B, n, 8, st);
  };
  V2S_TRY(linear(fo, PI, H_W1, H_B1, PH, EPI_BIAS_RELU_MASK, hb.a1, hb.y1, mo));
  V2S_TRY(linear_splitk(hb.y1, PH, H_W2, H_B2, PO, hb.z));
  V2S_TRY(linear(hb.z, PO, H_W3, H_B3, PO, EPI_BIAS_RELU_MASK, hb.y2, hb.y2b, nullptr));
  V2S_TRY(linear(hb.y2, PO, H_W4, H_B4, PO, EPI_STORE, hb.pr, nullptr, nullptr));
  V2S_TRY(linear(ft, PI, H_W1, H_B1, PH, EPI_BIAS_RELU_MASK, hb.a1t, hb.y1t, mt));
  V2S_TRY(linear_splitk(hb.y1t, PH, H_W2, H_B2, PO, hb.zt));
  if (pred_out) V2S_CUDA_OK(cudaMemcpyAsync(pred_out, hb.pr, (size_t)B * PO * 4, cudaMemcpyDeviceToDevice, st));
  if (tgt_out) V2S_CUDA_OK(cudaMemcpyAsync(tgt_out, hb.zt, (size_t)B * PO * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// autograd backward of the online branch of the heads: dpred [B,128] -> head grads (+=), dfeat_online
static int heads_backward_impl(const float* hp, float* hg, const float* fo, const float* mo, const float* dpred,
                               float* dfo, int B, void* ws, int64_t ws_bytes, cudaStream_t st) {
  if (!hp || !hg || !fo || !dfo) { set_error("heads_backward: null argument"); return 1; }
  HeadsBufs hb;
  V2S_TRY(heads_bufs(ws, ws_bytes, B, &hb));
  const float* dp = dpred ? dpred : hb.dp;
  const int PH = V2S_PROJ_HID, PO = V2S_PROJ_OUT, PI = V2S_PROJ_IN;
  auto wgrad = [&](const float* dy, int n, const float* x, int k, int64_t woff) -> int {
    GemmDesc d = make_gemm_desc();
    d.M = n; d.N = k; d.K = B; d.a_rs = 1; d.a_cs = n; d.b_rs = k; d.b_cs = 1; d.epi = EPI_ACCUM; d.ldc = k;
    d.A[0] = dy; d.B[0] = x; d.out[0] = hg + woff;
    return launch_gemm_simt(d, 0, 0, 0, st);
  };
  auto bgrad = [&](const float* dy, int n, int64_t boff) -> int {
    const void* s[1] = {dy}; float* o[1] = {hg + boff};
    return launch_colsum(s, o, 1, B, n, 0, st);
  };
  auto dgrad = [&](const float* dy, int n, int64_t woff, int k, float* out, int epi, const float* relu_src,
                   const float* mask) -> int {
    GemmDesc d = make_gemm_desc();
    d.M = B; d.N = k; d.K = n; d.a_rs = n; d.a_cs = 1; d.b_rs = k; d.b_cs = 1; d.epi = epi; d.ldc = k;
    d.A[0] = dy; d.B[0] = hp + woff; d.out[0] = out; d.aux[0] = relu_src; d.mask[0] = mask;
    return launch_gemm_simt(d, 0, 0, 0, st);
  };
  V2S_TRY(wgrad(dp, PO, hb.y2, PO, H_W4));
  V2S_TRY(bgrad(dp, PO, H_B4));
  V2S_TRY(dgrad(dp, PO, H_W4, PO, hb.dy2, EPI_DRELU_MASK, hb.y2, nullptr));
  V2S_TRY(wgrad(hb.dy2, PO, hb.z, PO, H_W3));
  V2S_TRY(bgrad(hb.dy2, PO, H_B3));
  V2S_TRY(dgrad(hb.dy2, PO, H_W3, PO, hb.dz, EPI_STORE, nullptr, nullptr));
  V2S_TRY(wgrad(hb.dz, PO, hb.y1, PH, H_W2));
  V2S_TRY(bgrad(hb.dz, PO, H_B2));
  V2S_TRY(dgrad(hb.dz, PO, H_W2, PH, hb.dy1, EPI_DRELU_MASK, hb.a1, mo));
  V2S_TRY(wgrad(hb.dy1, PH, fo, PI, H_W1));
  V2S_TRY(bgrad(hb.dy1, PH, H_B1));
  {   // d feat = dy1 W1: K = 1024 → split-K accumulate into the zeroed [B,384] output
    V2S_CUDA_OK(cudaMemsetAsync(dfo, 0, (size_t)B * PI * sizeof(float), st));
    GemmDesc d = make_gemm_desc();
    d.M = B; d.N = PI; d.K = PH; d.a_rs = PH; d.a_cs = 1; d.b_rs = PI; d.b_cs = 1; d.epi = EPI_ACCUM; d.ldc = PI; d.split_k = 8;
    d.A[0] = hb.dy1; d.B[0] = hp + H_W1; d.out[0] = dfo;
    V2S_TRY(launch_gemm_simt(d, 0, 0, 0, st));
  }
  return 0;
}

static int heads_impl(const float* hp, float* hg, const float* fo, const float* ft, const float* mo, const float* mt,
                      float* dfo, float* pred_out, float* tgt_out, float* loss, int B, int accum, float grad_scale,
                      int with_backward, void* ws, int64_t ws_bytes, cudaStream_t st, const float* grad_scale_dev = nullptr) {
  if (!loss) { set_error("heads: null loss"); return 1; }
  V2S_TRY(heads_forward_impl(hp, fo, ft, mo, mt, pred_out, tgt_out, B, ws, ws_bytes, st));
  HeadsBufs hb;
  V2S_TRY(heads_bufs(ws, ws_bytes, B, &hb));
  V2S_TRY(launch_cosine_loss(hb.pr, hb.zt, loss, with_backward ? hb.dp : nullptr, B, accum, grad_scale, st, grad_scale_dev));
  if (!with_backward) return 0;
  return heads_backward_impl(hp, hg, fo, mo, nullptr, dfo, B, ws, ws_bytes, st);
}

}  // namespace v2s

// =============================================================================================
// extern "C" boundary
// =============================================================================================
using namespace v2s;

extern "C" {

int v2s_abi_version(void) { return V2S_ABI_VERSION; }
const char* v2s_last_error(void) { return g_err; }

int v2s_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) { set_error("no CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(e)); return 1; }
  if (device < 0 || device >= n) { set_error("device %d out of range (%d devices)", device, n); return 1; }
  cudaDeviceProp prop;
  V2S_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
    return 1;
  }
  // per-device state is created with `device` current; the caller's current device is left as it was
  int prev = 0;
  V2S_CUDA_OK(cudaGetDevice(&prev));
  V2S_CUDA_OK(cudaSetDevice(device));
  const int rc = gemm_tc_init();
  cudaSetDevice(prev);
  return rc;
}

int64_t v2s_backbone_numel(void) { return BACKBONE_NUMEL; }
int64_t v2s_backbone_active_numel(void) { return OFF_ACTIVE_END; }
int64_t v2s_heads_numel(void) { return HEADS_NUMEL; }

int v2s_backbone_layout(int64_t* o) {
  if (!o) { set_error("null"); return 1; }
  int i = 0;
  o[i++] = OFF_CLS; o[i++] = OFF_POS; o[i++] = OFF_WPE; o[i++] = OFF_BPE;
  for (int l = 0; l < NL; ++l) {
    const int64_t lo = layer_off(l);
    o[i++] = lo + L_WQKV;               o[i++] = lo + L_BQKV;            // query
    o[i++] = lo + L_WQKV + D * D;       o[i++] = lo + L_BQKV + D;        // key
    o[i++] = lo + L_WQKV + 2 * D * D;   o[i++] = lo + L_BQKV + 2 * D;    // value
    o[i++] = lo + L_WO; o[i++] = lo + L_BO;
    o[i++] = lo + L_W1; o[i++] = lo + L_B1;
    o[i++] = lo + L_W2; o[i++] = lo + L_B2;
    o[i++] = lo + L_LN1W; o[i++] = lo + L_LN1B; o[i++] = lo + L_LN2W; o[i++] = lo + L_LN2B;
  }
  o[i++] = OFF_LNF_W; o[i++] = OFF_LNF_B; o[i++] = OFF_POOL_W; o[i++] = OFF_POOL_B;
  return i == 200 ? 0 : 1;
}

int v2s_heads_layout(int64_t* o) {
  if (!o) { set_error("null"); return 1; }
  o[0] = H_W1; o[1] = H_B1; o[2] = H_W2; o[3] = H_B2; o[4] = H_W3; o[5] = H_B3; o[6] = H_W4; o[7] = H_B4;
  return 0;
}

int64_t v2s_workspace_bytes(int batch, int mode, int n_groups, int n_saved) {
  if (batch < 1 || n_groups < 0 || n_groups > MAXG || n_saved < 0 || n_saved > MAXG) return -1;
  return plan_total(make_plan(batch, mode), n_groups, n_saved);
}

int v2s_backbone_forward(const v2s_group_t* groups, int n_groups, int batch, int mode, void* workspace,
                         int64_t workspace_bytes, void* stream) {
  if (!groups) { set_error("null groups"); return 1; }
  return backbone_forward_impl(groups, n_groups, batch, mode, workspace, workspace_bytes, (cudaStream_t)stream);
}

int v2s_backbone_backward(const v2s_group_t* groups, int n_groups, int batch, int mode, void* workspace,
                          int64_t workspace_bytes, void* stream) {
  if (!groups) { set_error("null groups"); return 1; }
  return backbone_backward_impl(groups, n_groups, batch, mode, workspace, workspace_bytes, (cudaStream_t)stream);
}

int v2s_backbone_backward_range(const v2s_group_t* groups, int n_groups, int batch, int mode, void* workspace,
                                int64_t workspace_bytes, int layer_hi, int layer_lo, void* stream) {
  if (!groups) { set_error("null groups"); return 1; }
  return backbone_backward_impl(groups, n_groups, batch, mode, workspace, workspace_bytes, (cudaStream_t)stream, layer_hi,
                                layer_lo);
}

int v2s_heads_loss_fwd_bwd(const float* head_params, float* head_grads, const float* feat_online,
                           const float* feat_target, const float* mask_online, const float* mask_target,
                           float* dfeat_online, float* pred, float* target_proj, float* loss, int batch,
                           int accumulation_steps, float grad_scale, int with_backward, void* workspace,
                           int64_t workspace_bytes, void* stream) {
  if (batch < 1 || accumulation_steps < 1) { set_error("heads: bad batch/accumulation_steps"); return 1; }
  return heads_impl(head_params, head_grads, feat_online, feat_target, mask_online, mask_target, dfeat_online, pred,
                    target_proj, loss, batch, accumulation_steps, grad_scale, with_backward, workspace,
                    workspace_bytes, (cudaStream_t)stream);
}

int v2s_heads_loss_fwd_bwd_amp(const float* head_params, float* head_grads, const float* feat_online,
                               const float* feat_target, const float* mask_online, const float* mask_target,
                               float* dfeat_online, float* pred, float* target_proj, float* loss, int batch,
                               int accumulation_steps, const float* grad_scale_dev, int with_backward, void* workspace,
                               int64_t workspace_bytes, void* stream) {
  if (batch < 1 || accumulation_steps < 1) { set_error("heads: bad batch/accumulation_steps"); return 1; }
  return heads_impl(head_params, head_grads, feat_online, feat_target, mask_online, mask_target, dfeat_online, pred,
                    target_proj, loss, batch, accumulation_steps, 1.0f, with_backward, workspace, workspace_bytes,
                    (cudaStream_t)stream, grad_scale_dev);
}

int v2s_heads_forward(const float* head_params, const float* feat_online, const float* feat_target,
                      const float* mask_online, const float* mask_target, float* pred, float* target_proj, int batch,
                      void* workspace, int64_t workspace_bytes, void* stream) {
  if (batch < 1) { set_error("heads_forward: bad batch"); return 1; }
  return heads_forward_impl(head_params, feat_online, feat_target, mask_online, mask_target, pred, target_proj, batch,
                            workspace, workspace_bytes, (cudaStream_t)stream);
}

int v2s_heads_backward(const float* head_params, float* head_grads, const float* feat_online,
                       const float* mask_online, const float* dpred, float* dfeat_online, int batch, void* workspace,
                       int64_t workspace_bytes, void* stream) {
  if (batch < 1 || !dpred) { set_error("heads_backward: bad argument"); return 1; }
  return heads_backward_impl(head_params, head_grads, feat_online, mask_online, dpred, dfeat_online, batch, workspace,
                             workspace_bytes, (cudaStream_t)stream);
}

int v2s_cosine_loss(const float* pred, const float* target_proj, float* loss, float* dpred, int batch,
                    int accumulation_steps, float grad_scale, void* stream) {
  if (!pred || !target_proj || !loss || batch < 1 || accumulation_steps < 1) { set_error("cosine_loss: bad argument"); return 1; }
  return launch_cosine_loss(pred, target_proj, loss, dpred, batch, accumulation_steps, grad_scale, (cudaStream_t)stream);
}

int v2s_infonce_loss(const float* pred, const float* keys, float* loss, float* row_loss, float* dpred, int batch,
                     int n_keys, int64_t label_offset, float temperature, int accumulation_steps, float grad_scale,
                     const float* grad_scale_dev, void* stream) {
  if (!pred || !keys || !loss || !row_loss || batch < 1 || accumulation_steps < 1) { set_error("infonce_loss: bad argument"); return 1; }
  return launch_infonce_loss(pred, keys, loss, row_loss, dpred, batch, n_keys, label_offset, temperature, accumulation_steps,
                             grad_scale, (cudaStream_t)stream, grad_scale_dev);
}

int v2s_dropout_mask(float* mask, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream) {
  if (!mask || n < 0 || p < 0.f || p >= 1.f) { set_error("dropout_mask: bad argument"); return 1; }
  if (n == 0) return 0;
  return launch_dropout_mask(mask, n, p, seed, offset, (cudaStream_t)stream);
}

int v2s_adam_step(const v2s_range_t* ranges, int n_ranges, int64_t step, double lr, double beta1, double beta2,
                  double eps, double weight_decay, double grad_scale, void* stream) {
  if (!ranges || step < 1) { set_error("adam: bad argument"); return 1; }
  for (int i = 0; i < n_ranges; ++i)
    if ((reinterpret_cast<uintptr_t>(ranges[i].params) | reinterpret_cast<uintptr_t>(ranges[i].grads) |
         reinterpret_cast<uintptr_t>(ranges[i].exp_avg) | reinterpret_cast<uintptr_t>(ranges[i].exp_avg_sq)) & 15) {
      set_error("adam: range %d is not 16-byte aligned", i);
      return 1;
    }
  return launch_adam(ranges, n_ranges, step, lr, beta1, beta2, eps, weight_decay, grad_scale, (cudaStream_t)stream);
}

static int check_ranges(const v2s_range_t* ranges, int n_ranges) {
  for (int i = 0; i < n_ranges; ++i)
    if ((reinterpret_cast<uintptr_t>(ranges[i].params) | reinterpret_cast<uintptr_t>(ranges[i].grads) |
         reinterpret_cast<uintptr_t>(ranges[i].exp_avg) | reinterpret_cast<uintptr_t>(ranges[i].exp_avg_sq)) & 15) {
      set_error("adam: range %d is not 16-byte aligned", i);
      return 1;
    }
  return 0;
}

int v2s_adam_step_lp(const v2s_range_t* ranges, int n_ranges, int64_t step, double lr, double beta1, double beta2,
                     double eps, double weight_decay, double grad_scale, int lp_format, void* stream) {
  if (!ranges || step < 1) { set_error("adam: bad argument"); return 1; }
  V2S_TRY(check_ranges(ranges, n_ranges));
  return launch_adam(ranges, n_ranges, step, lr, beta1, beta2, eps, weight_decay, grad_scale, (cudaStream_t)stream,
                     lp_format == V2S_LP_FP16);
}

int v2s_adam_step_amp(const v2s_range_t* ranges, int n_ranges, float* state8, double lr, double beta1, double beta2,
                      double eps, double weight_decay, double grad_multiplier, const float* grad_scale_dev,
                      const float* found_inf_dev, int lp_format, int advance_step, void* stream) {
  if (!ranges || !state8) { set_error("adam_amp: bad argument"); return 1; }
  V2S_TRY(check_ranges(ranges, n_ranges));
  return launch_adam_amp(ranges, n_ranges, state8, lr, beta1, beta2, eps, weight_decay, grad_multiplier, grad_scale_dev,
                         found_inf_dev, lp_format == V2S_LP_FP16, advance_step, (cudaStream_t)stream);
}

int v2s_ema_update(float* const* targets, const float* const* onlines, void* const* targets_lp, int n_pairs,
                   int64_t numel, double momentum, void* stream) {
  if (!targets || !onlines) { set_error("ema: null"); return 1; }
  return launch_ema(targets, onlines, targets_lp, n_pairs, numel, momentum, (cudaStream_t)stream);
}

int v2s_ema_update_lp(float* const* targets, const float* const* onlines, void* const* targets_lp, int n_pairs,
                      int64_t numel, double momentum, int lp_format, void* stream) {
  if (!targets || !onlines) { set_error("ema: null"); return 1; }
  return launch_ema(targets, onlines, targets_lp, n_pairs, numel, momentum, (cudaStream_t)stream, lp_format == V2S_LP_FP16);
}

int v2s_cast_lp(const float* src, void* dst, int64_t numel, int lp_format, void* stream) {
  if (!src || !dst || numel < 0) { set_error("cast: bad argument"); return 1; }
  if (numel == 0) return 0;
  return launch_cast_bf16(src, dst, numel, (cudaStream_t)stream, lp_format == V2S_LP_FP16);
}

int v2s_cast_bf16(const float* src, void* dst, int64_t numel, void* stream) {
  if (!src || !dst || numel < 0) { set_error("cast: bad argument"); return 1; }
  if (numel == 0) return 0;
  return launch_cast_bf16(src, dst, numel, (cudaStream_t)stream);
}

int v2s_preprocess_u8_patches(const uint8_t* src, void* patch_rows, int n_images, int lp_format, void* stream) {
  if (!src || !patch_rows || n_images < 1 || (lp_format != 0 && lp_format != 1)) { set_error("preprocess_patches: bad argument"); return 1; }
  return launch_preprocess_u8_patches(src, patch_rows, n_images, lp_format, (cudaStream_t)stream);
}

int v2s_preprocess_u8(const uint8_t* src, float* dst, int batch, void* stream) {
  if (!src || !dst || batch < 1) { set_error("preprocess: bad argument"); return 1; }
  return launch_preprocess_u8(src, dst, batch, (cudaStream_t)stream);
}

int v2s_augment_finish_u8(const uint8_t* src, int n, int in_size, const int32_t* bounds, const int32_t* coefs, int ksize,
                          const float* k1d, const int32_t* erase, const float* host_mean3, const float* host_std3,
                          float* dst, void* stream) {
  if (!src || !bounds || !coefs || !host_mean3 || !host_std3 || !dst || n < 1) { set_error("augment: bad argument"); return 1; }
  return launch_augment_finish(src, n, in_size, bounds, coefs, ksize, k1d, erase, host_mean3, host_std3, dst,
                               (cudaStream_t)stream);
}

int64_t v2s_launch_count(void) { return g_launch_count; }

int v2s_set_sm_limit(int n_sms) {
  tc_set_sm_limit(n_sms);
  return 0;
}

int v2s_prof_enable(int on) {
  prof::enabled = on != 0;
  prof::n_recs = 0;
  return 0;
}

// writes one line per kernel class: "<name> <launches> <total_ms> <work> <bytes>" (work = FLOPs for
// GEMM/attention classes, bytes for the memory-bound ones; bytes = algorithmic HBM bytes); synchronises.
int v2s_prof_report(char* host_buf, int64_t buf_bytes) {
  if (!host_buf || buf_bytes < 64) { set_error("prof_report: buffer too small"); return 1; }
  V2S_CUDA_OK(cudaDeviceSynchronize());
  double ms[prof::C_COUNT] = {0}, work[prof::C_COUNT] = {0}, byt[prof::C_COUNT] = {0};
  int cnt[prof::C_COUNT] = {0};
  for (int i = 0; i < prof::n_recs; ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, prof::recs[i].a, prof::recs[i].b) != cudaSuccess) continue;
    ms[prof::recs[i].cls] += t; work[prof::recs[i].cls] += prof::recs[i].work; cnt[prof::recs[i].cls]++;
    byt[prof::recs[i].cls] += prof::recs[i].bytes > 0 ? prof::recs[i].bytes : prof::recs[i].work;
  }
  int64_t off = 0;
  host_buf[0] = 0;
  for (int c = 0; c < prof::C_COUNT; ++c) {
    if (!cnt[c]) continue;
    int n = snprintf(host_buf + off, (size_t)(buf_bytes - off), "%s %d %.6f %.6e %.6e\n", prof::names[c], cnt[c], ms[c], work[c], byt[c]);
    if (n < 0 || off + n >= buf_bytes) break;
    off += n;
  }
  prof::n_recs = 0;
  return 0;
}

// test hook: which = 0 forward (qkv -> ctx, lse), 1 backward (qkv, ctx, lse, dctx -> dqkv);
// variant 0 = tensor-core kernels, 1 = SIMT reference kernels; bf16 tensors, one backbone
int v2s_test_attention(int which, const void* qkv, void* ctx, float* lse, const void* dctx, void* dqkv, int batch,
                       int variant, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const void* cq[1] = {qkv}; void* cc[1] = {ctx}; float* ls[1] = {lse};
  const void* cctx[1] = {ctx}; const float* cls[1] = {lse}; const void* cd[1] = {dctx}; void* dq[1] = {dqkv};
  // variant: bit 0 = SIMT reference kernels, bit 1 = fp16 instead of bf16 tensors
  const int f16 = (variant & 2) ? 1 : 0, at = f16 ? AT_F16 : AT_BF16;
  if (which == 0) {
    if (variant & 1) return launch_attn_fwd_simt(cq, cc, ls, 1, batch, at, st);
    return launch_attn_fwd_tc(cq, cc, ls, 1, batch, st, f16);
  }
  if (which == 1) {
    if (variant & 1) return launch_attn_bwd_simt(cq, cctx, cls, cd, dq, 1, batch, at, st);
    return launch_attn_bwd_tc(cq, cctx, cls, cd, dq, 1, batch, st, f16);
  }
  set_error("v2s_test_attention: which must be 0 or 1");
  return 1;
}

int v2s_debug_flag(void) { return gemm_tc_error_flag(); }
int v2s_debug_counters(int64_t* host32) { return gemm_tc_debug_counters(reinterpret_cast<long long*>(host32)); }

int v2s_test_mlp(int mode, const void* a, const void* w1, const void* w2, const float* b1, const float* b2, void* u,
                 void* h, const float* resid, void* out, void* ln_out, const float* ln_gamma, const float* ln_beta,
                 float* ln_mean, float* ln_rstd, int m, int lp_f16, void* stream) {
  MlpDesc d = make_mlp_desc();
  d.mode = mode; d.M = m; d.groups = 1; d.lp_f16 = lp_f16;
  d.a[0] = a; d.w1[0] = w1; d.w2[0] = w2; d.b1[0] = b1; d.b2[0] = b2; d.u[0] = u; d.h[0] = h; d.resid[0] = resid;
  d.out[0] = out; d.ln_out[0] = ln_out; d.ln_gamma[0] = ln_gamma; d.ln_beta[0] = ln_beta; d.ln_mean[0] = ln_mean;
  d.ln_rstd[0] = ln_rstd;
  return launch_mlp_tc(d, (cudaStream_t)stream);
}

int v2s_test_gemm(int which, const void* a, const void* b, void* c, int m, int n, int k, int variant, void* stream) {
  return gemm_tc_test(which, a, b, c, m, n, k, variant, (cudaStream_t)stream);
}

}  // extern "C"
