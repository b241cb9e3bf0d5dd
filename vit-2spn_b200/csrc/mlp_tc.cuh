// Chained tcgen05 GEMM pair for the MLP half of a ViT block (HF:modeling_vit.py:296-312, 340-346): the 768-wide
// intermediate never round-trips through HBM between the two GEMMs.
#pragma once
#include "common.cuh"

namespace v2s {

enum MlpMode : int {
  MLP_FWD = 0,   // x_out = x_mid + b2 + gelu(xn2 W1^T + b1) W2^T   (+ optional LayerNorm of x_out for the next block)
  MLP_BWD = 1,   // du = (dx W2) * gelu'(u);  dxn2 = du W1
};

struct MlpDesc {
  int mode;
  int M;                       // token rows
  int groups;
  int lp_f16;                  // 16-bit format: 0 = bf16, 1 = fp16
  const void* a[MAXG];         // fwd: xn2 [M,192];  bwd: dx [M,192]            (16-bit)
  const void* w1[MAXG];        // intermediate.dense.weight [768,192]           (16-bit shadow)
  const void* w2[MAXG];        // output.dense.weight [192,768]                 (16-bit shadow)
  const float* b1[MAXG];       // fwd: [768]
  const float* b2[MAXG];       // fwd: [192]
  void* u[MAXG];               // fwd: optional out, pre-GELU activation [M,768]; bwd: in
  void* h[MAXG];               // fwd: optional out, gelu(u) [M,768];             bwd: out, du [M,768]
  const float* resid[MAXG];    // fwd: x_mid fp32 [M,192]
  void* out[MAXG];             // fwd: x_out fp32 [M,192];  bwd: dxn2 16-bit [M,192]
  void* ln_out[MAXG];          // fwd, optional: LayerNorm(x_out) * gamma + beta, 16-bit [M,192]
  const float* ln_gamma[MAXG];
  const float* ln_beta[MAXG];
  float* ln_mean[MAXG];        // optional [M]
  float* ln_rstd[MAXG];
  int late_wait;               // see GemmDesc::late_wait
};

inline MlpDesc make_mlp_desc() {
  MlpDesc d;
  memset(&d, 0, sizeof(d));
  d.groups = 1;
  return d;
}

int launch_mlp_tc(const MlpDesc& d, cudaStream_t stream);

}  // namespace v2s
