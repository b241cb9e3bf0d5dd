// Shared device helpers of the tensor-core kernels: the 16-bit operand formats (bf16 / fp16) and the fast erf-GELU
// used by the GELU-type epilogues.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "ptx.cuh"

namespace v2s {

// ---- 16-bit operand formats -------------------------------------------------------------------
// The tcgen05 kind::f16 MMA takes bf16 or fp16 operands (format field of the instruction descriptor); everything
// else that differs between the two "low-precision" compute modes is how a pair of floats is packed / unpacked.
//   bf16: torch.autocast(bfloat16) — no loss scaling needed
//   fp16: the reference's actual CUDA precision (torch.autocast default + GradScaler, ref:ssp_vit2spn_tiny.py:175,209-217)
struct LpBf16 {
  static constexpr uint32_t kIdescFmt = 1;      // a_format / b_format = BF16
  static constexpr bool kIsF16 = false;
  static constexpr uint32_t kOnePair = 0x3F803F80u;   // (1.0, 1.0)
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  static __device__ __forceinline__ float lo(uint32_t w) { return __uint_as_float(w << 16); }
  static __device__ __forceinline__ float hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
};
struct LpF16 {
  static constexpr uint32_t kIdescFmt = 0;      // a_format / b_format = F16
  static constexpr bool kIsF16 = true;
  static constexpr uint32_t kOnePair = 0x3C003C00u;   // (1.0, 1.0)
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  static __device__ __forceinline__ float lo(uint32_t w) { return __low2float(*reinterpret_cast<const __half2*>(&w)); }
  static __device__ __forceinline__ float hi(uint32_t w) { return __high2float(*reinterpret_cast<const __half2*>(&w)); }
};

// instruction descriptor for kind::f16 with a given operand format: 16-bit x 16-bit -> fp32, M x N tile
__host__ __device__ constexpr uint32_t make_idesc_lp(uint32_t fmt, int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- fast erf-GELU ------------------------------------------------------------------------------
// Abramowitz-Stegun 7.1.26, |erf error| < 1.5e-7 (far below 16-bit resolution); one MUFU.RCP + one MUFU.EX2 per
// element, the rest FMA-pipe work.  (A cheaper fitted logistic form, 7 FP + 2 MUFU and 5.7e-5 abs error, was
// measured: fc1 1.06 -> 1.00 ms but no change of the step, so the more accurate form stays.)
__device__ __forceinline__ float gelu_fast(float x) {
  // 0.5 x (1 + erf(x/sqrt2)) = x/2 + |x/2| erf(|x|/sqrt2): no sign transfer, constants folded (13 FP + 2 MUFU)
  const float t = ptx::rcp_approx(fmaf(0.3275911f * 0.70710678118654752f, fabsf(x), 1.0f));
  const float e = ptx::ex2_approx((-0.72134752044448170f * x) * x);   // exp(-x^2/2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = fmaf(-(poly * t), e, 1.0f);
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), erf_abs, hx);
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  // cdf(x) + x pdf(x), cdf = 1/2 + copysign(erf(|x|/sqrt2)/2, x); the 1/2 is folded into the polynomial
  const float t = ptx::rcp_approx(fmaf(0.3275911f * 0.70710678118654752f, fabsf(x), 1.0f));
  const float e = ptx::ex2_approx((-0.72134752044448170f * x) * x);   // exp(-x^2/2)
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  const float half_erf = fmaf(-(poly * t), e, 0.5f);
  const float cdf = 0.5f + copysignf(half_erf, x);
  return fmaf(x * 0.39894228040143268f, e, cdf);
}

}  // namespace v2s
