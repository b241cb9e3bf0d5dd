// GPU half of the reference's augmentation pipeline: everything from
// `transforms.Resize((224, 224))` on (ref:ssp_vit2spn_tiny.py:90-95).  The PIL-side, parameter-drawing
// augmentations stay on the host at 28x28 (784 bytes per view); this kernel turns each 28x28 uint8 view into the
// fp32 [3,224,224] network input:
//
//   Pillow BILINEAR resize of an 8-bit image (Resample.c: separable, horizontal pass first, 22-bit fixed-point
//   coefficients, rounding and clipping to uint8 after EACH pass - reproduced bit-exactly from the host-computed
//   coefficient tables)  ->  ToTensor (/255)  ->  GaussianBlur 3x3 (reflect padding, outer-product kernel)
//   ->  RandomErasing (rectangle := 0)  ->  Normalize, replicated to the 3 (identical) channels.
//
// Two CTAs per view (upper / lower half of the output rows): source and both resampling passes live in shared memory
// (44 KB), the 602 KB output is written once with 128-bit stores.  HBM-bound: 3*224*224*4 B written per view.
#include "common.cuh"
#include "kernels.cuh"

namespace v2s {

namespace {

constexpr int AUG_OUT = 224;
constexpr int AUG_MAX_IN = 64;
constexpr int AUG_PREC = 22;

struct AugP {
  const uint8_t* src;        // [n, in, in]
  const int32_t* bounds;     // [224, 2]  first source index, tap count
  const int32_t* coefs;      // [224, ksize]
  const float* k1d;          // [n, 3] Gaussian taps (NULL: no blur anywhere); a view with k1d[1] == 1 is not blurred
  const int32_t* erase;      // [n, 4] top, left, height, width (NULL or height <= 0: nothing erased)
  float* dst;                // [n, 3, 224, 224]
  float mean[3], inv_unused[3], std[3];
  int in_size, ksize;
};

__device__ __forceinline__ int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

constexpr int AUG_HALF = AUG_OUT / 2;          // output rows per CTA
constexpr int AUG_ROWS = AUG_HALF + 2;         // resampled rows a CTA keeps: its own plus one halo row per side

__global__ void __launch_bounds__(256) augment_finish_kernel(const __grid_constant__ AugP p) {
  extern __shared__ uint8_t sm[];
  const int in = p.in_size;
  uint8_t* s_src = sm;                                  // [in][in]
  uint8_t* s_h = s_src + AUG_MAX_IN * AUG_MAX_IN;       // [in][224]          after the horizontal pass
  uint8_t* s_r = s_h + AUG_MAX_IN * AUG_OUT;            // [AUG_ROWS][224]    after the vertical pass (rows r0..)
  __shared__ int s_b[AUG_OUT][2];
  __shared__ int s_c[AUG_OUT][4];
  __shared__ float s_lut[256];                          // ToTensor: v / 255 in fp32
  const int n = blockIdx.x, half = blockIdx.y, tid = threadIdx.x;
  const int y_begin = half * AUG_HALF;
  const int r0 = y_begin > 0 ? y_begin - 1 : 0;
  const int r1 = min(y_begin + AUG_HALF, AUG_OUT - 1);  // last resampled row needed (inclusive)
  for (int i = tid; i < in * in; i += 256) s_src[i] = p.src[(int64_t)n * in * in + i];
  for (int i = tid; i < AUG_OUT; i += 256) {
    s_b[i][0] = p.bounds[2 * i]; s_b[i][1] = p.bounds[2 * i + 1];
    for (int k = 0; k < 4; ++k) s_c[i][k] = k < p.ksize ? p.coefs[i * p.ksize + k] : 0;
  }
  s_lut[tid] = __fdiv_rn((float)tid, 255.0f);
  __syncthreads();
  for (int i = tid; i < in * AUG_OUT; i += 256) {       // horizontal pass (all source rows: the tables decide which are used)
    const int y = i / AUG_OUT, xx = i - y * AUG_OUT;
    int acc = 1 << (AUG_PREC - 1);
    const int x0 = s_b[xx][0], cnt = s_b[xx][1];
    for (int k = 0; k < cnt; ++k) acc += (int)s_src[y * in + x0 + k] * s_c[xx][k];
    s_h[i] = (uint8_t)clip8(acc >> AUG_PREC);
  }
  __syncthreads();
  for (int i = tid; i < (r1 - r0 + 1) * AUG_OUT; i += 256) {  // vertical pass, rows r0..r1
    const int ry = i / AUG_OUT, xx = i - ry * AUG_OUT, yy = r0 + ry;
    int acc = 1 << (AUG_PREC - 1);
    const int y0 = s_b[yy][0], cnt = s_b[yy][1];
    for (int k = 0; k < cnt; ++k) acc += (int)s_h[(y0 + k) * AUG_OUT + xx] * s_c[yy][k];
    s_r[i] = (uint8_t)clip8(acc >> AUG_PREC);
  }
  __syncthreads();
  float k0 = 0.f, k1 = 1.f, k2 = 0.f;
  bool blur = false;
  if (p.k1d != nullptr) {
    k0 = p.k1d[3 * n]; k1 = p.k1d[3 * n + 1]; k2 = p.k1d[3 * n + 2];
    blur = !(k1 == 1.0f && k0 == 0.0f && k2 == 0.0f);
  }
  int et = 0, el = 0, eh = 0, ew = 0;
  if (p.erase != nullptr) { et = p.erase[4 * n]; el = p.erase[4 * n + 1]; eh = p.erase[4 * n + 2]; ew = p.erase[4 * n + 3]; }
  const float kk[3] = {k0, k1, k2};
  float w2[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) w2[a][b] = __fmul_rn(kk[a], kk[b]);     // torch.mm(k[:, None], k[None, :])
  float* out = p.dst + (int64_t)n * 3 * AUG_OUT * AUG_OUT;
  for (int i = tid; i < AUG_HALF * AUG_OUT / 4; i += 256) {     // 4 pixels of one row per thread
    const int yl = (i * 4) / AUG_OUT, x = (i * 4) - yl * AUG_OUT, y = y_begin + yl;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int xx = x + e;
      float val;
      if (blur) {
        val = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          int ya = y + a - 1;
          ya = ya < 0 ? -ya : (ya >= AUG_OUT ? 2 * AUG_OUT - 2 - ya : ya);         // reflect (no edge repeat)
          const uint8_t* rrow = s_r + (ya - r0) * AUG_OUT;
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            int xb = xx + b - 1;
            xb = xb < 0 ? -xb : (xb >= AUG_OUT ? 2 * AUG_OUT - 2 - xb : xb);
            val = fmaf(w2[a][b], s_lut[rrow[xb]], val);
          }
        }
      } else {
        val = s_lut[s_r[(y - r0) * AUG_OUT + xx]];
      }
      if (eh > 0 && y >= et && y < et + eh && xx >= el && xx < el + ew) val = 0.f;
      v[e] = val;
    }
    const int o4 = (y * AUG_OUT + x) >> 2;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float4 o;
      o.x = __fdiv_rn(v[0] - p.mean[c], p.std[c]); o.y = __fdiv_rn(v[1] - p.mean[c], p.std[c]);
      o.z = __fdiv_rn(v[2] - p.mean[c], p.std[c]); o.w = __fdiv_rn(v[3] - p.mean[c], p.std[c]);
      reinterpret_cast<float4*>(out + (int64_t)c * AUG_OUT * AUG_OUT)[o4] = o;
    }
  }
}

}  // namespace

int launch_augment_finish(const uint8_t* src, int n, int in_size, const int32_t* bounds, const int32_t* coefs, int ksize,
                          const float* k1d, const int32_t* erase, const float* mean3, const float* std3, float* dst,
                          cudaStream_t s) {
  if (in_size < 2 || in_size > AUG_MAX_IN) { set_error("augment: source size %d outside [2, %d]", in_size, AUG_MAX_IN); return 1; }
  if (ksize < 1 || ksize > 4) { set_error("augment: %d filter taps (expected <= 4: bilinear up-sampling)", ksize); return 1; }
  if (reinterpret_cast<uintptr_t>(dst) & 15) { set_error("augment: dst must be 16-byte aligned"); return 1; }
  AugP p;
  memset(&p, 0, sizeof(p));
  p.src = src; p.bounds = bounds; p.coefs = coefs; p.k1d = k1d; p.erase = erase; p.dst = dst;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean3[c]; p.std[c] = std3[c]; }
  p.in_size = in_size; p.ksize = ksize;
  const size_t smem = AUG_MAX_IN * AUG_MAX_IN + AUG_MAX_IN * AUG_OUT + AUG_ROWS * AUG_OUT;
  static bool attr[MAX_DEVICES] = {false};
  if (!attr[cur_device()]) {
    V2S_CUDA_OK(cudaFuncSetAttribute(augment_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[cur_device()] = true;
  }
  augment_finish_kernel<<<dim3(n, 2), 256, smem, s>>>(p);
  V2S_LAUNCH_CHECK();
  return 0;
}

}  // namespace v2s
