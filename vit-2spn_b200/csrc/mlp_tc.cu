// Chained tcgen05 GEMM pair for the MLP half of a ViT block: the 768-wide intermediate stays on chip.
//
//   forward  (HF:modeling_vit.py:296-312,340-346)   u = xn2 W1^T + b1;  h = gelu(u);  x_out = x_mid + b2 + h W2^T
//                                                   (+ LayerNorm of x_out = LN1 of the next block, fused)
//   backward (dgrad chain of the same two layers)   du = (dx W2) * gelu'(u);  dxn2 = du W1
//
// Persistent CTA (640 threads, one per SM) over the (backbone, 128-row m-tile) jobs.  The intermediate dimension is
// walked in six chunks of 128 columns:
//   stage 1   acc1[c&1] (TMEM, 128 cols, double-buffered) = A1[128x192] * B1_chunk          12 UMMAs 128x128x16
//   epilogue  all 16 epilogue warps take the chunk together (thread = one row x 32 columns): bias+GELU or gelu'(u)
//             product -> 16-bit -> written as the K-major SWIZZLE_128B A operand of stage 2 into two [128 x 64]
//             k-block tiles in shared memory
//   stage 2   acc2 (TMEM, 192 cols) += A2_chunk[128x128] * B2_chunk                          8 UMMAs 128x192x16
//   after six chunks the epilogue warps drain acc2: forward = + bias + fp32 residual -> x_out, row statistics,
//   normalised 16-bit row; backward = 16-bit dxn2.
// The MMA warp issues S1(c+1) before S2(c), so the next accumulator is ready when the epilogue warps finish a chunk.
//
// Global traffic of the epilogues goes through TMA only (a thread-per-row access pattern costs 32 LSU wavefronts
// per instruction: measured 2-3x the whole kernel).  Four 16 KB "epilogue buffers" EB0..3:
//   EB0/EB1  the A2 k-block tiles; the same tiles are the source of the TMA stores of h (forward, online
//            backbones only) / du (backward): no separate staging
//   EB2/EB3  forward: staging of the pre-GELU activation u (stored for the backward pass of the online backbones);
//            backward: the u tiles, TMA-loaded one chunk ahead
//   while acc2 is drained no chunk epilogue runs, so all four buffers stage the residual (TMA load, updated in
//   place, TMA store) and the normalised rows.
// A store warp owns every TMA store / epilogue load and the buffer recycling; weights stream from L2 through a
// 4-slot TMA ring (W1 / W2 chunks: 576 KB per m-tile); A1 (48 KB) is resident per m-tile.
#include <string.h>

#include "gemm_tc.cuh"
#include "mlp_tc.cuh"
#include "ptx.cuh"
#include "tc_math.cuh"

namespace v2s {

namespace {

constexpr int BM = 128;                    // rows per m-tile
constexpr int CW = 128;                    // chunk width along the 768-wide intermediate dimension
constexpr int NCH = DF / CW;               // 6 chunks
constexpr int KB1 = D / 64;                // 3 k-blocks in stage 1 (K = 192)
constexpr int KB2 = CW / 64;               // 2 k-blocks per chunk in stage 2
constexpr int KBLK = BM * 128;             // one [128 rows x 128 B] tile: 16384 B
constexpr int A1_BYTES = KB1 * KBLK;       // 49152
constexpr int W_SLOT = 24576;              // holds a stage-1 B k-block (16 KB) or a stage-2 B k-block (24 KB)
constexpr int W_SLOTS = 3;
constexpr int OFF_W = A1_BYTES;
constexpr int N_EB = 6;
constexpr int OFF_EB = OFF_W + W_SLOTS * W_SLOT;       // 122880: EB0..5
constexpr int OFF_BAR = OFF_EB + N_EB * KBLK;          // 221184
constexpr int OFF_LN = OFF_BAR + 1024;                 // [4 column quarters][128 rows] float2
constexpr int SMEM_BYTES = OFF_LN + 4 * BM * 8 + 1024;
constexpr int TM_ACC2 = 256;               // TMEM columns: acc1[0] 0..127, acc1[1] 128..255, acc2 256..447
constexpr int N_THREADS = 640;             // producer, MMA, store warp, spare warp, 16 epilogue warps
constexpr int EPI_THREADS = 512;
constexpr int XCH = D / 32;                // 6 column chunks of 32 when acc2 is drained
static_assert(SMEM_BYTES <= 232448, "shared-memory budget");
static_assert(OFF_W % 1024 == 0 && OFF_EB % 1024 == 0 && W_SLOT % 1024 == 0, "swizzled tiles need 1024-byte alignment");

struct alignas(64) MlpParams {
  CUtensorMap tmA[MAXG], tmB1[MAXG], tmB2[MAXG];
  CUtensorMap tmH[MAXG];      // [M,768] 16-bit, box 64 x 128: h (forward, optional) / du (backward)
  CUtensorMap tmU[MAXG];      // [M,768] 16-bit, box 64 x 128: u out (forward, optional) / u in (backward)
  CUtensorMap tmRes[MAXG];    // forward: x_mid fp32 [M,192], box 32 x 128
  CUtensorMap tmOut[MAXG];    // forward: x_out fp32 [M,192], box 32 x 128;  backward: dxn2 16-bit [M,192], box 64 x 128
  CUtensorMap tmXn[MAXG];     // forward, optional: normalised row 16-bit [M,192], box 64 x 128
  const float* b1[MAXG]; const float* b2[MAXG];
  const float* ln_gamma[MAXG]; const float* ln_beta[MAXG]; float* ln_mean[MAXG]; float* ln_rstd[MAXG];
  int has_h[MAXG], has_u[MAXG], has_ln[MAXG];
  int M, tiles_m, total_tiles, late_wait;
  int* err_flag;
  long long* dbg;   // optional cycle counters of CTA 0 (V2S_GEMM_DEBUG): [0..7] MMA warp waits, [8..19] epilogue thread waits
};

template <int MODE, typename LP, bool DBG>
__global__ void __launch_bounds__(N_THREADS, 1) mlp_tc_kernel(const __grid_constant__ MlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* a1_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* a1_empty = a1_full + 1;
  uint64_t* w_full = a1_empty + 1;          // [4]
  uint64_t* w_empty = w_full + W_SLOTS;     // [4]
  uint64_t* acc1_full = w_empty + W_SLOTS;  // [2]
  uint64_t* acc1_empty = acc1_full + 2;     // [2] one arrival per epilogue warp
  uint64_t* a2_full = acc1_empty + 2;       // [2] 256 arrivals: the threads that write the k-block tile
  uint64_t* a2_free = a2_full + 2;          // [2] 2 arrivals: stage-2 UMMAs retired + store warp (h / du store has read it)
  uint64_t* u_full = a2_free + 2;           // [2] forward: 256 arrivals (u staged); backward: TMA load of the u tile landed
  uint64_t* u_free = u_full + 2;            // [2] forward: store warp (u store has read it); backward: 256 readers done
  uint64_t* acc2_full = u_free + 2;
  uint64_t* acc2_empty = acc2_full + 1;     // one arrival per epilogue warp
  uint64_t* rs_full = acc2_empty + 1;       // [6] residual chunk k landed in its buffer
  uint64_t* st_full = rs_full + XCH;        // [6] 512 arrivals: output chunk k finished in its buffer
  uint64_t* xn_free = st_full + XCH;        // store warp: the three buffers of the normalised row / dxn2 tiles may be overwritten
  uint64_t* xn_full = xn_free + 1;          // 512 arrivals: those three tiles are written
  uint64_t* e2_done = xn_full + 1;          // store warp: every store of the tile's acc2 drain has read its buffer
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(e2_done + 1);
  float2* ln_part = reinterpret_cast<float2*>(smem + OFF_LN);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int g = 0; g * p.tiles_m < p.total_tiles; ++g) {
      ptx::prefetch_tmap(&p.tmA[g]); ptx::prefetch_tmap(&p.tmB1[g]); ptx::prefetch_tmap(&p.tmB2[g]);
    }
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(a1_full, 1); ptx::mbar_init(a1_empty, 1);
    for (int s = 0; s < W_SLOTS; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&acc1_full[s], 1); ptx::mbar_init(&acc1_empty[s], 16);
      ptx::mbar_init(&a2_full[s], 256); ptx::mbar_init(&a2_free[s], 2);
      ptx::mbar_init(&u_full[s], MODE == MLP_FWD ? 256 : 1); ptx::mbar_init(&u_free[s], MODE == MLP_FWD ? 1 : 256);
    }
    ptx::mbar_init(acc2_full, 1); ptx::mbar_init(acc2_empty, 16);
    for (int s = 0; s < XCH; ++s) { ptx::mbar_init(&rs_full[s], 1); ptx::mbar_init(&st_full[s], EPI_THREADS); }
    ptx::mbar_init(xn_free, 1); ptx::mbar_init(xn_full, EPI_THREADS); ptx::mbar_init(e2_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  if (!p.late_wait) {      // programmatic dependent launch: see the note in gemm_tc.cu
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
  }

  const int n_my = ((int)blockIdx.x < p.total_tiles) ? (p.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto tile_at = [&](int i, int& g, int& mt) {
    const int t = blockIdx.x + i * gridDim.x;
    g = t / p.tiles_m;
    mt = t - g * p.tiles_m;
  };
  uint8_t* const eb = smem + OFF_EB;
  long long tk[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_begin = DBG ? clock64() : 0;
#define V2S_WAIT(slot, bar, par, code)                                   \
  do {                                                                  \
    const long long _t0 = DBG ? clock64() : 0;                          \
    ptx::mbar_wait(bar, par, p.err_flag, code);                         \
    if (DBG) tk[slot] += clock64() - _t0;                               \
  } while (0)

  if (warp == 0) {
    // ================= TMA producer: A1 per tile, weight k-blocks in the order the MMA warp consumes them =========
    int ws = 0; uint32_t wph = 0;
    auto advance = [&]() { if (++ws == W_SLOTS) { ws = 0; wph ^= 1; } };
    auto load_b1 = [&](int g, int j) {
#pragma unroll 1
      for (int kb = 0; kb < KB1; ++kb) {
        ptx::mbar_wait(&w_empty[ws], wph ^ 1, p.err_flag, 41);
        if (ptx::elect_one()) {
          uint8_t* slot = smem + OFF_W + ws * W_SLOT;
          ptx::mbar_arrive_expect_tx(&w_full[ws], CW * 128);
          if (MODE == MLP_FWD) {      // W1 rows [j*128, +128), k-block kb: K-major
            ptx::tma_load_2d(slot, &p.tmB1[g], &w_full[ws], kb * 64, j * CW);
          } else {                    // W2 k-rows [kb*64, +64), columns [j*128, +128): MN-major, two 64-column boxes
            ptx::tma_load_2d(slot, &p.tmB1[g], &w_full[ws], j * CW, kb * 64);
            ptx::tma_load_2d(slot + 8192, &p.tmB1[g], &w_full[ws], j * CW + 64, kb * 64);
          }
        }
        __syncwarp();
        advance();
      }
    };
    auto load_b2 = [&](int g, int j) {
#pragma unroll 1
      for (int kb = 0; kb < KB2; ++kb) {
        ptx::mbar_wait(&w_empty[ws], wph ^ 1, p.err_flag, 42);
        if (ptx::elect_one()) {
          uint8_t* slot = smem + OFF_W + ws * W_SLOT;
          ptx::mbar_arrive_expect_tx(&w_full[ws], D * 128);
          if (MODE == MLP_FWD) {      // W2 all 192 rows, k columns [j*128 + kb*64, +64): K-major
            ptx::tma_load_2d(slot, &p.tmB2[g], &w_full[ws], j * CW + kb * 64, 0);
          } else {                    // W1 k-rows [j*128 + kb*64, +64), all 192 columns: MN-major, three boxes
            ptx::tma_load_2d(slot, &p.tmB2[g], &w_full[ws], 0, j * CW + kb * 64);
            ptx::tma_load_2d(slot + 8192, &p.tmB2[g], &w_full[ws], 64, j * CW + kb * 64);
            ptx::tma_load_2d(slot + 16384, &p.tmB2[g], &w_full[ws], 128, j * CW + kb * 64);
          }
        }
        __syncwarp();
        advance();
      }
    };
    // stage-2 operands lag two chunks behind stage-1 operands (the MMA warp's order): (g, j) of the last two chunks
    int pg1 = -1, pj1 = 0, pg2 = -1, pj2 = 0;
    for (int i = 0; i < n_my; ++i) {
      int g, mt;
      tile_at(i, g, mt);
      ptx::mbar_wait(a1_empty, (i & 1) ^ 1, p.err_flag, 43);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(a1_full, A1_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB1; ++kb) ptx::tma_load_2d(smem + kb * KBLK, &p.tmA[g], a1_full, kb * 64, mt * BM);
      }
      __syncwarp();
#pragma unroll 1
      for (int j = 0; j < NCH; ++j) {
        load_b1(g, j);
        if (pg2 >= 0) load_b2(pg2, pj2);
        pg2 = pg1; pj2 = pj1; pg1 = g; pj1 = j;
      }
    }
    if (pg2 >= 0) load_b2(pg2, pj2);
    if (pg1 >= 0) load_b2(pg1, pj1);
  } else if (warp == 1) {
    // ================= MMA issuer: S1(c), then S2(c-2) =================
    // (S1(c+1) only needs the accumulator drained by the chunk epilogue c-1, which happens at its very start; issuing
    //  it ahead of S2(c-1), which needs that epilogue's END, has the next accumulator ready a whole chunk early)
    constexpr uint32_t b_mn = (MODE == MLP_BWD) ? 1u : 0u;
    const uint32_t idesc1 = make_idesc_lp(LP::kIdescFmt, BM, CW, 0, b_mn);
    const uint32_t idesc2 = make_idesc_lp(LP::kIdescFmt, BM, D, 0, b_mn);
    const uint32_t sbase = ptx::smem_u32(smem);
    const uint32_t b_lbo = b_mn ? 8192u : 16u, b_step = b_mn ? (2048u >> 4) : (32u >> 4);
    int ws = 0; uint32_t wph = 0;
    auto advance = [&]() { if (++ws == W_SLOTS) { ws = 0; wph ^= 1; } };
    auto stage1 = [&](int i, int j, int c) {
      const int b = c & 1;
      V2S_WAIT(0, &acc1_empty[b], ((c >> 1) & 1) ^ 1, 44);
      if (j == 0) V2S_WAIT(1, a1_full, i & 1, 45);
      const uint32_t d_tmem = tmem_base + b * CW;
#pragma unroll 1
      for (int kb = 0; kb < KB1; ++kb) {
        V2S_WAIT(2, &w_full[ws], wph, 46);
        ptx::tc_fence_after();
        const uint32_t a_lo = ptx::desc_lo(sbase + kb * KBLK, 16);
        const uint32_t b_lo = ptx::desc_lo(sbase + OFF_W + ws * W_SLOT, b_lbo);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_lohi(d_tmem, a_lo + k * 2, b_lo + k * b_step, ptx::DESC_HI_SW128_SBO1024, idesc1,
                                (kb > 0 || k > 0) ? 1u : 0u);
          ptx::umma_commit(&w_empty[ws]);
          if (kb == KB1 - 1) {
            ptx::umma_commit(&acc1_full[b]);
            if (j == NCH - 1) ptx::umma_commit(a1_empty);      // the tile's A1 is dead once these retire
          }
        }
        __syncwarp();
        advance();
      }
    };
    auto stage2 = [&](int i, int j, int c) {
      if (j == 0) V2S_WAIT(3, acc2_empty, (i & 1) ^ 1, 47);
      const uint32_t d_tmem = tmem_base + TM_ACC2;
#pragma unroll 1
      for (int kb = 0; kb < KB2; ++kb) {
        V2S_WAIT(4, &a2_full[kb], c & 1, 48);
        V2S_WAIT(5, &w_full[ws], wph, 49);
        ptx::tc_fence_after();
        const uint32_t a_lo = ptx::desc_lo(sbase + OFF_EB + kb * KBLK, 16);
        const uint32_t b_lo = ptx::desc_lo(sbase + OFF_W + ws * W_SLOT, b_lbo);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_lohi(d_tmem, a_lo + k * 2, b_lo + k * b_step, ptx::DESC_HI_SW128_SBO1024, idesc2,
                                (j > 0 || kb > 0 || k > 0) ? 1u : 0u);
          ptx::umma_commit(&w_empty[ws]);
          ptx::umma_commit(&a2_free[kb]);
          if (j == NCH - 1 && kb == KB2 - 1) ptx::umma_commit(acc2_full);
        }
        __syncwarp();
        advance();
      }
    };
    int c = 0, pi1 = -1, pj1 = 0, pi2 = -1, pj2 = 0;
    for (int i = 0; i < n_my; ++i) {
#pragma unroll 1
      for (int j = 0; j < NCH; ++j, ++c) {
        stage1(i, j, c);
        if (pi2 >= 0) stage2(pi2, pj2, c - 2);
        pi2 = pi1; pj2 = pj1; pi1 = i; pj1 = j;
      }
    }
    if (pi2 >= 0) stage2(pi2, pj2, c - 2);
    if (pi1 >= 0) stage2(pi1, pj1, c - 1);
    if (DBG && blockIdx.x == 0 && lane == 0) {
      for (int k = 0; k < 6; ++k) p.dbg[k] = tk[k];
      p.dbg[6] = clock64() - t_begin; p.dbg[7] = n_my;
    }
  } else if (warp == 2) {
    // ================= store warp: every TMA store, the epilogue-side TMA loads, buffer recycling ==============
    // (lane 0 issues: bulk async-groups are per thread)
    int c = 0, nu = 0, nl = 0;
    if (MODE == MLP_BWD && n_my > 0 && lane == 0) {     // u tiles of the very first chunk
      int g, mt;
      tile_at(0, g, mt);
      for (int kb = 0; kb < 2; ++kb) {
        ptx::mbar_arrive_expect_tx(&u_full[kb], KBLK);
        ptx::tma_load_2d(eb + (2 + kb) * KBLK, &p.tmU[g], &u_full[kb], kb * 64, mt * BM);
      }
    }
    __syncwarp();
    for (int i = 0; i < n_my; ++i) {
      int g, mt;
      tile_at(i, g, mt);
      const int m0 = mt * BM;
      const bool has_h = p.has_h[g] != 0, has_u = p.has_u[g] != 0, has_ln = p.has_ln[g] != 0;
      if (MODE == MLP_FWD && lane == 0) {
        // residual chunks 0, 1 of this tile into their dedicated buffers EB4 / EB5, long before the drain needs them
        for (int k = 0; k < 2; ++k) {
          ptx::mbar_arrive_expect_tx(&rs_full[k], KBLK);
          ptx::tma_load_2d(eb + (4 + k) * KBLK, &p.tmRes[g], &rs_full[k], k * 32, m0);
        }
      }
      __syncwarp();
#pragma unroll 1
      for (int j = 0; j < NCH; ++j, ++c) {
        if (MODE == MLP_BWD) {
          // u tiles of the next chunk, as soon as this chunk's readers have them in registers
          int g2 = g, mt2 = mt, j2 = j + 1;
          bool have_next = true;
          if (j2 == NCH) { j2 = 0; have_next = i + 1 < n_my; if (have_next) tile_at(i + 1, g2, mt2); }
          if (have_next) {
            for (int kb = 0; kb < 2; ++kb) {
              ptx::mbar_wait(&u_free[kb], c & 1, p.err_flag, 53);
              if (lane == 0) {
                ptx::mbar_arrive_expect_tx(&u_full[kb], KBLK);
                ptx::tma_load_2d(eb + (2 + kb) * KBLK, &p.tmU[g2], &u_full[kb], j2 * CW + kb * 64, mt2 * BM);
              }
              __syncwarp();
            }
          }
        }
        ptx::mbar_wait(&a2_full[0], c & 1, p.err_flag, 54);
        ptx::mbar_wait(&a2_full[1], c & 1, p.err_flag, 54);
        if (has_h && lane == 0) {
          ptx::tma_store_2d(&p.tmH[g], eb, j * CW, m0);
          ptx::tma_store_2d(&p.tmH[g], eb + KBLK, j * CW + 64, m0);
          ptx::tma_commit_group();
        }
        __syncwarp();
        if (MODE == MLP_FWD && has_u) {
          ptx::mbar_wait(&u_full[0], nu & 1, p.err_flag, 55);
          ptx::mbar_wait(&u_full[1], nu & 1, p.err_flag, 55);
          if (lane == 0) {
            ptx::tma_store_2d(&p.tmU[g], eb + 2 * KBLK, j * CW, m0);
            ptx::tma_store_2d(&p.tmU[g], eb + 3 * KBLK, j * CW + 64, m0);
            ptx::tma_commit_group();
          }
          __syncwarp();
          ++nu;
        }
        if (lane == 0) {
          if (has_h || (MODE == MLP_FWD && has_u)) ptx::tma_wait_group_read<0>();
          ptx::mbar_arrive(&a2_free[0]); ptx::mbar_arrive(&a2_free[1]);
          if (MODE == MLP_FWD && has_u) { ptx::mbar_arrive(&u_free[0]); ptx::mbar_arrive(&u_free[1]); }
        }
        __syncwarp();
      }
      // ---- acc2 drain ----
      ptx::mbar_wait(acc2_full, i & 1, p.err_flag, 56);      // all stage-2 UMMAs of the tile retired: EB0..3 are idle
      if (MODE == MLP_FWD) {
        // chunk k lives in buffer EB[k < 2 ? 4 + k : k - 2]: six buffers for six chunks, no recycling inside a tile
        if (lane == 0) {
          for (int k = 2; k < XCH; ++k) {
            ptx::mbar_arrive_expect_tx(&rs_full[k], KBLK);
            ptx::tma_load_2d(eb + (k - 2) * KBLK, &p.tmRes[g], &rs_full[k], k * 32, m0);
          }
        }
        __syncwarp();
#pragma unroll 1
        for (int k = 0; k < XCH; ++k) {
          ptx::mbar_wait(&st_full[k], i & 1, p.err_flag, 57);
          if (lane == 0) {
            ptx::tma_store_2d(&p.tmOut[g], eb + (k < 2 ? 4 + k : k - 2) * KBLK, k * 32, m0);
            ptx::tma_commit_group();
          }
          __syncwarp();
        }
        if (has_ln) {
          if (lane == 0) {
            ptx::tma_wait_group_read<3>();        // the stores of chunks 0..2 have read EB4, EB5, EB0
            ptx::mbar_arrive(xn_free);
          }
          __syncwarp();
          ptx::mbar_wait(xn_full, nl & 1, p.err_flag, 58);
          ++nl;
          if (lane == 0) {
            ptx::tma_store_2d(&p.tmXn[g], eb + 4 * KBLK, 0, m0);
            ptx::tma_store_2d(&p.tmXn[g], eb + 5 * KBLK, 64, m0);
            ptx::tma_store_2d(&p.tmXn[g], eb, 128, m0);
            ptx::tma_commit_group();
          }
          __syncwarp();
        }
      } else {
        ptx::mbar_wait(xn_full, i & 1, p.err_flag, 59);
        if (lane == 0) {
          ptx::tma_store_2d(&p.tmOut[g], eb + 4 * KBLK, 0, m0);
          ptx::tma_store_2d(&p.tmOut[g], eb + 5 * KBLK, 64, m0);
          ptx::tma_store_2d(&p.tmOut[g], eb, 128, m0);
          ptx::tma_commit_group();
        }
        __syncwarp();
      }
      if (lane == 0) {
        ptx::tma_wait_group_read<0>();
        ptx::mbar_arrive(e2_done);
      }
      __syncwarp();
    }
    if (lane == 0) ptx::tma_wait_group<0>();
    __syncwarp();
  } else if (warp >= 4) {
    // ================= epilogue warps =================
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int cq = (warp - 4) >> 2;              // column quarter
    const int kbq = cq >> 1;                     // k-block tile of the chunk this thread writes
    const int row = q * 32 + lane;
    const int rx = row & 7;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t eb_row = ptx::smem_u32(eb) + row * 128;      // this thread's row in EB0
    int c = 0, nu = 0, nl = 0;
    for (int i = 0; i < n_my; ++i) {
      int g, mt;
      tile_at(i, g, mt);
      const int64_t grow = (int64_t)mt * BM + row;
      const bool has_u = p.has_u[g] != 0, has_ln = p.has_ln[g] != 0;
#pragma unroll 1
      for (int j = 0; j < NCH; ++j, ++c) {
        const int b = c & 1;
        const int col0 = j * CW + cq * 32;                      // this thread's 32 columns of the 768-wide dimension
        uint4 ux[4];
        if (MODE == MLP_BWD) {                                  // pre-GELU activation of the forward pass
          V2S_WAIT(1, &u_full[kbq], c & 1, 61);
#pragma unroll
          for (int t = 0; t < 4; ++t) ux[t] = ptx::lds128(eb_row + (2 + kbq) * KBLK + ((((cq & 1) * 4 + t) ^ rx) << 4));
          // The tile is handed back to the TMA engine (next chunk's u): the loads above must have READ shared memory
          // before the arrival is visible.  Nothing below depends on their data yet, and the hardware lets the
          // barrier arrival overtake loads that are merely issued (observed: a few rows picked up the next chunk's u),
          // so wait for them explicitly.
          __threadfence_block();
          ptx::mbar_arrive(&u_free[kbq]);
        }
        V2S_WAIT(2, &acc1_full[b], (c >> 1) & 1, 50);
        ptx::tc_fence_after();
        uint32_t r[32];
        ptx::tmem_ld_32x32(tl + b * CW + cq * 32, r);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        if (lane == 0) ptx::mbar_arrive(&acc1_empty[b]);
        uint32_t o[16], pu[16];
        if (MODE == MLP_FWD) {
          const float4* b4 = reinterpret_cast<const float4*>(p.b1[g] + col0);
          float v[32];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float4 bb = __ldg(b4 + t);
            v[4 * t] = __uint_as_float(r[4 * t]) + bb.x; v[4 * t + 1] = __uint_as_float(r[4 * t + 1]) + bb.y;
            v[4 * t + 2] = __uint_as_float(r[4 * t + 2]) + bb.z; v[4 * t + 3] = __uint_as_float(r[4 * t + 3]) + bb.w;
          }
          if (has_u) {
#pragma unroll
            for (int k = 0; k < 16; ++k) pu[k] = LP::pack(v[2 * k], v[2 * k + 1]);
          }
#pragma unroll
          for (int k = 0; k < 16; ++k) o[k] = LP::pack(gelu_fast(v[2 * k]), gelu_fast(v[2 * k + 1]));
        } else {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint32_t w[4] = {ux[t].x, ux[t].y, ux[t].z, ux[t].w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              o[4 * t + e] = LP::pack(__uint_as_float(r[8 * t + 2 * e]) * gelu_grad_fast(LP::lo(w[e])),
                                      __uint_as_float(r[8 * t + 2 * e + 1]) * gelu_grad_fast(LP::hi(w[e])));
          }
        }
        // A operand of stage 2: k-block tile kbq of the chunk, 16-byte pieces (cq & 1) * 4 + t of this row
        if (j == 0 && i > 0) V2S_WAIT(0, e2_done, (i - 1) & 1, 60);   // the previous tile's drain has left the buffers
        V2S_WAIT(3, &a2_free[kbq], (c & 1) ^ 1, 51);
#pragma unroll
        for (int t = 0; t < 4; ++t)
          ptx::sts128(eb_row + kbq * KBLK + ((((cq & 1) * 4 + t) ^ rx) << 4), o[4 * t], o[4 * t + 1], o[4 * t + 2], o[4 * t + 3]);
        ptx::fence_proxy_async();
        ptx::mbar_arrive(&a2_full[kbq]);
        if (MODE == MLP_FWD && has_u) {
          V2S_WAIT(4, &u_free[kbq], (nu & 1) ^ 1, 62);
#pragma unroll
          for (int t = 0; t < 4; ++t)
            ptx::sts128(eb_row + (2 + kbq) * KBLK + ((((cq & 1) * 4 + t) ^ rx) << 4), pu[4 * t], pu[4 * t + 1], pu[4 * t + 2],
                        pu[4 * t + 3]);
          ptx::fence_proxy_async();
          ptx::mbar_arrive(&u_full[kbq]);
          ++nu;
        }
      }
      // ---- drain acc2: in every 32-column chunk k this thread owns columns 32 k + 8 cq .. + 7 of its row ----
      const long long t_e2 = DBG ? clock64() : 0;
      V2S_WAIT(5, acc2_full, i & 1, 52);
      ptx::tc_fence_after();
      uint32_t r[48];
#pragma unroll
      for (int k = 0; k < XCH; ++k) ptx::tmem_ld_32x8(tl + TM_ACC2 + k * 32 + cq * 8, r + 8 * k);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      if (lane == 0) ptx::mbar_arrive(acc2_empty);
      long long t_ph = DBG ? clock64() : 0;
      if (DBG) tk[9] += t_ph - t_e2;
      if (MODE == MLP_FWD) {
        float v[48];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < XCH; ++k) {
          const int bq = k < 2 ? 4 + k : k - 2;
          const float4 ba = __ldg(reinterpret_cast<const float4*>(p.b2[g] + k * 32 + cq * 8));
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b2[g] + k * 32 + cq * 8 + 4));
          V2S_WAIT(6, &rs_full[k], i & 1, 63);
          const uint32_t s0 = eb_row + bq * KBLK + (((2 * cq) ^ rx) << 4), s1a = eb_row + bq * KBLK + (((2 * cq + 1) ^ rx) << 4);
          const float4 xa = ptx::lds128f(s0), xb = ptx::lds128f(s1a);
          float* x = v + 8 * k;
          x[0] = __uint_as_float(r[8 * k]) + ba.x + xa.x; x[1] = __uint_as_float(r[8 * k + 1]) + ba.y + xa.y;
          x[2] = __uint_as_float(r[8 * k + 2]) + ba.z + xa.z; x[3] = __uint_as_float(r[8 * k + 3]) + ba.w + xa.w;
          x[4] = __uint_as_float(r[8 * k + 4]) + bb.x + xb.x; x[5] = __uint_as_float(r[8 * k + 5]) + bb.y + xb.y;
          x[6] = __uint_as_float(r[8 * k + 6]) + bb.z + xb.z; x[7] = __uint_as_float(r[8 * k + 7]) + bb.w + xb.w;
          ptx::sts128f(s0, x[0], x[1], x[2], x[3]);
          ptx::sts128f(s1a, x[4], x[5], x[6], x[7]);
          if (k & 1) {       // one proxy fence per pair of chunks
            ptx::fence_proxy_async();
            ptx::mbar_arrive(&st_full[k - 1]);
            ptx::mbar_arrive(&st_full[k]);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) { s1 += x[e]; s2 = fmaf(x[e], x[e], s2); }
        }
        if (DBG) { const long long t1 = clock64(); tk[10] += t1 - t_ph; t_ph = t1; }
        if (has_ln) {
          // LayerNorm over the 192-wide row: four threads per row exchange partial sums through shared memory
          ln_part[cq * BM + row] = make_float2(s1, s2);
          ptx::bar_sync(1, EPI_THREADS);
          if (DBG) { const long long t1 = clock64(); tk[11] += t1 - t_ph; t_ph = t1; }
          s1 = 0.f; s2 = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) { const float2 o2 = ln_part[k * BM + row]; s1 += o2.x; s2 += o2.y; }
          const float mean = s1 * (1.0f / D);
          const float var = fmaxf(s2 * (1.0f / D) - mean * mean, 0.f);
          const float rstd = 1.0f / sqrtf(var + LN_EPS);
          if (cq == 0 && grow < p.M && p.ln_mean[g] != nullptr) { p.ln_mean[g][grow] = mean; p.ln_rstd[g][grow] = rstd; }
          V2S_WAIT(7, xn_free, nl & 1, 64);
          ++nl;
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            const int bq = t < 2 ? 4 + t : 0;
            uint32_t w[8];
#pragma unroll
            for (int hk = 0; hk < 2; ++hk) {
              const int k = 2 * t + hk;
              const float4 ga = __ldg(reinterpret_cast<const float4*>(p.ln_gamma[g] + k * 32 + cq * 8));
              const float4 gb = __ldg(reinterpret_cast<const float4*>(p.ln_gamma[g] + k * 32 + cq * 8 + 4));
              const float4 ea = __ldg(reinterpret_cast<const float4*>(p.ln_beta[g] + k * 32 + cq * 8));
              const float4 eb4 = __ldg(reinterpret_cast<const float4*>(p.ln_beta[g] + k * 32 + cq * 8 + 4));
              const float* x = v + 8 * k;
              w[4 * hk] = LP::pack((x[0] - mean) * rstd * ga.x + ea.x, (x[1] - mean) * rstd * ga.y + ea.y);
              w[4 * hk + 1] = LP::pack((x[2] - mean) * rstd * ga.z + ea.z, (x[3] - mean) * rstd * ga.w + ea.w);
              w[4 * hk + 2] = LP::pack((x[4] - mean) * rstd * gb.x + eb4.x, (x[5] - mean) * rstd * gb.y + eb4.y);
              w[4 * hk + 3] = LP::pack((x[6] - mean) * rstd * gb.z + eb4.z, (x[7] - mean) * rstd * gb.w + eb4.w);
            }
            // tile t holds columns [64 t, +64): chunk 2t -> 16-byte piece cq, chunk 2t+1 -> piece 4 + cq
            ptx::sts128(eb_row + bq * KBLK + ((cq ^ rx) << 4), w[0], w[1], w[2], w[3]);
            ptx::sts128(eb_row + bq * KBLK + (((4 + cq) ^ rx) << 4), w[4], w[5], w[6], w[7]);
          }
          ptx::fence_proxy_async();
          ptx::mbar_arrive(xn_full);
        }
      } else {
        V2S_WAIT(7, &a2_free[0], (c & 1) ^ 1, 65);       // EB0: the last chunk's du store and UMMAs have read it
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const int bq = t < 2 ? 4 + t : 0;
          const uint32_t* x = r + 16 * t;
          ptx::sts128(eb_row + bq * KBLK + ((cq ^ rx) << 4), LP::pack(__uint_as_float(x[0]), __uint_as_float(x[1])),
                      LP::pack(__uint_as_float(x[2]), __uint_as_float(x[3])), LP::pack(__uint_as_float(x[4]), __uint_as_float(x[5])),
                      LP::pack(__uint_as_float(x[6]), __uint_as_float(x[7])));
          ptx::sts128(eb_row + bq * KBLK + (((4 + cq) ^ rx) << 4), LP::pack(__uint_as_float(x[8]), __uint_as_float(x[9])),
                      LP::pack(__uint_as_float(x[10]), __uint_as_float(x[11])), LP::pack(__uint_as_float(x[12]), __uint_as_float(x[13])),
                      LP::pack(__uint_as_float(x[14]), __uint_as_float(x[15])));
        }
        ptx::fence_proxy_async();
        ptx::mbar_arrive(xn_full);
      }
      if (DBG) tk[8] += clock64() - t_e2;
    }
    if (DBG && blockIdx.x == 0 && threadIdx.x == 128) {
      for (int k = 0; k < 9; ++k) p.dbg[8 + k] = tk[k];
      p.dbg[17] = clock64() - t_begin;
      p.dbg[18] = tk[9]; p.dbg[19] = tk[10]; p.dbg[20] = tk[11];
    }
  }
#undef V2S_WAIT

  if (p.late_wait) { ptx::pdl_wait(); ptx::pdl_launch_dependents(); }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

template <int MODE, typename LP>
int launch_impl(const MlpParams& p, int grid, cudaStream_t stream) {
  static bool attr_set[MAX_DEVICES] = {false};
  if (!attr_set[cur_device()]) {
    V2S_CUDA_OK(cudaFuncSetAttribute(mlp_tc_kernel<MODE, LP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    V2S_CUDA_OK(cudaFuncSetAttribute(mlp_tc_kernel<MODE, LP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set[cur_device()] = true;
  }
  if (p.dbg) V2S_CUDA_OK(launch_pdl(mlp_tc_kernel<MODE, LP, true>, dim3(grid), dim3(N_THREADS), (size_t)SMEM_BYTES, stream, p));
  else V2S_CUDA_OK(launch_pdl(mlp_tc_kernel<MODE, LP, false>, dim3(grid), dim3(N_THREADS), (size_t)SMEM_BYTES, stream, p));
  V2S_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int launch_mlp_tc(const MlpDesc& d, cudaStream_t stream) {
  if (!tc_enabled()) { set_error("mlp_tc: tensor-core path not initialised"); return 1; }
  if (d.groups < 1 || d.groups > MAXG || d.M < 1) { set_error("mlp_tc: bad shape"); return 1; }
  MlpParams p;
  memset(&p, 0, sizeof(p));
  p.M = d.M;
  p.tiles_m = (d.M + BM - 1) / BM;
  p.total_tiles = p.tiles_m * d.groups;
  p.late_wait = d.late_wait ? 1 : 0;
  p.err_flag = tc_err_flag();
  p.dbg = tc_dbg_counters();
  for (int g = 0; g < d.groups; ++g) {
    if (!d.a[g] || !d.w1[g] || !d.w2[g] || !d.out[g]) { set_error("mlp_tc: null operand (group %d)", g); return 1; }
    V2S_TRY(tmap_get_2d(&p.tmA[g], d.a[g], D, d.M, D, 64, BM, true, 128));
    if (d.mode == MLP_FWD) {
      if (!d.b1[g] || !d.b2[g] || !d.resid[g]) { set_error("mlp_tc: forward needs biases and the residual"); return 1; }
      V2S_TRY(tmap_get_2d(&p.tmB1[g], d.w1[g], D, DF, D, 64, CW, true, 128));       // W1 [768,192]: rows x k
      V2S_TRY(tmap_get_2d(&p.tmB2[g], d.w2[g], DF, D, DF, 64, D, true, 128));       // W2 [192,768]: rows x k
      V2S_TRY(tmap_get_2d(&p.tmRes[g], d.resid[g], D, d.M, D, 32, BM, false, 128));
      V2S_TRY(tmap_get_2d(&p.tmOut[g], d.out[g], D, d.M, D, 32, BM, false, 128));
      if (d.h[g]) { V2S_TRY(tmap_get_2d(&p.tmH[g], d.h[g], DF, d.M, DF, 64, BM, true, 128)); p.has_h[g] = 1; }
      if (d.u[g]) { V2S_TRY(tmap_get_2d(&p.tmU[g], d.u[g], DF, d.M, DF, 64, BM, true, 128)); p.has_u[g] = 1; }
      if (d.ln_out[g]) {
        if (!d.ln_gamma[g] || !d.ln_beta[g]) { set_error("mlp_tc: fused LayerNorm needs gamma and beta"); return 1; }
        V2S_TRY(tmap_get_2d(&p.tmXn[g], d.ln_out[g], D, d.M, D, 64, BM, true, 128));
        p.has_ln[g] = 1;
      }
    } else {
      if (!d.u[g] || !d.h[g]) { set_error("mlp_tc: backward needs u and the du buffer"); return 1; }
      V2S_TRY(tmap_get_2d(&p.tmB1[g], d.w2[g], DF, D, DF, 64, 64, true, 128));      // W2 [192(k),768(n)]
      V2S_TRY(tmap_get_2d(&p.tmB2[g], d.w1[g], D, DF, D, 64, 64, true, 128));       // W1 [768(k),192(n)]
      V2S_TRY(tmap_get_2d(&p.tmH[g], d.h[g], DF, d.M, DF, 64, BM, true, 128));
      V2S_TRY(tmap_get_2d(&p.tmU[g], d.u[g], DF, d.M, DF, 64, BM, true, 128));
      V2S_TRY(tmap_get_2d(&p.tmOut[g], d.out[g], D, d.M, D, 64, BM, true, 128));
      p.has_h[g] = 1; p.has_u[g] = 1;
    }
    p.b1[g] = d.b1[g]; p.b2[g] = d.b2[g];
    p.ln_gamma[g] = d.ln_gamma[g]; p.ln_beta[g] = d.ln_beta[g]; p.ln_mean[g] = d.ln_mean[g]; p.ln_rstd[g] = d.ln_rstd[g];
  }
  const int sms = tc_num_sms();
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  if (d.mode == MLP_FWD)
    return d.lp_f16 ? launch_impl<MLP_FWD, LpF16>(p, grid, stream) : launch_impl<MLP_FWD, LpBf16>(p, grid, stream);
  return d.lp_f16 ? launch_impl<MLP_BWD, LpF16>(p, grid, stream) : launch_impl<MLP_BWD, LpBf16>(p, grid, stream);
}

}  // namespace v2s
