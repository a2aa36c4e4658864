// Exact mode on tensor cores: fp32 grid tensors and fp32 weights, every operand split into two fp16 terms (hi = fp16(v),
// lo = fp16(v - hi)), three tcgen05.mma per 16-channel step -- a_hi * w_hi + a_lo * w_hi + a_hi * w_lo, fp32 accumulation in TMEM.
// The dropped a_lo * w_lo term is 2^-22 of the product, so the result is fp32-grade (the <= 1e-3 gate of the image is met with three
// orders of magnitude to spare) at a third of the fp16 MMA rate instead of CUDA-core FFMA speed (conv_direct.cuh, which stays as the
// cross-check).  Same itg_conv_desc, same packed fp32 weights [tap][n_pad][k_pad], same epilogue code as the direct kernel: all modes
// (3x3, 1x1, folded up-sampling conv), any K, windows into wider buffers, residuals, SSM modulation, image / pre-tanh outputs.
//
// One CTA = 128 output-grid pixels (TH x TW tile) x <= 64 GEMM columns x one up-sampling phase; thread r owns pixel r (= TMEM lane r).
// Per (tap, 16-channel chunk): the threads fetch the tile's 128 x 16 values in memory order (the local-padding gather is the tap offset into
// the framed tensor), split them and writes the two K-major no-swizzle operand tiles; the weight chunk is split the same way; one thread issues
// the three MMAs and commits to the stage's mbarrier.  Two stages; several CTAs per SM (28 KB of shared memory, 64 TMEM columns each)
// hide the load latency.  Not tuned: it replaces a CUDA-core kernel, not one of the fp16 kernels.  Measured: cfg2 3.4 ms per pass against
// 7.8 ms on CUDA cores (0.38 ms of both is the fp32 attention kernel), a 15 x 15-patch SSM window 17.9 against 77.1 ms.  The thin
// full-resolution layers are bound by the LSU's line requests of the per-tap fp32 gathers: fetching in memory order instead of one pixel
// per thread gave 3.9 -> 3.4 / 23.9 -> 17.9 ms, while 32 channels per step with the next step's loads in flight, and a persistent variant
// packing two taps per step, changed nothing / were slower.  The next step would be the halo-tile load of conv_tile.cuh.
#pragma once
#include "ssm_fused.cuh"

namespace itg {

constexpr int SPLIT_NB = 64;                                   // GEMM columns per CTA
constexpr int SPLIT_A_BYTES = 2 * 128 * 16;                    // one 16-channel A tile: 2 planes x 128 pixels x 8 channels (fp16)
constexpr int SPLIT_B_BYTES = 2 * SPLIT_NB * 16;
constexpr int SPLIT_STAGE = 2 * SPLIT_A_BYTES + 2 * SPLIT_B_BYTES;      // hi + lo of both operands: 12 KB
constexpr int SPLIT_SMEM = 128 + 2 * SPLIT_STAGE + 128;

struct SplitParams {
  const float* in;
  int in_h, in_w, in_pitch, in_c, in_c_off, k;
  const float* w;
  int n_pad, k_pad;
  int mode;
  int tw_log2;          // tile width = 1 << tw_log2, tile height = 128 >> tw_log2
  int tiles_x;
  EpiParams ep;
};

__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(v[2 * i] - back.x, v[2 * i + 1] - back.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(128) conv_split_kernel(const SplitParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t bar_stage = sbase;                // [2] MMAs that read the stage have completed
  const uint32_t bar_done = sbase + 16;
  const uint32_t tmem_slot = sbase + 32;
  const uint32_t stages = sbase + 128;
  const int r = threadIdx.x, warp = r >> 5;

  const int tile = blockIdx.x;
  const int n0 = blockIdx.y * SPLIT_NB;
  const int nb = (p.n_pad - n0) < SPLIT_NB ? (p.n_pad - n0) : SPLIT_NB;      // multiple of 16
  const int phase = blockIdx.z;
  const int tw = 1 << p.tw_log2;
  const int tx = r & (tw - 1), ty = r >> p.tw_log2;
  const int y = (tile / p.tiles_x) * (128 >> p.tw_log2) + ty;
  const int x = (tile % p.tiles_x) * tw + tx;
  const bool valid = (y < p.in_h) && (x < p.in_w);

  if (r == 0) {
    mbar_init(bar_stage, 1);
    mbar_init(bar_stage + 8, 1);
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const uint32_t idesc = (1u << 4) | ((uint32_t)(nb >> 3) << 17) | ((128u >> 4) << 24);      // fp16 x fp16 -> fp32, M = 128, N = nb
  const int ntaps = (p.mode == ITG_CONV3X3) ? 9 : (p.mode == ITG_CONV1X1 ? 1 : 4);
  int it = 0;
  for (int t = 0; t < ntaps; ++t) {
    int dy, dx, wt;
    tap_offsets(p.mode, phase, t, dy, dx, wt);
    // gather in memory order: thread r fetches 8-channel group (r & 1) of pixels (r >> 1) and 64 + (r >> 1) of the tile -- a warp
    // instruction covers 16 consecutive pixels x 64 contiguous bytes instead of 32 pixels x 16 bytes (a quarter of the line requests)
    const float* a_src[2];
    bool a_ok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int px = i * 64 + (r >> 1);
      const int py = (tile / p.tiles_x) * (128 >> p.tw_log2) + (px >> p.tw_log2), pxx = (tile % p.tiles_x) * tw + (px & (tw - 1));
      a_ok[i] = (py < p.in_h) && (pxx < p.in_w);
      a_src[i] = p.in + grid_off_pitch(a_ok[i] ? py + dy : 0, a_ok[i] ? pxx + dx : 0, p.in_pitch, p.in_c, p.in_c_off) + 8 * (r & 1);
    }
    const float* w_tap = p.w + ((size_t)wt * p.n_pad + n0) * (size_t)p.k_pad;
    for (int k0 = 0; k0 < p.k; k0 += 16, ++it) {
      const int s = it & 1;
      const uint32_t a_hi = stages + (uint32_t)(s * SPLIT_STAGE), a_lo = a_hi + SPLIT_A_BYTES;
      const uint32_t b_hi = a_lo + SPLIT_A_BYTES, b_lo = b_hi + SPLIT_B_BYTES;
      // 8 channels of two pixels (k and the channel offsets are multiples of 8; channels beyond k read as zeros)
      float a[2][8];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (a_ok[j] && k0 + 8 * (r & 1) < p.k) load8(a_src[j] + k0, a[j]);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) a[j][i] = 0.f;
        }
      }
      // weight chunk [nb][16]: thread -> (row, 8-channel half)
      float wv[8];
      const int wr = r >> 1, wj = r & 1;
      const bool w_ok = wr < nb;
      if (w_ok && k0 + 8 * wj < p.k_pad) load8(w_tap + (size_t)wr * p.k_pad + k0 + 8 * wj, wv);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) wv[i] = 0.f;
      }
      if (it >= 2) mbar_wait(bar_stage + 8 * s, (uint32_t)((it >> 1) - 1) & 1u);      // the MMAs of iteration it - 2 have read this stage
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint4 hi, lo;
        split8(a[j], hi, lo);
        sts128(a_hi + (uint32_t)((r & 1) * 2048 + (j * 64 + (r >> 1)) * 16), hi.x, hi.y, hi.z, hi.w);
        sts128(a_lo + (uint32_t)((r & 1) * 2048 + (j * 64 + (r >> 1)) * 16), lo.x, lo.y, lo.z, lo.w);
      }
      if (w_ok) {
        uint4 hi, lo;
        split8(wv, hi, lo);
        sts128(b_hi + (uint32_t)(wj * SPLIT_NB * 16 + wr * 16), hi.x, hi.y, hi.z, hi.w);
        sts128(b_lo + (uint32_t)(wj * SPLIT_NB * 16 + wr * 16), lo.x, lo.y, lo.z, lo.w);
      }
      fence_proxy_async();                          // generic writes -> async proxy (tensor core) reads
      tc_fence_before();
      __syncthreads();
      if (r == 0) {
        tc_fence_after();
        const uint64_t dah = desc_noswz(a_hi >> 4, 2048 / 16, 8), dal = desc_noswz(a_lo >> 4, 2048 / 16, 8);
        const uint64_t dbh = desc_noswz(b_hi >> 4, SPLIT_NB, 8), dbl = desc_noswz(b_lo >> 4, SPLIT_NB, 8);
        umma_f16(tmem_base, dal, dbh, idesc, it > 0 ? 1u : 0u);      // small terms first
        umma_f16(tmem_base, dah, dbl, idesc, 1u);
        umma_f16(tmem_base, dah, dbh, idesc, 1u);
        umma_commit(bar_stage + 8 * s);
      }
    }
  }
  if (r == 0) umma_commit(bar_done);
  mbar_wait(bar_done, 0u);
  tc_fence_after();

  int oy = y, ox = x;
  if (p.mode == ITG_UPCONV) { oy = 2 * y + (phase >> 1); ox = 2 * x + (phase & 1); }
  const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < nb; c += 16) {
    float v[16];
    tmem_ld16(trow + (uint32_t)c, v);               // warp-collective: every lane takes part, valid or not
    if (!valid) continue;
    if (p.ep.mod_x != nullptr) {
      epilogue_ssm16<float>(p.ep, oy, ox, n0 + c, v);
    } else {
      float a8[8], b8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { a8[i] = v[i]; b8[i] = v[8 + i]; }
      epilogue8<float>(p.ep, oy, ox, n0 + c, a8);
      epilogue8<float>(p.ep, oy, ox, n0 + c + 8, b8);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

}  // namespace itg
