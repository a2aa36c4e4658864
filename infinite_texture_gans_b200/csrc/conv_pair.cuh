// 3x3 local-padding convolution (conv2d_lp, models/layers.py:29-36) with K <= 128 input channels per tap on CTA PAIRS
// (tcgen05.mma.cta_group::2): the pipeline of the SSM pair kernel (ssm_fused2.cuh) fed from global memory instead of from the
// mlp_shared GEMM.  It serves the 104 -> 52 / 52 -> 52 / 52 -> 26 / 26 -> 26 layers of the 34 Generator's last two blocks.
//
// What it replaces.  The streaming kernel (conv_umma.cuh) fetches a fresh 128-pixel activation tile PER TAP (9 x 32 KB of A plus
// 9 x 16 KB of weights per tile out of L2: block4.conv1 of cfg3 ran at the L2's bandwidth, 14.6 k cycles per tile for 3.5 k cycles of
// MMA); the thin-layer kernel (conv_tile.cuh) issues M = 128 MMAs whose cost is the 40-cycle issue / operand floor whatever N is.
// Here, per 16 x 8 output tile and CTA:
//   * six loader warps fetch the tile's (16+2) x (8+2) halo ONCE (16-byte cp.async, zero-fill outside the buffer) into K/8 planes of
//     [halo pixel][8 channels]; the nine taps are nine 16-byte-granular shifts of the MMA's no-swizzle A descriptor;
//   * the weights of all taps for the CTA's half of the pair's <= 128 GEMM columns (<= 144 KB) stay in shared memory for the whole
//     launch: in steady state the kernel reads each activation once (+ halo) and writes its outputs;
//   * planes are handed over in groups of eight (64 channels; four for K = 32) through a ring of three (six) group slots, all but one of
//     them in flight per loader warp: 1.5 / 3 / 6 tiles of K = 128 / 64 / 32 are buffered;
//   * one instruction covers M = 256 pixels (both CTAs' tiles) x N columns: half the instructions of the single-CTA kernels for the
//     same tile, each CTA reading its own A and only its half of B;
//   * the leader CTA's MMA warp issues from an elect-guarded block (uniform-datapath descriptors); commits are cluster-multicast,
//     loaders / epilogue warps of both CTAs arrive on the leader's mbarriers;
//   * eight epilogue warps per CTA drain the TMEM accumulator ring (2 or 4 buffers) through the shared fused epilogue (bias, residual,
//     BN + activation, raw / activated outputs, frame: epilogue8, itg_common.cuh).
#pragma once
#include "ssm_fused2.cuh"

namespace itg {

constexpr int PAIR_KG_MAX = 16;                                        // 8-channel planes per tile (K <= 128)
constexpr int PAIR_PLANE = PLANE_BYTES + 16;                           // plane pitch 2896 B = 16 (mod 128): the eight 16-byte chunks of a pixel, written by
                                                                       // eight lanes of one cp.async to eight planes, land in eight different bank groups
constexpr int PAIR_A_PLANES = 24;                                      // ring of group slots: 3 slots x 8 planes (K >= 64) or 6 slots x 4 planes (K = 32)
constexpr int PAIR_HDR = 1024;
constexpr int PAIR_OFF_A = PAIR_HDR;
constexpr int PAIR_OFF_W = PAIR_OFF_A + PAIR_A_PLANES * PAIR_PLANE;    // [tap 9][k-group 16, kg used][64 rows, n_half used][16 B]
constexpr int PAIR_LOADERS = 6;                                        // warps 8..13
static_assert(PAIR_OFF_W % 128 == 0, "operand alignment");

constexpr int PAIR_SMEM = PAIR_OFF_W + 9 * PAIR_KG_MAX * SSM_NBLK_MAX * 16 + 1024;     // 219 008 B: one CTA per SM

struct PairParams {
  int m_h, m_w;            // M-grid size (input interior == output size)
  int tiles_x, ntiles;
  const void* in;          // framed grid tensor (buffer origin)
  int in_c, in_pitch;      // storage channels, pixels per buffer row
  int buf_h, buf_w;        // buffer extent in pixels (interior + frame)
  int in_cg_off;           // first 8-channel group of the input slice
  int kg;                  // 8-channel planes per tile: 4, 8 or 16 (k_pad / 8)
  int ksteps;              // K = 16 steps that hold real channels (<= kg / 2)
  const void* w;           // [9][n_pad][k_pad] operand dtype
  int n_pad, k_pad;
  int n_blk, nblocks;      // GEMM columns per CTA pair (each CTA parks n_blk / 2), column blocks
  int nbuf;                // TMEM accumulator buffers: 2 or 4 (nbuf * n_blk <= 256)
  int inflight;            // groups a loader warp keeps in flight (1 .. ring slots - 1)
  uint32_t idesc;
  int exp;                 // developer experiments (ITG_TILE_EXP with ITG_TILE_DBG; WRONG RESULTS, timing only): 1 no loads, 2 no epilogue memory traffic, 4 one tap
  unsigned long long* dbg; // optional [16] cycle counters of CTA 0 (ITG_TILE_DBG=1 on a -DITG_SSM_DBG build), NULL in production
  EpiParams ep;
};

template <typename T, int F>
__global__ void __launch_bounds__(SSM_THREADS, 1)
conv_pair_kernel(const PairParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;        // the dynamic window starts at the same offset in both CTAs
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                              // 0 = leader
  const int n_half = p.n_blk >> 1;
  const int gp = p.kg < 8 ? p.kg : 8;                                    // planes per hand-over group (= ring slot): 4 (K = 32) or 8
  const int groups = p.kg / gp;                                          // groups per tile: 2 (K = 128) or 1
  const uint32_t nring = (uint32_t)(PAIR_A_PLANES / gp);                 // ring slots: 6 or 3

  const uint32_t bar_a_full = sbase;               // [8]  loaders of both CTAs -> leader's MMA warp        (count 12)
  const uint32_t bar_a_empty = sbase + 64;         // [8]  MMA commit (multicast) -> loaders
  const uint32_t bar_acc_full = sbase + 128;       // [4]  MMA commit (multicast) -> epilogue
  const uint32_t bar_acc_empty = sbase + 160;      // [4]  epilogue warps of both CTAs -> leader's MMA warp (count 16)
  const uint32_t tmem_slot = sbase + 192;

  const int pair = (int)blockIdx.x >> 1, npairs = (int)gridDim.x >> 1;
  const int nbp = pair % p.nblocks;                // the pair's block of GEMM columns
  const int slot = pair / p.nblocks, nslots = npairs / p.nblocks;
  const int npt = (p.ntiles + 1) >> 1;             // pair-tiles
  const int n_my = (slot < npt && slot < nslots) ? (npt - slot + nslots - 1) / nslots : 0;
  const int n0 = nbp * p.n_blk;                    // first GEMM column of the pair's block

  pdl_launch_dependents();
  if (warp == SSM_WARP_MMA && lane == 0) {
    for (int i = 0; i < 8; ++i) {
      mbar_init(bar_a_full + 8 * i, 2 * PAIR_LOADERS);
      mbar_init(bar_a_empty + 8 * i, 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 16);
    }
    fence_barrier_init();
  }
  if (warp == SSM_WARP_PROD) tmem_alloc2(tmem_slot, 256);

  // ---- park this CTA's half of the weights: rows [n0 + rank * n_half, + n_half) of every tap, fixed 64-row pitch (compile-time descriptor
  //      offsets in the MMA loop).  Weights are launch constants: read before griddepcontrol.wait, overlapping the previous launch's tail ----
  {
    const T* wg = reinterpret_cast<const T*>(p.w);
    const int chunks = 9 * p.kg * n_half;
    const uint32_t ws = sbase + PAIR_OFF_W;
    for (int i = threadIdx.x; i < chunks; i += SSM_THREADS) {
      const int j = i % p.kg, n = (i / p.kg) % n_half, t = i / (p.kg * n_half);
      const int ng = n0 + (int)rank * n_half + n;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (ng < p.n_pad) v = *reinterpret_cast<const uint4*>(wg + ((size_t)t * p.n_pad + ng) * p.k_pad + j * 8);
      sts128(ws + (uint32_t)(((t * PAIR_KG_MAX + j) * SSM_NBLK_MAX + n) * 16), v.x, v.y, v.z, v.w);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();                                       // the previous launch's outputs (our activations, residual) are complete from here on

  if (warp == SSM_WARP_MMA) {
    if (rank == 0) {
      // ---- MMA warp of the leader: issues for both CTAs ----
      const uint32_t w16 = (sbase + PAIR_OFF_W) >> 4, a16 = (sbase + PAIR_OFF_A) >> 4;
      constexpr uint32_t nh16 = SSM_NBLK_MAX;
      constexpr uint32_t tap16 = PAIR_KG_MAX * nh16;                    // one tap of the weight image, in 16-byte units (fixed pitch: literal offsets)
      uint32_t s = 0, sph = 0;                                          // ring slot and its phase
      unsigned long long dacc[3] = {0, 0, 0};
      long long tl = p.dbg ? clock64() : 0;
      for (int it = 0; it < n_my; ++it) {
        const int b = it & (p.nbuf - 1);
        if (lane == 0) mbar_wait(bar_acc_empty + 8 * b, (((uint32_t)it / (uint32_t)p.nbuf) & 1u) ^ 1u);
        __syncwarp();
        ITG_SACC(0, tl);
        const uint32_t d = tmem_base + (uint32_t)(b * p.n_blk);
#pragma unroll 1
        for (int g = 0; g < groups; ++g) {
          if (lane == 0) mbar_wait(bar_a_full + 8 * s, sph);
          __syncwarp();
          ITG_SACC(1, tl);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t ak = a16 + s * (uint32_t)gp * (PAIR_PLANE / 16);
            const uint32_t wk = w16 + (uint32_t)(gp * g) * nh16;
            const int ks0 = (gp >> 1) * g;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              if (2 * k4 < gp && ks0 + k4 < p.ksteps) {
#pragma unroll
                for (int t = 0; t < ((p.exp & 4) ? 1 : 9); ++t) {
                  const uint32_t shift16 = (uint32_t)((t / 3) * HALO_W + (t % 3));
                  umma2_f16(d, desc_noswz(ak + (uint32_t)(2 * k4) * (PAIR_PLANE / 16) + shift16, PAIR_PLANE / 16, HALO_W),
                            desc_noswz(wk + (uint32_t)(2 * k4) * nh16 + (uint32_t)t * tap16, nh16, 8), p.idesc, (ks0 + k4 > 0 || t > 0) ? 1u : 0u);
                }
              }
            }
            umma2_commit(bar_a_empty + 8 * s);
            if (g == groups - 1) umma2_commit(bar_acc_full + 8 * b);
          }
          __syncwarp();
          ITG_SACC(2, tl);
          if (++s == nring) { s = 0; sph ^= 1u; }
        }
      }
      if (p.dbg && blockIdx.x == 0 && lane == 0) for (int i = 0; i < 3; ++i) p.dbg[i] = dacc[i];
    }
  } else if (warp >= SSM_WARP_CVT && warp < SSM_WARP_CVT + PAIR_LOADERS) {
    // ---- loaders (both CTAs, each for its own tile): halo tile -> planes, one group of four planes (32 channels) per hand-over.
    //      All but one ring slot stay in flight (cp.async groups complete in order: the oldest is published when the window is
    //      full); everything that has landed is published before the warp sleeps on a free slot. ----
    const int lt = (warp - SSM_WARP_CVT) * 32 + lane;                   // 0..191
    const int cg_total = p.in_c >> 3;
    const T* in = reinterpret_cast<const T*>(p.in);
    const uint32_t a_smem = sbase + PAIR_OFF_A;
    uint32_t s = 0, sph = 0;                                            // slot being filled, its phase
    uint32_t ps = 0;                                                    // oldest unpublished slot
    int unpub = 0;                                                      // committed, unpublished groups
    unsigned long long dacc[3] = {0, 0, 0};
    long long tl = p.dbg ? clock64() : 0;
    auto publish = [&](int keep) {                                      // publish until at most `keep` groups are unpublished
      if (unpub > keep) {
        cp_async_wait_dyn(keep);
        fence_proxy_async();
        __syncwarp();
        while (unpub > keep) {
          if (lane == 0) mbar_arrive_cluster(bar_a_full + 8 * ps, 0u);
          if (++ps == nring) ps = 0;
          --unpub;
        }
      }
    };
    int pt = slot;
    for (int it = 0; it < n_my; ++it, pt += nslots) {
      const int tile = 2 * pt + (int)rank;
      const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
      const int y0 = ty * TILE_H, x0 = tx * TILE_W;                    // halo origin in buffer pixels
      const bool tile_ok = tile < p.ntiles;                             // odd tile count: the last pair's second CTA feeds zeros
      for (int g = 0; g < groups; ++g) {
        uint32_t ready = 0;
        if (lane == 0) ready = mbar_try_wait(bar_a_empty + 8 * s, sph ^ 1u) ? 1u : 0u;
        ready = __shfl_sync(0xffffffffu, ready, 0);
        if (!ready) {
          publish(0);
          ITG_SACC(2, tl);
          if (lane == 0) mbar_wait(bar_a_empty + 8 * s, sph ^ 1u);      // the MMAs that read this slot have completed
          __syncwarp();
        }
        ITG_SACC(0, tl);
        // lane -> (pixel, chunk) with the chunk index fastest: the gp lanes of a pixel read gp x 16 contiguous bytes and consecutive pixels
        // of a halo row are adjacent in memory -- few 128-byte lines per instruction (tools/ldgsts_probe.cu: 8-11 cycles per warp
        // instruction against 19 for four-lane runs and 46 for one pixel per lane)
        const uint32_t dst = a_smem + s * (uint32_t)(gp * PAIR_PLANE);
        const int gsh = gp == 8 ? 3 : 2;
#pragma unroll 2
        for (int idx = lt; idx < ((p.exp & 1) ? 0 : (HALO_PX << gsh)); idx += PAIR_LOADERS * 32) {
          const int px = idx >> gsh, j = idx & (gp - 1);
          const int hy = (px * 205) >> 11, hx = px - hy * HALO_W;
          const int cg = p.in_cg_off + gp * g + j;
          const bool valid = tile_ok && cg < cg_total && (y0 + hy < p.buf_h) && (x0 + hx < p.buf_w);
          const T* src = valid ? in + ((size_t)(y0 + hy) * p.in_pitch + (x0 + hx)) * (size_t)p.in_c + cg * 8 : in;
          cp_async16_zfill(dst + (uint32_t)(j * PAIR_PLANE + px * 16), src, valid);
        }
        cp_async_commit();
        ++unpub;
        ITG_SACC(1, tl);
        publish(p.inflight - 1);
        ITG_SACC(2, tl);
        if (++s == nring) { s = 0; sph ^= 1u; }
      }
    }
    publish(0);
    if (p.dbg && blockIdx.x == 0 && warp == SSM_WARP_CVT && lane == 0) for (int i = 0; i < 3; ++i) p.dbg[4 + i] = dacc[i];
  } else if (warp < 8) {
    // ---- epilogue (both CTAs): own tile x all columns of the pair's block; group eg takes the 16-column chunks c = eg, eg + 2, ...
    //      (<= 4 per warp).  A same-resolution 16-bit residual is fetched before the warp waits for the accumulator. ----
    constexpr bool PRE = (F & EF_RES) != 0 && (F & EF_GENERIC) == 0;
    const int eg = warp >> 2, q = warp & 3;
    const int row = q * 32 + lane;
    const EpiParams& ep = p.ep;
    unsigned long long dacc[2] = {0, 0};
    long long tl = p.dbg ? clock64() : 0;
    int pt = slot;
    for (int it = 0; it < n_my; ++it, pt += nslots) {
      const int b = it & (p.nbuf - 1);
      const int tile = 2 * pt + (int)rank;
      const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
      const int y = ty * TILE_H + (row >> 3), x = tx * TILE_W + (row & 7);
      const bool valid = tile < p.ntiles && (y < p.m_h) && (x < p.m_w);
      uint4 pre[8];
      if (PRE && valid && !(p.exp & 2)) {
        const T* rp = reinterpret_cast<const T*>(ep.res) + grid_off(y >> ep.res_shift, x >> ep.res_shift, ep.res_w, ep.res_c, 0);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int n = n0 + 16 * (eg + 2 * cc);
          if (16 * (eg + 2 * cc) < p.n_blk) {
            if (n < ep.out_c) pre[2 * cc] = *reinterpret_cast<const uint4*>(rp + n);
            if (n + 8 < ep.out_c) pre[2 * cc + 1] = *reinterpret_cast<const uint4*>(rp + n + 8);
          }
        }
      }
      if (lane == 0) mbar_wait(bar_acc_full + 8 * b, ((uint32_t)it / (uint32_t)p.nbuf) & 1u);
      __syncwarp();
      ITG_SACC(0, tl);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * p.n_blk);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = eg + 2 * cc;
        if (16 * c < p.n_blk && n0 + 16 * c < p.n_pad) {
          float v[16];
          tmem_ld16(trow + (uint32_t)(16 * c), v);
          if (valid && !(p.exp & 2)) {
            float a8[8], b8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { a8[i] = v[i]; b8[i] = v[8 + i]; }
            epilogue8<T, F>(ep, y, x, n0 + 16 * c, a8, PRE ? &pre[2 * cc] : nullptr);
            epilogue8<T, F>(ep, y, x, n0 + 16 * c + 8, b8, PRE ? &pre[2 * cc + 1] : nullptr);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar_acc_empty + 8 * b, 0u);
      ITG_SACC(1, tl);
    }
    if (p.dbg && blockIdx.x == 0 && warp == 0 && lane == 0) for (int i = 0; i < 2; ++i) p.dbg[8 + i] = dacc[i];
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // nobody frees tensor memory / exits while the peer may still signal or be read
  if (warp == SSM_WARP_PROD) tmem_dealloc2(tmem_base, 256);
}

}  // namespace itg
