// 3x3 (and 1x1) local-padding convolution (conv2d_lp, models/layers.py:29-36) with K <= 128 input channels per tap on CTA PAIRS
// (tcgen05.mma.cta_group::2): the pipeline of the SSM pair kernel (ssm_fused2.cuh) fed from global memory instead of from the
// mlp_shared GEMM.  It serves the 104 -> 104 / 104 -> 52 / 52 -> 52 / 52 -> 26 / 26 -> 26 layers and the 1x1 shortcuts of the 34 Generator's
// last blocks (MODE = ITG_CONV1X1: one tap, only the tile interior is fetched).
//
// What it replaces.  The streaming kernel (conv_umma.cuh) fetches a fresh 128-pixel activation tile PER TAP (9 x 32 KB of A plus
// 9 x 16 KB of weights per tile out of L2: block4.conv1 of cfg3 ran at the L2's bandwidth, 14.6 k cycles per tile for 3.5 k cycles of
// MMA); the thin-layer kernel (conv_tile.cuh) issues M = 128 MMAs whose cost is the 40-cycle issue / operand floor whatever N is, and
// its one-pixel-per-lane cp.async costs a 128-byte line request per lane.  Here, per 16 x 8 output tile and CTA:
//   * six loader warps fetch the tile's (16+2) x (8+2) halo ONCE into K/8 planes of [halo pixel][8 channels] (16-byte cp.async,
//     zero-fill outside the buffer); the nine taps are nine 16-byte-granular shifts of the MMA's no-swizzle A descriptor.  Lanes walk
//     the halo ROW (10 pixels x all channels = one contiguous run of global memory) 16 bytes at a time, so that a warp instruction
//     touches 4-5 lines instead of 32; the plane pitch of 2896 B (= 16 mod 128) spreads the chunks of a pixel over the bank groups
//     (tools/ldgsts_probe.cu: 17-24 cycles per warp instruction against 180 for one pixel per lane).  The lane -> (global offset,
//     shared offset) table is tile-independent and lives in registers;
//   * the weights of all taps for the CTA's half of the pair's <= 64 GEMM columns (80 KB with the shortcut's tenth tap) stay in shared memory for the whole launch:
//     in steady state the kernel reads each activation once (+ halo) and writes its outputs; wider layers run as column blocks;
//   * whole tiles are handed over through a ring of 3 (K = 104) / 5 (K = 64) / 8 (K = 32) slots, up to three in flight per loader warp;
//   * optional second input (itg_conv_desc.in2): the block's 1x1 shortcut folded into its conv2 -- the loaders also fetch the tile interior
//     of the shortcut's input into further planes of the slot, the MMA warp appends its k-steps (centre tap, weights of a tenth tap);
//   * one instruction covers M = 256 pixels (both CTAs' tiles) x N columns: half the instructions of the single-CTA kernels for the
//     same tile, each CTA reading its own A and only its half of B;
//   * the leader CTA's MMA warp issues from an elect-guarded block (uniform-datapath descriptors, literal offsets); commits are
//     cluster-multicast, loaders / epilogue warps of both CTAs arrive on the leader's mbarriers;
//   * eight epilogue warps per CTA drain a ring of four TMEM accumulators (bias, residual, BN + activation, raw / activated outputs,
//     frame -- the semantics of epilogue8, itg_common.cuh) with coalesced global accesses: a lane owns a pixel, but residual loads and
//     output stores go through a per-warp transposition buffer so that one instruction covers 8 consecutive pixels x 64 bytes.  The two
//     warps of a TMEM lane quarter split the columns of a tile, or alternate tiles when the pair has <= 32 columns.
#pragma once
#include "ssm_fused2.cuh"

namespace itg {

constexpr int PAIR_KG_MAX = 16;                                        // 8-channel planes per tile (K <= 128)
constexpr int PAIR_NH = 32;                                            // weight rows parked per CTA (row pitch of the weight image)
constexpr int PAIR_NBLK_MAX = 2 * PAIR_NH;                             // GEMM columns per CTA pair
constexpr int PAIR_PLANE = PLANE_BYTES + 16;                           // plane pitch 2896 B = 16 (mod 128), see TILE_PLANE
constexpr int PAIR_A_PLANES = 45;                                      // activation ring: 3 x 14 | 5 x 8 | 8 x 4 planes (2 x 16 for K = 128 exactly)
constexpr int PAIR_W_TAPS = 10;                                        // nine taps + the folded 1x1 shortcut
constexpr int PAIR_MAX_SLOTS = 8;
constexpr int PAIR_HDR = 1024;                                         // barriers | at 256: bias, scale, shift of the pair's 64 columns (fp32)
constexpr int PAIR_OFF_VEC = 256;
constexpr int PAIR_OFF_STAGE = PAIR_HDR;                               // 8 epilogue warps x [32 pixels][64 B]: transposition buffer of the coalesced stores
constexpr int PAIR_OFF_W = PAIR_OFF_STAGE + 8 * 2048;                  // [tap 10][k-group 16, kg used][32 rows, n_half used][16 B]
constexpr int PAIR_OFF_A = PAIR_OFF_W + PAIR_W_TAPS * PAIR_KG_MAX * PAIR_NH * 16;
constexpr int PAIR_LOADERS = 6;                                        // warps 8..13
constexpr int PAIR_LD_ITERS = 16;                                      // 16-byte chunks per loader thread and tile (launch_pair checks the count)
static_assert(PAIR_OFF_A % 128 == 0, "operand alignment");

constexpr int PAIR_SMEM = PAIR_OFF_A + PAIR_A_PLANES * PAIR_PLANE + 1024;     // 230 672 B: one CTA per SM
static_assert(PAIR_SMEM <= 227 * 1024, "shared memory budget");

struct PairParams {
  int m_h, m_w;            // M-grid size (input interior == output size)
  int tiles_x, ntiles;
  const void* in;          // framed grid tensor (buffer origin)
  int in_c, in_pitch;      // storage channels, pixels per buffer row
  int buf_h, buf_w;        // buffer extent in pixels (interior + frame)
  int in_cg_off;           // first 8-channel group of the input slice
  int np;                  // 8-channel planes loaded per tile: k / 8 (<= 16)
  const void* in2;         // optional second input (1x1 shortcut folded into a 3x3 conv): same geometry as `in`; NULL = off
  int in2_c, in2_cg_off;
  int np2, ksteps2;        // its planes per tile (interior only) and K = 16 steps; the planes follow the first input's 2 * ksteps in the slot
  const void* w2;          // [1][n_pad][k2_pad]
  int k2_pad;
  int slot_planes;         // planes per ring slot: 2 * (ksteps + ksteps2)
  int nring;               // ring slots (<= 8)
  int ksteps;              // K = 16 steps per tap
  const void* w;           // [9][n_pad][k_pad] operand dtype
  int n_pad, k_pad;
  int n_blk, nblocks;      // GEMM columns per CTA pair (each CTA parks n_blk / 2 <= 32), column blocks
  int nbuf;                // TMEM accumulator buffers (nbuf * n_blk <= 256)
  int inflight;            // tiles a loader warp keeps in flight (1 .. ring slots - 1)
  uint32_t idesc;
  int exp;                 // developer experiments (ITG_TILE_EXP with ITG_TILE_DBG; WRONG RESULTS, timing only): 1 no loads, 2 no epilogue memory traffic, 4 one tap
  unsigned long long* dbg; // optional [16] cycle counters of CTA 0 (ITG_TILE_DBG=1 on a -DITG_SSM_DBG build), NULL in production
  EpiParams ep;
};

template <typename T, int F, int MODE>
__global__ void __launch_bounds__(SSM_THREADS, 1)
conv_pair_kernel(const PairParams p) {
  static_assert(MODE == ITG_CONV3X3 || MODE == ITG_CONV1X1, "conv_pair: 3x3 or 1x1");
  constexpr int NTAPS = MODE == ITG_CONV3X3 ? 9 : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;        // the dynamic window starts at the same offset in both CTAs
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                              // 0 = leader
  const int n_half = p.n_blk >> 1;
  const uint32_t nring = (uint32_t)p.nring;
  const bool split_tiles = p.n_blk <= 32;            // epilogue: the two warps of a TMEM lane quarter alternate tiles (else they split the columns)
  const uint32_t slot_bytes = (uint32_t)(p.slot_planes * PAIR_PLANE);

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
    for (int i = 0; i < PAIR_MAX_SLOTS; ++i) {
      mbar_init(bar_a_full + 8 * i, 2 * PAIR_LOADERS);
      mbar_init(bar_a_empty + 8 * i, 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, split_tiles ? 8 : 16);
    }
    fence_barrier_init();
  }
  if (warp == SSM_WARP_PROD) tmem_alloc2(tmem_slot, 256);

  // ---- park this CTA's half of the weights: rows [n0 + rank * n_half, + n_half) of every tap, fixed 32-row pitch (literal descriptor
  //      offsets in the MMA loop).  Weights are launch constants: read before griddepcontrol.wait, overlapping the previous launch's tail.
  //      Planes of a slot beyond the loaded ones (an odd plane count) are read by the last k-step against zero weights: zero them once. ----
  {
    const T* wg = reinterpret_cast<const T*>(p.w);
    const int kg = p.k_pad >> 3;
    const int chunks = NTAPS * kg * n_half;
    const uint32_t ws = sbase + PAIR_OFF_W;
    for (int i = threadIdx.x; i < chunks; i += SSM_THREADS) {
      const int j = i % kg, n = (i / kg) % n_half, t = i / (kg * n_half);
      const int ng = n0 + (int)rank * n_half + n;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (ng < p.n_pad) v = *reinterpret_cast<const uint4*>(wg + ((size_t)t * p.n_pad + ng) * p.k_pad + j * 8);
      sts128(ws + (uint32_t)(((t * PAIR_KG_MAX + j) * PAIR_NH + n) * 16), v.x, v.y, v.z, v.w);
    }
    if (MODE == ITG_CONV3X3 && p.in2 != nullptr) {                     // the shortcut's weights: tap 9 of the image
      const T* w2g = reinterpret_cast<const T*>(p.w2);
      const int kg2 = p.k2_pad >> 3;
      for (int i = threadIdx.x; i < kg2 * n_half; i += SSM_THREADS) {
        const int j = i % kg2, n = i / kg2;
        const int ng = n0 + (int)rank * n_half + n;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (ng < p.n_pad) v = *reinterpret_cast<const uint4*>(w2g + (size_t)ng * p.k2_pad + j * 8);
        sts128(ws + (uint32_t)(((9 * PAIR_KG_MAX + j) * PAIR_NH + n) * 16), v.x, v.y, v.z, v.w);
      }
    }
    const int pad1 = 2 * p.ksteps - p.np, pad2 = 2 * p.ksteps2 - p.np2, pad_planes = pad1 + pad2;      // 0 or 1 each
    for (int i = threadIdx.x; i < p.nring * pad_planes * (PAIR_PLANE / 16); i += SSM_THREADS) {
      const int c = i % (PAIR_PLANE / 16), r = i / (PAIR_PLANE / 16);
      const int s = r / pad_planes, w = r % pad_planes;
      const int j = (w < pad1) ? p.np : 2 * p.ksteps + p.np2;
      sts128(sbase + PAIR_OFF_A + (uint32_t)s * slot_bytes + (uint32_t)(j * PAIR_PLANE + c * 16), 0u, 0u, 0u, 0u);
    }
    if (threadIdx.x < 3 * PAIR_NBLK_MAX) {                            // bias | scale | shift of this pair's columns (zeros / ones where absent)
      const int v = threadIdx.x / PAIR_NBLK_MAX, n = n0 + threadIdx.x % PAIR_NBLK_MAX;
      const float* src = v == 0 ? p.ep.bias : (v == 1 ? p.ep.scale : p.ep.shift);
      const float val = (src != nullptr && n < p.n_pad) ? src[n] : (v == 1 ? 1.f : 0.f);
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbase + PAIR_OFF_VEC + (uint32_t)threadIdx.x * 4u), "f"(val) : "memory");
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
      constexpr uint32_t nh16 = PAIR_NH;
      constexpr uint32_t tap16 = PAIR_KG_MAX * nh16;                    // one tap of the weight image, in 16-byte units
      uint32_t s = 0, sph = 0;                                          // ring slot and its phase
      unsigned long long dacc[3] = {0, 0, 0};
      long long tl = p.dbg ? clock64() : 0;
      for (int it = 0; it < n_my; ++it) {
        const int b = it & (p.nbuf - 1);
        if (lane == 0) {
          mbar_wait(bar_acc_empty + 8 * b, (((uint32_t)it / (uint32_t)p.nbuf) & 1u) ^ 1u);
          ITG_SACC(0, tl);
          mbar_wait(bar_a_full + 8 * s, sph);
        }
        __syncwarp();
        ITG_SACC(1, tl);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t d = tmem_base + (uint32_t)(b * p.n_blk);
          const uint32_t ak = a16 + s * (slot_bytes >> 4);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            if (ks < p.ksteps) {
#pragma unroll
              for (int t = 0; t < NTAPS; ++t) {
#ifdef ITG_SSM_DBG
                if (t > 0 && (p.exp & 4)) break;
#endif
                const uint32_t shift16 = MODE == ITG_CONV3X3 ? (uint32_t)((t / 3) * HALO_W + (t % 3)) : (uint32_t)(HALO_W + 1);
                umma2_f16(d, desc_noswz(ak + (uint32_t)(2 * ks) * (PAIR_PLANE / 16) + shift16, PAIR_PLANE / 16, HALO_W),
                          desc_noswz(w16 + (uint32_t)(2 * ks) * nh16 + (uint32_t)t * tap16, nh16, 8), p.idesc, (ks > 0 || t > 0) ? 1u : 0u);
              }
            }
          }
          if (MODE == ITG_CONV3X3 && p.ksteps2 > 0) {                   // folded 1x1 shortcut: centre tap of the second input's planes, weights of tap 9
            const uint32_t ak2 = ak + (uint32_t)(2 * p.ksteps) * (PAIR_PLANE / 16) + (uint32_t)(HALO_W + 1);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              if (ks < p.ksteps2)
                umma2_f16(d, desc_noswz(ak2 + (uint32_t)(2 * ks) * (PAIR_PLANE / 16), PAIR_PLANE / 16, HALO_W),
                          desc_noswz(w16 + (uint32_t)(2 * ks) * nh16 + 9u * tap16, nh16, 8), p.idesc, 1u);
            }
          }
          umma2_commit(bar_a_empty + 8 * s);
          umma2_commit(bar_acc_full + 8 * b);
        }
        __syncwarp();
        ITG_SACC(2, tl);
        if (++s == nring) { s = 0; sph ^= 1u; }
      }
      if (p.dbg && blockIdx.x == 0 && lane == 0) for (int i = 0; i < 3; ++i) p.dbg[i] = dacc[i];
    }
  } else if (warp >= SSM_WARP_CVT && warp < SSM_WARP_CVT + PAIR_LOADERS) {
    // ---- loaders (both CTAs, each for its own tile).  Chunk q = lt + 192 * i of a tile: halo row hy = q / (10 * np), then pixel hx and
    //      channel group c within the row's contiguous run.  All but one ring slot stay in flight (cp.async groups complete in order: the
    //      oldest tile is published when the window is full); everything that has landed is published before the warp sleeps. ----
    const int lt = (warp - SSM_WARP_CVT) * 32 + lane;                   // 0..191
    const T* in = reinterpret_cast<const T*>(p.in) + (size_t)p.in_cg_off * 8;
    const uint32_t a_smem = sbase + PAIR_OFF_A;
    // (a 1x1 conv reads no neighbours: only the 16 x 8 interior of the halo tile is fetched; the ring around it is never addressed)
    constexpr int ROWS = MODE == ITG_CONV3X3 ? HALO_H : TILE_H, COLS = MODE == ITG_CONV3X3 ? HALO_W : TILE_W, OFF = MODE == ITG_CONV3X3 ? 0 : 1;
    const int row_chunks = COLS * p.np, tile_chunks = ROWS * row_chunks;
    const int row_chunks2 = TILE_W * p.np2, tile_chunks2 = (MODE == ITG_CONV3X3 && p.in2 != nullptr) ? TILE_H * row_chunks2 : 0;
    // element offset from the halo origin | shared offset in 16-byte units, hy << 16, hx << 24, second input << 31 (hy = 31: none)
    uint32_t goff[PAIR_LD_ITERS], soff[PAIR_LD_ITERS];
#pragma unroll
    for (int i = 0; i < PAIR_LD_ITERS; ++i) {
      const int q = lt + i * PAIR_LOADERS * 32;
      if (q < tile_chunks) {
        const int ry = q / row_chunks, r = q - ry * row_chunks;
        const int rx = r / p.np, c = r - rx * p.np;
        const int hy = ry + OFF, hx = rx + OFF;
        goff[i] = (uint32_t)((hy * p.in_pitch + hx) * p.in_c + c * 8);
        soff[i] = (uint32_t)(c * (PAIR_PLANE / 16) + hy * HALO_W + hx) | ((uint32_t)hy << 16) | ((uint32_t)hx << 24);
      } else if (q < tile_chunks + tile_chunks2) {                      // the second input: interior of the tile only (a 1x1 conv)
        const int q2 = q - tile_chunks;
        const int ry = q2 / row_chunks2, r = q2 - ry * row_chunks2;
        const int rx = r / p.np2, c = r - rx * p.np2;
        const int hy = ry + 1, hx = rx + 1;
        goff[i] = (uint32_t)((hy * p.in_pitch + hx) * p.in2_c + c * 8);
        soff[i] = (uint32_t)((2 * p.ksteps + c) * (PAIR_PLANE / 16) + hy * HALO_W + hx) | ((uint32_t)hy << 16) | ((uint32_t)hx << 24) | (1u << 31);
      } else {
        goff[i] = 0;
        soff[i] = 31u << 16;
      }
    }
    const T* in2 = reinterpret_cast<const T*>(p.in2) + (size_t)p.in2_cg_off * 8;
    uint32_t s = 0, sph = 0;                                            // slot being filled, its phase
    uint32_t ps = 0;                                                    // oldest unpublished slot
    int unpub = 0;                                                      // committed, unpublished tiles
    unsigned long long dacc[3] = {0, 0, 0};
    long long tl = p.dbg ? clock64() : 0;
    auto publish = [&](int keep) {                                      // publish until at most `keep` tiles are unpublished
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
      const int hy_lim = tile_ok ? p.buf_h - y0 : 0, hx_lim = p.buf_w - x0;
      uint32_t ready = 0;
      if (lane == 0) ready = mbar_try_wait(bar_a_empty + 8 * s, sph ^ 1u) ? 1u : 0u;
      ready = __shfl_sync(0xffffffffu, ready, 0);
      if (!ready) {
        publish(0);
        ITG_SACC(2, tl);
        if (lane == 0) mbar_wait(bar_a_empty + 8 * s, sph ^ 1u);        // the MMAs that read this slot have completed
        __syncwarp();
      }
      ITG_SACC(0, tl);
      const uint32_t dst = a_smem + s * slot_bytes;
      const T* src0 = in + ((size_t)y0 * p.in_pitch + x0) * (size_t)p.in_c;
      const T* src2 = in2 + ((size_t)y0 * p.in_pitch + x0) * (size_t)p.in2_c;
      if (!(p.exp & 1)) {
#pragma unroll
        for (int i = 0; i < PAIR_LD_ITERS; ++i) {
          const int hy = (int)((soff[i] >> 16) & 31u), hx = (int)((soff[i] >> 24) & 31u);
          if (hy != 31) {
            const bool valid = hy < hy_lim && hx < hx_lim;
            const T* src = (soff[i] >> 31) ? src2 : src0;
            cp_async16_zfill(dst + ((soff[i] & 0xffffu) << 4), valid ? src + goff[i] : in, valid);
          }
        }
      }
      cp_async_commit();
      ++unpub;
      ITG_SACC(1, tl);
      publish(p.inflight - 1);
      ITG_SACC(2, tl);
      if (++s == nring) { s = 0; sph ^= 1u; }
    }
    publish(0);
    if (p.dbg && blockIdx.x == 0 && warp == SSM_WARP_CVT && lane == 0) for (int i = 0; i < 3; ++i) p.dbg[4 + i] = dacc[i];
  } else if (warp < 8) {
    // ---- epilogue (both CTAs): own tile; the two warps of a TMEM lane quarter split the pair's columns in halves.  Every global access
    //      is coalesced through a per-warp transposition buffer: a lane owns a pixel (a TMEM row) but loads / stores 16-byte units in
    //      memory order, 4 lanes per pixel and 8 consecutive pixels of a tile row per instruction -- 8 line requests instead of 32.
    //      (With one pixel per lane the epilogue's 1 024 line requests per 64-column tile took longer than the tile's MMAs.) ----
    constexpr bool G = (F & EF_GENERIC) != 0;
    const EpiParams& ep = p.ep;
    const bool has_res = G ? (ep.res_kind == ITG_RES_GRID) : ((F & EF_RES) != 0);
    const bool has_raw = G ? (ep.out_raw != nullptr) : ((F & EF_RAW) != 0);
    const bool has_act = G ? (ep.out_act != nullptr) : ((F & EF_ACT) != 0);
    const int eg = warp >> 2, q = warp & 3;
    // <= 32 columns per pair: a warp takes every other tile with all columns (one tile's epilogue is a chain of TMEM / shared / global
    // latencies longer than the tile's 18-36 MMAs); otherwise both warps work on every tile, half of the columns each
    const int c_lo = split_tiles ? 0 : eg * n_half;                     // first column of this warp, relative to n0
    const int ncols = split_tiles ? p.n_blk : n_half;
    int nch = (ep.out_c - (n0 + c_lo) + 7) >> 3;                        // 8-channel chunks this warp stores (columns beyond out_c are padding)
    nch = nch < 0 ? 0 : (nch > (ncols >> 3) ? (ncols >> 3) : nch);
    const int it0 = split_tiles ? eg : 0, it_step = split_tiles ? 2 : 1;
    const uint32_t stage = sbase + PAIR_OFF_STAGE + (uint32_t)warp * 2048u;
    const uint32_t vec = sbase + PAIR_OFF_VEC + (uint32_t)c_lo * 4u;
    // unit r of this lane in memory order: pixel pl = (lane + 32 r) / nch of the warp's 32, chunk k = (lane + 32 r) % nch
    int u_pl[4], u_k[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int u = lane + 32 * r;
      u_pl[r] = nch > 0 ? u / nch : 0;
      u_k[r] = nch > 0 ? u - u_pl[r] * nch : 0;
    }
    const uint32_t my_row = stage + (uint32_t)lane * 64u;
    const int my_swz = (lane >> 1) & 3;
    unsigned long long dacc[2] = {0, 0};
    long long tl = p.dbg ? clock64() : 0;
    // residual of one tile, in memory order (coalesced): fetched one tile ahead, the loads land while the previous tile is processed
    uint4 rn[4];
    auto load_res = [&](int pt_) {
      const int tile_ = 2 * pt_ + (int)rank;
      const int ty_ = tile_ / p.tiles_x, tx_ = tile_ - ty_ * p.tiles_x;
      const T* rp = reinterpret_cast<const T*>(ep.res);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (r < nch) {
          const int yy = ty_ * TILE_H + q * 4 + (u_pl[r] >> 3), xx = tx_ * TILE_W + (u_pl[r] & 7);
          rn[r] = make_uint4(0, 0, 0, 0);
          if (tile_ < p.ntiles && !(p.exp & 2) && yy < p.m_h && xx < p.m_w)
            rn[r] = *reinterpret_cast<const uint4*>(rp + grid_off(yy >> ep.res_shift, xx >> ep.res_shift, ep.res_w, ep.res_c, n0 + c_lo + 8 * u_k[r]));
        }
      }
    };
    if (has_res && it0 < n_my) load_res(slot + it0 * nslots);
    int pt = slot + it0 * nslots;
    for (int it = it0; it < n_my; it += it_step, pt += it_step * nslots) {
      const int b = it & (p.nbuf - 1);
      const int tile = 2 * pt + (int)rank;
      const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
      const int ybase = ty * TILE_H + q * 4, xbase = tx * TILE_W;     // this warp's 4 x 8 pixels
      const int y = ybase + (lane >> 3), x = xbase + (lane & 7);
      const bool tile_ok = tile < p.ntiles && !(p.exp & 2);
      const bool valid = tile_ok && (y < p.m_h) && (x < p.m_w);
      // residual: memory order -> transposition buffer -> this lane's pixel
      uint4 pre[4];
      if (has_res) {
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (r < nch) sts128(stage + (uint32_t)(u_pl[r] * 64 + ((u_k[r] ^ ((u_pl[r] >> 1) & 3)) << 4)), rn[r].x, rn[r].y, rn[r].z, rn[r].w);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < nch) asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(pre[k].x), "=r"(pre[k].y), "=r"(pre[k].z), "=r"(pre[k].w) : "r"(my_row + (uint32_t)((k ^ my_swz) << 4)));
        if (it + it_step < n_my) load_res(pt + it_step * nslots);
      }
      if (lane == 0) mbar_wait(bar_acc_full + 8 * b, ((uint32_t)it / (uint32_t)p.nbuf) & 1u);
      __syncwarp();
      ITG_SACC(0, tl);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * p.n_blk + c_lo);
      float v[32];
      {
        float lo[16], hi[16];
        tmem_ld16(trow, lo);
        if (nch > 2) tmem_ld16(trow + 16u, hi);
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i] = lo[i]; v[16 + i] = nch > 2 ? hi[i] : 0.f; }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar_acc_empty + 8 * b, 0u);   // the accumulator is in registers: the MMA warp may overwrite it
      // bias, residual
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < nch) {
          const float4 b0 = lds_f4(vec + (uint32_t)(k * 32)), b1 = lds_f4(vec + (uint32_t)(k * 32 + 16));
          v[8 * k] += b0.x; v[8 * k + 1] += b0.y; v[8 * k + 2] += b0.z; v[8 * k + 3] += b0.w;
          v[8 * k + 4] += b1.x; v[8 * k + 5] += b1.y; v[8 * k + 6] += b1.z; v[8 * k + 7] += b1.w;
          if (has_res) {
            const Vec8<T> t = *reinterpret_cast<const Vec8<T>*>(&pre[k]);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[8 * k + i] += Op<T>::to_f(t.v[i]);
          }
        }
      }
      // one pass per output tensor: pack -> transposition buffer -> coalesced stores (+ the frame pixels of edge pixels, directly)
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        if (pass == 0 ? !has_raw : !has_act) continue;
        T* out = reinterpret_cast<T*>(pass == 0 ? ep.out_raw : ep.out_act);
        const int border = pass == 0 ? (int)ITG_BORDER_NONE : ep.border;
        __syncwarp();                                                 // the buffer's previous readers are done
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < nch) {
            float w8[8];
            if (pass == 0) {
#pragma unroll
              for (int i = 0; i < 8; ++i) w8[i] = v[8 * k + i];
            } else {
              const float4 s0 = lds_f4(vec + 256u + (uint32_t)(k * 32)), s1 = lds_f4(vec + 256u + (uint32_t)(k * 32 + 16));
              const float4 t0 = lds_f4(vec + 512u + (uint32_t)(k * 32)), t1 = lds_f4(vec + 512u + (uint32_t)(k * 32 + 16));
              const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, sh[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float a = fmaf(sc[i], v[8 * k + i], sh[i]);
                w8[i] = ep.act_linear ? a : act_fn(a, ep.leak);
              }
            }
            sts128(my_row + (uint32_t)((k ^ my_swz) << 4), pack2<T>(w8[0], w8[1]), pack2<T>(w8[2], w8[3]), pack2<T>(w8[4], w8[5]), pack2<T>(w8[6], w8[7]));
          }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (r < nch) {
            const int yy = ybase + (u_pl[r] >> 3), xx = xbase + (u_pl[r] & 7);
            uint4 t;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w)
                         : "r"(stage + (uint32_t)(u_pl[r] * 64 + ((u_k[r] ^ ((u_pl[r] >> 1) & 3)) << 4))));
            if (tile_ok && yy < p.m_h && xx < p.m_w)
              *reinterpret_cast<uint4*>(out + grid_off(yy, xx, ep.out_w, ep.out_c, n0 + c_lo + 8 * u_k[r])) = t;
          }
        }
        if (border != ITG_BORDER_NONE && valid) {                    // frame pixels that mirror an edge pixel (F.pad of layers.py:82), or zeros
          const int h = ep.out_h, w = ep.out_w;
          const bool ey = (y == 0) | (y == h - 1), ex = (x == 0) | (x == w - 1);
          if (ey | ex) {
            const int fy = (y == 0) ? -1 : ((y == h - 1) ? h : y), fx = (x == 0) ? -1 : ((x == w - 1) ? w : x);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k < nch) {
                uint4 z = make_uint4(0, 0, 0, 0);
                if (border == ITG_BORDER_REPLICATE)
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(z.x), "=r"(z.y), "=r"(z.z), "=r"(z.w) : "r"(my_row + (uint32_t)((k ^ my_swz) << 4)));
                const int n = n0 + c_lo + 8 * k;
                if (ey) *reinterpret_cast<uint4*>(out + grid_off(fy, x, w, ep.out_c, n)) = z;
                if (ex) *reinterpret_cast<uint4*>(out + grid_off(y, fx, w, ep.out_c, n)) = z;
                if (ey && ex) *reinterpret_cast<uint4*>(out + grid_off(fy, fx, w, ep.out_c, n)) = z;
              }
            }
          }
        }
      }
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
