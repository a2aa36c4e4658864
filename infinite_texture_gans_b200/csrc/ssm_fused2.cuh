// StochasticSpatialModulation.forward (models/layers.py:228-234) on CTA PAIRS: the ssm_fused.cuh pipeline with
// tcgen05.mma.cta_group::2 (one instruction = M 256 x N <= 128 x K 16 across the two SMs of a cluster).
//
// Why pairs: with the embed weights parked in shared memory a single CTA can hold at most 64 GEMM columns, and an M = 128,
// N = 64 tcgen05.mma is bound by its operand reads from shared memory (4 KB of A + 2 KB of B at 128 B/cycle = 48 cycles for
// 32 cycles of math, measured with tools/umma_probe.cu).  In a CTA pair every SM still parks 64 columns, but the instruction
// covers 128 columns: each SM reads its own 4 KB of A and only ITS half of B, and the pair's tensor cores exchange the B halves
// -- 48 cycles of operand traffic for 64 cycles of math.  Per SM the kernel then runs at the tensor pipe's rate while it still
// loads nothing but the 1-channel noise map, x and its output.
//
// Layout of a pair: CTA rank r owns output tile 2 * pt + r of pair-tile pt (its own taps, m1 planes, accumulator rows in its own
// TMEM) and the weight rows of columns [r * N/2, (r + 1) * N/2) of the pair's column block.  Only the leader (rank 0) issues
// MMAs -- the mlp_shared GEMM (M 256 = both CTAs' halo pixels, N 128 = W1 split 64 / 64) and the embed conv -- and commits with
// a cluster multicast, so both CTAs' converters / epilogue warps / producers are released by the same instruction; in the
// other direction every role arrives on the LEADER's mbarriers (mapa + mbarrier.arrive.release.cluster).  Everything else
// (roles, barriers, phases) is as in ssm_fused.cuh.
#pragma once
#include "ssm_fused.cuh"

namespace itg {

constexpr int SSM2_NPAIR_MAX = 128;                                    // GEMM columns per CTA pair (64 resident per CTA)
constexpr int SSM2_HDR = 2048;                                         // barriers | bias (128 floats) | mean (64) | rstd (64)
constexpr int SSM2_VEC_BIAS = 512, SSM2_VEC_MEAN = 1024, SSM2_VEC_RSTD = 1280;
constexpr int SSM2_OFF_W1 = SSM2_HDR;                                  // this CTA's 64 rows of W1: [k-group 2][64][16 B]
constexpr int SSM2_OFF_TAPS = SSM2_OFF_W1 + 2 * 64 * 16;
constexpr int SSM2_OFF_WIN = SSM2_OFF_TAPS + 2 * SSM_TAPS_BYTES;
constexpr int SSM2_OFF_A = SSM2_OFF_WIN + 512;
constexpr int SSM2_OFF_W2 = SSM2_OFF_A + SSM_KG * PLANE_BYTES;          // [tap 9][k-group 16][n_half][16 B]
static_assert(SSM2_OFF_A % 128 == 0 && SSM2_OFF_W2 % 128 == 0, "operand alignment");

__host__ __device__ constexpr int ssm2_smem_bytes(int) { return SSM2_OFF_W2 + 9 * SSM_KG * SSM_NBLK_MAX * 16 + 1024; }     // weight image: fixed row pitch

__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16_pred(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// all MMAs issued so far by this thread arrive, when complete, on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma2_commit_pred(uint32_t bar, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %2;\n\t}"
      ::"r"(bar), "r"(leader), "h"((uint16_t)3)
      : "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster.  The arriving warp has already made its shared-memory writes
// visible to the async proxy (fence.proxy.async) / finished its TMEM reads (tcgen05.fence::before_thread_sync), and the data it guards
// never leaves its own SM: the default (CTA-scope) release is what CUTLASS's ClusterBarrier::arrive(cta_id) uses for the same hand-off.
// A cluster-scope release here cost ~800 cycles per arrival (measured: 965 vs 147 cycles per converted plane pair).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta)
      : "memory");
}
// (waits use mbar_wait: the default CTA-scope acquire on the local barrier, as CUTLASS's ClusterBarrier::wait does)

template <typename T>
__global__ void __launch_bounds__(SSM_THREADS, 1)
ssm_fused2_kernel(const SsmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;        // the dynamic window starts at the same offset in both CTAs
  uint8_t* const sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                              // 0 = leader
  const int n_half = p.n_blk >> 1;

  const uint32_t bar_taps_full = sbase;            // [2]  both producers -> leader's MMA warp            (count 2)
  const uint32_t bar_taps_empty = sbase + 16;      // [2]  MMA commit (multicast) -> producers
  const uint32_t bar_mlp_full = sbase + 32;        //      MMA commit (multicast) -> converters
  const uint32_t bar_mlp_empty = sbase + 40;       //      converters of both CTAs -> leader's MMA warp    (count 12)
  const uint32_t bar_a_full = sbase + 64;          // [8]  converters of both CTAs -> leader's MMA warp    (count 12)
  const uint32_t bar_a_empty = sbase + 128;        // [8]  MMA commit (multicast) -> converters
  const uint32_t bar_acc_full = sbase + 192;       // [2]  MMA commit (multicast) -> epilogue
  const uint32_t bar_acc_empty = sbase + 208;      // [2]  epilogue warps of both CTAs -> leader's MMA warp (count 16)
  const uint32_t tmem_slot = sbase + 224;

  const int pair = (int)blockIdx.x >> 1, npairs = (int)gridDim.x >> 1;
  const int nbp = pair % p.nblocks;                // the pair's block of GEMM columns
  const int slot = pair / p.nblocks, nslots = npairs / p.nblocks;
  const int npt = (p.ntiles + 1) >> 1;             // pair-tiles
  const int n_my = (slot < npt && slot < nslots) ? (npt - slot + nslots - 1) / nslots : 0;     // (pairs beyond nslots * nblocks idle)
  const int n0 = nbp * p.n_blk;                    // first GEMM column of the pair's block

  pdl_launch_dependents();
  if (warp == SSM_WARP_MMA && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_taps_full + 8 * i, 2);
      mbar_init(bar_taps_empty + 8 * i, 1);
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 16);
    }
    mbar_init(bar_mlp_full, 1);
    mbar_init(bar_mlp_empty, 12);
    for (int i = 0; i < SSM_GROUPS; ++i) {
      mbar_init(bar_a_full + 8 * i, 12);
      mbar_init(bar_a_empty + 8 * i, 1);
    }
    fence_barrier_init();
  }
  if (warp == SSM_WARP_PROD) tmem_alloc2(tmem_slot, 512);

  // ---- park this CTA's half of the weights: embed rows [n0 + rank * n_half, + n_half), W1 rows [rank * 64, + 64) ----
  {
    const T* w2 = reinterpret_cast<const T*>(p.w2);
    const int chunks = 9 * SSM_KG * n_half;
    const uint32_t w2s = sbase + SSM2_OFF_W2;
    for (int i = threadIdx.x; i < chunks; i += SSM_THREADS) {
      const int n = i % n_half, j = (i / n_half) % SSM_KG, t = i / (n_half * SSM_KG);
      const int ng = n0 + (int)rank * n_half + n;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (ng < p.n_pad) v = *reinterpret_cast<const uint4*>(w2 + ((size_t)t * p.n_pad + ng) * SSM_K + j * 8);
      sts128(w2s + (uint32_t)(((t * SSM_KG + j) * SSM_NBLK_MAX + n) * 16), v.x, v.y, v.z, v.w);     // row pitch fixed at 64: compile-time descriptors
    }
    const T* w1 = reinterpret_cast<const T*>(p.w1);
    for (int i = threadIdx.x; i < 2 * 64; i += SSM_THREADS) {
      const int n = i % 64, j = i / 64;
      const uint4 v = *reinterpret_cast<const uint4*>(w1 + ((int)rank * 64 + n) * 16 + j * 8);
      sts128(sbase + SSM2_OFF_W1 + (uint32_t)i * 16u, v.x, v.y, v.z, v.w);
    }
    for (int i = threadIdx.x; i < 2 * SSM_TAPS_BYTES / 16; i += SSM_THREADS)
      sts128(sbase + SSM2_OFF_TAPS + (uint32_t)i * 16u, 0u, 0u, 0u, 0u);
    fence_proxy_async();
    float* const vec = reinterpret_cast<float*>(sptr);
    if (threadIdx.x < SSM2_NPAIR_MAX) {
      const int n = n0 + (int)threadIdx.x;
      vec[SSM2_VEC_BIAS / 4 + threadIdx.x] = ((int)threadIdx.x < p.n_blk && n < p.n_pad) ? p.ep.bias[n] : 0.f;
    } else if (threadIdx.x < SSM2_NPAIR_MAX + SSM2_NPAIR_MAX / 2) {
      const int i = (int)threadIdx.x - SSM2_NPAIR_MAX, ch = (n0 >> 1) + i;
      const bool ok = 2 * i < p.n_blk && ch < p.ep.out_c;
      vec[SSM2_VEC_MEAN / 4 + i] = ok ? p.ep.mod_mean[ch] : 0.f;
      vec[SSM2_VEC_RSTD / 4 + i] = ok ? p.ep.mod_rstd[ch] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // the peer's barriers are initialised and its weights parked before anybody signals / issues
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  if (warp == SSM_WARP_PROD) {
    // ---- taps producer (both CTAs, each for its own tile) ----
    T* const win = reinterpret_cast<T*>(sptr + SSM2_OFF_WIN);
    const int map_h = p.h + 4, map_w = p.w + 4;
    const T one_t = Op<T>::from_f(1.f);
    const uint32_t one = (uint32_t)(*reinterpret_cast<const unsigned short*>(&one_t));
    int pt = slot;
    for (int it = 0; it < n_my; ++it, pt += nslots) {
      const int tb = it & 1;
      const int tile = 2 * pt + (int)rank;          // tile >= ntiles (odd tile count): every map read falls outside and yields zeros
      if (lane == 0) mbar_wait(bar_taps_empty + 8 * tb, (((uint32_t)it >> 1) & 1u) ^ 1u);
      __syncwarp();
      const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
      const int y0 = ty * TILE_H, x0 = tx * TILE_W;
#pragma unroll
      for (int k = 0; k < (SSM_WIN_N + 31) / 32; ++k) {
        const int idx = lane + 32 * k;
        if (idx < SSM_WIN_N) {
          const int r = idx / SSM_WIN_W, c = idx - r * SSM_WIN_W;
          const int yy = y0 + r, xx = x0 + c;
          const float v = (yy < map_h && xx < map_w) ? p.map[(size_t)yy * p.map_pitch + xx] : 0.f;
          win[idx] = Op<T>::from_f(v);
        }
      }
      __syncwarp();
      const uint32_t dst = sbase + SSM2_OFF_TAPS + (uint32_t)tb * SSM_TAPS_BYTES;
#pragma unroll
      for (int k = 0; k < (HALO_PX + 31) / 32; ++k) {
        const int hp = lane + 32 * k;
        if (hp < HALO_PX) {
          const int hy = hp / HALO_W, hx = hp - hy * HALO_W;
          const unsigned short* wp = reinterpret_cast<const unsigned short*>(win) + hy * SSM_WIN_W + hx;
          uint32_t t[9];
#pragma unroll
          for (int q = 0; q < 9; ++q) t[q] = wp[(q / 3) * SSM_WIN_W + (q % 3)];
          sts128(dst + (uint32_t)hp * 16u, t[0] | (t[1] << 16), t[2] | (t[3] << 16), t[4] | (t[5] << 16), t[6] | (t[7] << 16));
          sts128(dst + (uint32_t)(SSM_TAPS_ROWS * 16 + hp * 16), t[8] | (one << 16), one, 0u, 0u);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar_taps_full + 8 * tb, 0u);
    }
  } else if (warp == SSM_WARP_MMA) {
    if (rank == 0) {
      // ---- MMA warp of the LEADER: issues for both CTAs (elect-guarded blocks, descriptors on the uniform datapath: see ssm_fused.cuh) ----
      const uint32_t w1_16 = (sbase + SSM2_OFF_W1) >> 4, w2_16 = (sbase + SSM2_OFF_W2) >> 4, a16 = (sbase + SSM2_OFF_A) >> 4;
      constexpr uint32_t nh16 = SSM_NBLK_MAX;                          // row pitch of the parked weight image (the MMA's N is in idesc)
      unsigned long long dacc[4] = {0, 0, 0, 0};
      long long tl = p.dbg ? clock64() : 0;
      auto issue_mlp = [&](int it) {
        const int tb = it & 1;
        if (lane == 0) {
          mbar_wait(bar_taps_full + 8 * tb, ((uint32_t)it >> 1) & 1u);
          mbar_wait(bar_mlp_empty, ((uint32_t)it & 1u) ^ 1u);
        }
        __syncwarp();
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t t16 = (sbase + SSM2_OFF_TAPS + (uint32_t)tb * SSM_TAPS_BYTES) >> 4;
          const uint64_t bdesc = desc_noswz(w1_16, 64, 8);
          umma2_f16(tmem_base + SSM_TMEM_MLP, desc_noswz(t16, SSM_TAPS_ROWS, 8), bdesc, p.idesc_mlp, 0u);
          umma2_f16(tmem_base + SSM_TMEM_MLP + SSM_K, desc_noswz(t16 + 128, SSM_TAPS_ROWS, 8), bdesc, p.idesc_mlp, 0u);
          umma2_commit(bar_taps_empty + 8 * tb);
          umma2_commit(bar_mlp_full);
        }
        __syncwarp();
      };
      if (n_my > 0) issue_mlp(0);
      ITG_SACC(0, tl);
      for (int it = 0; it < n_my; ++it) {
        const int b = it & 1;
        if (lane == 0) mbar_wait(bar_acc_empty + 8 * b, (((uint32_t)it >> 1) & 1u) ^ 1u);
        __syncwarp();
        ITG_SACC(1, tl);
        const uint32_t d = tmem_base + (uint32_t)(b * p.n_blk);
#pragma unroll 1
        for (int g = 0; g < SSM_GROUPS; ++g) {
          if (lane == 0) mbar_wait(bar_a_full + 8 * g, (uint32_t)it & 1u);
          __syncwarp();
          ITG_SACC(2, tl);
          tc_fence_after();
          if (elect_one_sync()) {
#pragma unroll
            for (int k2 = 0; k2 < 2; ++k2) {
              const int ks = 2 * g + k2;
              const uint32_t ak = a16 + (uint32_t)(2 * ks) * (PLANE_BYTES / 16);
              const uint32_t wk = w2_16 + (uint32_t)(2 * ks) * nh16;
#pragma unroll
              for (int t = 0; t < 9; ++t) {
                if ((p.exp & 8) && t > 0) break;
                const uint32_t shift16 = (uint32_t)((t / 3) * HALO_W + (t % 3));
                umma2_f16(d, desc_noswz(ak + shift16, PLANE_BYTES / 16, HALO_W), desc_noswz(wk + (uint32_t)(t * SSM_KG) * nh16, nh16, 8),
                          p.idesc_emb, (ks > 0 || t > 0) ? 1u : 0u);
              }
            }
            umma2_commit(bar_a_empty + 8 * g);
            if (g == SSM_GROUPS - 1) umma2_commit(bar_acc_full + 8 * b);
          }
          __syncwarp();
          ITG_SACC(3, tl);
          if (g == 0 && it + 1 < n_my) { issue_mlp(it + 1); ITG_SACC(0, tl); }
        }
      }
      if (p.dbg && blockIdx.x == 0 && lane == 0) for (int i = 0; i < 4; ++i) p.dbg[i] = dacc[i];
    }
  } else if (warp >= SSM_WARP_CVT) {
    // ---- converters (both CTAs): own rows of the m1 accumulator -> ReLU -> operand type -> own A planes ----
    const int rb = (warp - SSM_WARP_CVT) >> 2, q = warp & 3;
    const int hp = rb * 128 + q * 32 + lane;
    const bool hp_ok = hp < HALO_PX;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(SSM_TMEM_MLP + rb * SSM_K);
    const uint32_t dst = sbase + SSM2_OFF_A + (uint32_t)hp * 16u;
    unsigned long long dacc[3] = {0, 0, 0};
    long long tl = p.dbg ? clock64() : 0;
    const int hy = hp / HALO_W, hx = hp - hy * HALO_W;
    int pt = slot;
    for (int it = 0; it < n_my; ++it, pt += nslots) {
      bool ring = false;                              // non-local Generator: the hidden map is zero outside the image (see ssm_fused.cuh)
      if (p.zero_ring) {
        const int tile = 2 * pt + (int)rank;
        const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
        const int my = ty * TILE_H + hy, mx = tx * TILE_W + hx;
        ring = my == 0 || mx == 0 || my >= p.h + 1 || mx >= p.w + 1;
      }
      if (lane == 0) mbar_wait(bar_mlp_full, (uint32_t)it & 1u);
      __syncwarp();
      ITG_SACC(0, tl);
      tc_fence_after();
      // groups of 32 channels (two k-steps, four planes): one TMEM load, one proxy fence and one (remote) arrival per group; the next
      // group's load is in flight while this one is packed and stored
      uint32_t r[32];
      tmem_ld32_issue(trow, r);
      tmem_ld_wait(r);
#pragma unroll 1
      for (int g = 0; g < SSM_GROUPS; ++g) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = ring ? 0u : pack2<T>(fmaxf(__uint_as_float(r[2 * i]), 0.f), fmaxf(__uint_as_float(r[2 * i + 1]), 0.f));
        if (g + 1 < SSM_GROUPS) tmem_ld32_issue(trow + (uint32_t)(32 * (g + 1)), r);
        if (lane == 0) mbar_wait(bar_a_empty + 8 * g, ((uint32_t)it & 1u) ^ 1u);
        __syncwarp();
        ITG_SACC(1, tl);
        if (hp_ok && !(p.exp & 2)) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            sts128(dst + (uint32_t)((4 * g + q4) * PLANE_BYTES), w[4 * q4], w[4 * q4 + 1], w[4 * q4 + 2], w[4 * q4 + 3]);
        }
        if (!(p.exp & 1)) fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(bar_a_full + 8 * g, 0u);
        if (g + 1 < SSM_GROUPS) tmem_ld_wait(r);
        if (g + 2 == SSM_GROUPS) {                                     // the last group is in registers: the m1 accumulator may be overwritten
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(bar_mlp_empty, 0u);
        }
        ITG_SACC(2, tl);
      }
    }
    if (p.dbg && blockIdx.x == 0 && warp == SSM_WARP_CVT && lane == 0) for (int i = 0; i < 3; ++i) p.dbg[4 + i] = dacc[i];
  } else {
    // ---- epilogue (both CTAs): own tile x all columns of the pair's block; group g takes the 16-column chunks c = g, g + 2, g + 4, g + 6 ----
    const int eg = warp >> 2, q = warp & 3;
    const int row = q * 32 + lane;
    const EpiParams& ep = p.ep;
    const uint32_t vb = sbase + SSM2_VEC_BIAS, vm = sbase + SSM2_VEC_MEAN, vr = sbase + SSM2_VEC_RSTD;
    unsigned long long dacc[2] = {0, 0};
    long long tl = p.dbg ? clock64() : 0;
    int pt = slot;
    for (int it = 0; it < n_my; ++it, pt += nslots) {
      const int b = it & 1;
      const int tile = 2 * pt + (int)rank;
      const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
      const int y = ty * TILE_H + (row >> 3), x = tx * TILE_W + (row & 7);
      const bool valid = tile < p.ntiles && (y < p.h) && (x < p.w);
      bool live[4], ok[4];
      uint4 xr[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = eg + 2 * j, n = n0 + 16 * c;
        live[j] = 16 * c < p.n_blk && n < p.n_pad;
        ok[j] = live[j] && valid && (n >> 1) < ep.out_c;
        xr[j] = make_uint4(0, 0, 0, 0);
        if (ok[j])
          xr[j] = *reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(ep.mod_x) +
                                                  grid_off(y >> ep.mod_shift, x >> ep.mod_shift, ep.mod_w, ep.mod_c, n >> 1));
      }
      if (lane == 0) mbar_wait(bar_acc_full + 8 * b, ((uint32_t)it >> 1) & 1u);
      __syncwarp();
      ITG_SACC(0, tl);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * p.n_blk);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = eg + 2 * j;
        if (!live[j]) continue;
        float v[16];
        tmem_ld16(trow + (uint32_t)(16 * c), v);
        if (!ok[j] || (p.exp & 4)) continue;
        float xf[8], yv[8];
        {
          const Vec8<T> t8 = *reinterpret_cast<const Vec8<T>*>(&xr[j]);
#pragma unroll
          for (int i = 0; i < 8; ++i) xf[i] = Op<T>::to_f(t8.v[i]);
        }
        const float4 m0 = lds_f4(vm + (uint32_t)(32 * c)), m1 = lds_f4(vm + (uint32_t)(32 * c + 16));
        const float4 r0 = lds_f4(vr + (uint32_t)(32 * c)), r1 = lds_f4(vr + (uint32_t)(32 * c + 16));
        const float mean[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
        const float rstd[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float4 bb = lds_f4(vb + (uint32_t)(64 * c + 16 * h));
          const float g0 = v[4 * h] + bb.x, b0 = v[4 * h + 1] + bb.y, g1 = v[4 * h + 2] + bb.z, b1 = v[4 * h + 3] + bb.w;
          const float u0 = (1.f + g0) * ((xf[2 * h] - mean[2 * h]) * rstd[2 * h]) + b0;
          const float u1 = (1.f + g1) * ((xf[2 * h + 1] - mean[2 * h + 1]) * rstd[2 * h + 1]) + b1;
          yv[2 * h] = ep.act_linear ? u0 : act_fn(u0, ep.leak);
          yv[2 * h + 1] = ep.act_linear ? u1 : act_fn(u1, ep.leak);
        }
        store8_framed(reinterpret_cast<T*>(ep.out_act), y, x, ep.out_h, ep.out_w, ep.out_c, (n0 + 16 * c) >> 1, yv, ep.border);
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
  if (warp == SSM_WARP_PROD) tmem_dealloc2(tmem_base, 512);
}

}  // namespace itg
