// StochasticSpatialModulation.forward (models/layers.py:228-234) as ONE persistent tcgen05 kernel:
//
//     m1          = relu(conv3x3_valid(map) + b1)          mlp_shared: 1 -> 128 channels
//     [gamma|beta] = conv3x3_valid(m1) + b2                 embed: 128 -> 2C channels   (93 % of the SSM config's FLOPs)
//     out         = [act]((1 + gamma) * bn0(x) + beta)      modulation (+ activation, + outer padding of the frame)
//
// The 128-channel hidden map m1 (256 B per pixel) never exists in global memory.  Per 16 x 8 output tile:
//   1. a producer warp reads the tile's (16+4) x (8+4) window of the fp32 noise map and writes the nine 3x3 taps of each of
//      the (16+2) x (8+2) halo pixels as one K = 16 operand row (taps 0..8, two constant ones that carry the bias, zeros);
//   2. the MMA warp runs  m1_acc[halo pixel][128] = taps[halo pixel][16] x W1[128][16]^T  (two M = 128 tcgen05.mma, the bias
//      b1 rides in the two constant K slots as an fp16 hi + lo pair) into tensor memory;
//   3. six converter warps drain that accumulator 16 channels at a time (tcgen05.ld), apply ReLU, round to the operand type
//      and store the values as the embed conv's A operand: 16 planes of [halo pixel][8 channels] in shared memory (the
//      halo-tile layout of conv_tile.cuh), released plane pair by plane pair through mbarriers;
//   4. the MMA warp accumulates the embed conv over 8 k-steps x 9 taps: the tap shift is a 16-byte-granular offset of the
//      no-swizzle K-major A descriptor (SBO = one halo-tile row, LBO = one plane), so the tile's m1 values are produced
//      once and read nine times from shared memory.  The embed weights of the CTA's <= 64 GEMM columns (all taps, K = 128:
//      <= 144 KB) are parked in shared memory for the whole launch: in steady state the kernel loads nothing but the
//      1-channel map, x and its own output.  Wider layers (N = 2C > 64) are split into N blocks over the CTAs
//      (CTA c serves block c % nblocks; m1 is recomputed per block, which costs two small MMAs and one conversion);
//   5. eight epilogue warps drain the embed accumulator (double-buffered in TMEM) through the modulation epilogue while the
//      next tile is being computed; x is fetched before they sleep on the MMA barrier, the per-channel vectors sit in shared memory.
// The plane-pair granularity of step 3/4 lets the conversion of tile i+1 trail the embed MMAs of tile i by one k-step, so
// one 45 KB A buffer suffices next to the parked weights.  Every mbarrier has one producing and one consuming role that
// walk its phases in order (see conv_tile.cuh for why that matters).
#pragma once
#include "conv_tile.cuh"

namespace itg {

constexpr int SSM_K = 128;                                   // nhidden of StochasticSpatialModulation (models/layers.py:220)
constexpr int SSM_KG = SSM_K / 8;                            // 8-channel planes of the A operand
constexpr int SSM_KSTEPS = SSM_K / 16;
// Warp roles.  A warp's scheduler is warp % 4 and so is the TMEM lane quarter it may read; within a scheduler the hardware favours the
// HIGHER warp id (measured: with the epilogue warps on top, their barrier polling starved the MMA warp and the converters -- 25 % slower).
// So the roles that feed the tensor pipe sit on top: 15 MMA issuer, 14 taps producer (+ TMEM allocator), 12-13 converters of halo pixels
// 128..179, 8-11 converters of halo pixels 0..127, and the two epilogue groups 0-3 / 4-7 at the bottom.
constexpr int SSM_WARPS = 16;
constexpr int SSM_WARP_MMA = 15, SSM_WARP_PROD = 14, SSM_WARP_CVT = 8;
constexpr int SSM_THREADS = 32 * SSM_WARPS;
constexpr int SSM_NBLK_MAX = 64;                             // GEMM columns whose weights fit shared memory next to the A planes
constexpr int SSM_TAPS_ROWS = 256;                           // 180 halo pixels padded to two M = 128 row blocks
constexpr int SSM_TAPS_BYTES = 2 * SSM_TAPS_ROWS * 16;       // two 8-element K groups
constexpr int SSM_W1_BYTES = 2 * SSM_K * 16;
constexpr int SSM_WIN_W = TILE_W + 4, SSM_WIN_H = TILE_H + 4;   // map window of one tile: 12 x 20
constexpr int SSM_WIN_N = SSM_WIN_W * SSM_WIN_H;
constexpr int SSM_HDR = 1024;                                 // barriers | embed bias (64 floats) | mean (32) | rstd (32) of this CTA's column block
constexpr int SSM_VEC_BIAS = 256, SSM_VEC_MEAN = 512, SSM_VEC_RSTD = 640;
constexpr int SSM_OFF_W1 = SSM_HDR;
constexpr int SSM_OFF_TAPS = SSM_OFF_W1 + SSM_W1_BYTES;                // 2 buffers
constexpr int SSM_OFF_WIN = SSM_OFF_TAPS + 2 * SSM_TAPS_BYTES;         // 16-bit staging of the map window
constexpr int SSM_OFF_A = SSM_OFF_WIN + 512;                           // 16 planes of 180 halo pixels x 16 B
constexpr int SSM_OFF_W2 = SSM_OFF_A + SSM_KG * PLANE_BYTES;           // [tap 9][k-group 16][64 rows, n_blk used][16 B]
constexpr int SSM_TMEM_MLP = 256;                                      // TMEM columns 256..511: the two m1 row blocks; 0..2*n_blk: embed accumulators
static_assert(SSM_WIN_N * 2 <= 512, "map window staging");
static_assert(SSM_OFF_A % 128 == 0 && SSM_OFF_W2 % 128 == 0, "operand alignment");

__host__ __device__ constexpr int ssm_smem_bytes(int) { return SSM_OFF_W2 + 9 * SSM_KG * SSM_NBLK_MAX * 16 + 1024; }      // weight image: fixed row pitch

struct SsmParams {
  int h, w;                 // interior size of the modulated tensor / output
  int tiles_x, ntiles;
  const float* map;         // fp32 (h+4) x (w+4) noise map of this level (utils.py:246), `map_pitch` floats per row
  int map_pitch;
  const void* w1;           // [128][16] operand dtype: mlp_shared taps 0..8, bias hi / lo in columns 9 / 10
  const void* w2;           // [9][n_pad][128] operand dtype (gamma / beta rows interleaved, packing.pack_ssm_embed)
  int n_pad, n_blk, nblocks;
  uint32_t idesc_mlp, idesc_emb;
  unsigned long long* dbg;  // optional cycle counters of CTA 0 (ITG_TILE_DBG=1)
  int zero_ring;            // 1: hidden map forced to zero on the ring outside the image (non-local Generator, zero-padded convs)
  int exp;                  // developer experiments (ITG_SSM_EXP bit mask: WRONG RESULTS, timing only): 1 no proxy fence, 2 no plane stores,
                            // 4 no epilogue math / stores, 8 one tap per k-step
  EpiParams ep;
};

// per-role cycle counters: compiled in only with -DITG_SSM_DBG (they cost code size in every hot loop)
#ifdef ITG_SSM_DBG
#define ITG_SACC(slot, tvar) do { if (p.dbg) { const long long now_ = clock64(); dacc[slot] += (unsigned long long)(now_ - tvar); tvar = now_; } } while (0)
#else
#define ITG_SACC(slot, tvar) do { (void)tvar; } while (0)
#endif

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// 32 consecutive TMEM columns of this thread's lane, without waiting: the registers may only be read after tmem_ld_wait(r)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// ... the wait names the registers as in/out operands so that no use of them can be scheduled before it
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]),
                 "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]),
                 "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}
constexpr int SSM_GROUPS = SSM_KSTEPS / 2;                   // the converters hand the A planes over in groups of two k-steps (32 channels)

template <typename T>
__global__ void __launch_bounds__(SSM_THREADS, 1)
ssm_fused_kernel(const SsmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t bar_taps_full = sbase;            // [2]  producer (32 lanes) -> MMA
  const uint32_t bar_taps_empty = sbase + 16;      // [2]  MMA commit -> producer
  const uint32_t bar_mlp_full = sbase + 32;        //      MMA commit -> converters
  const uint32_t bar_mlp_empty = sbase + 40;       //      converters (6 warps) -> MMA
  const uint32_t bar_a_full = sbase + 64;          // [8]  converters (6 warps) -> MMA, one per plane pair
  const uint32_t bar_a_empty = sbase + 128;        // [8]  MMA commit -> converters
  const uint32_t bar_acc_full = sbase + 192;       // [2]  MMA commit -> epilogue
  const uint32_t bar_acc_empty = sbase + 208;      // [2]  epilogue (8 warps) -> MMA
  const uint32_t tmem_slot = sbase + 224;

  const int nb = (int)blockIdx.x % p.nblocks;      // this CTA's block of GEMM columns (weights resident)
  const int slot = (int)blockIdx.x / p.nblocks, nslots = (int)gridDim.x / p.nblocks;
  const int n_my = slot < p.ntiles ? (p.ntiles - slot + nslots - 1) / nslots : 0;

  pdl_launch_dependents();
  if (warp == SSM_WARP_MMA && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_taps_full + 8 * i, 32);
      mbar_init(bar_taps_empty + 8 * i, 1);
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 8);
    }
    mbar_init(bar_mlp_full, 1);
    mbar_init(bar_mlp_empty, 6);
    for (int i = 0; i < SSM_GROUPS; ++i) {
      mbar_init(bar_a_full + 8 * i, 6);
      mbar_init(bar_a_empty + 8 * i, 1);
    }
    fence_barrier_init();
  }
  if (warp == SSM_WARP_PROD) tmem_alloc(tmem_slot, 512);

  // ---- park the weights (they do not depend on the previous launch): embed block [tap][k-group][n][8 ch], W1 [k-group][n][8 ch] ----
  {
    const T* w2 = reinterpret_cast<const T*>(p.w2);
    const int chunks = 9 * SSM_KG * p.n_blk;
    const uint32_t w2s = sbase + SSM_OFF_W2;
    for (int i = threadIdx.x; i < chunks; i += SSM_THREADS) {
      const int n = i % p.n_blk, j = (i / p.n_blk) % SSM_KG, t = i / (p.n_blk * SSM_KG);
      const int ng = nb * p.n_blk + n;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (ng < p.n_pad) v = *reinterpret_cast<const uint4*>(w2 + ((size_t)t * p.n_pad + ng) * SSM_K + j * 8);
      sts128(w2s + (uint32_t)(((t * SSM_KG + j) * SSM_NBLK_MAX + n) * 16), v.x, v.y, v.z, v.w);     // row pitch fixed at 64: compile-time descriptors
    }
    const T* w1 = reinterpret_cast<const T*>(p.w1);
    for (int i = threadIdx.x; i < 2 * SSM_K; i += SSM_THREADS) {
      const int n = i % SSM_K, j = i / SSM_K;
      const uint4 v = *reinterpret_cast<const uint4*>(w1 + n * 16 + j * 8);
      sts128(sbase + SSM_OFF_W1 + (uint32_t)i * 16u, v.x, v.y, v.z, v.w);
    }
    for (int i = threadIdx.x; i < 2 * SSM_TAPS_BYTES / 16; i += SSM_THREADS)       // rows 180..255 of the taps operand stay zero
      sts128(sbase + SSM_OFF_TAPS + (uint32_t)i * 16u, 0u, 0u, 0u, 0u);
    fence_proxy_async();                                                           // generic writes -> async proxy (UMMA) reads
    float* const vec = reinterpret_cast<float*>(sptr);
    if (threadIdx.x < SSM_NBLK_MAX) {
      const int n = nb * p.n_blk + (int)threadIdx.x;
      vec[SSM_VEC_BIAS / 4 + threadIdx.x] = ((int)threadIdx.x < p.n_blk && n < p.n_pad) ? p.ep.bias[n] : 0.f;
    } else if (threadIdx.x < SSM_NBLK_MAX + SSM_NBLK_MAX / 2) {
      const int i = (int)threadIdx.x - SSM_NBLK_MAX, ch = ((nb * p.n_blk) >> 1) + i;
      const bool ok = 2 * i < p.n_blk && ch < p.ep.out_c;
      vec[SSM_VEC_MEAN / 4 + i] = ok ? p.ep.mod_mean[ch] : 0.f;
      vec[SSM_VEC_RSTD / 4 + i] = ok ? p.ep.mod_rstd[ch] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();                                       // x (and, in a captured step, the map) of the previous launches are complete from here on

  if (warp == SSM_WARP_PROD) {
    // ---- taps producer: map window -> K = 16 operand rows of the tile's 180 halo pixels ----
    T* const win = reinterpret_cast<T*>(sptr + SSM_OFF_WIN);
    const int map_h = p.h + 4, map_w = p.w + 4;
    const T one_t = Op<T>::from_f(1.f);
    const uint32_t one = (uint32_t)(*reinterpret_cast<const unsigned short*>(&one_t));     // the constant K slots that carry the bias
    int tile = slot;
    for (int it = 0; it < n_my; ++it, tile += nslots) {
      const int tb = it & 1;
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
      const uint32_t dst = sbase + SSM_OFF_TAPS + (uint32_t)tb * SSM_TAPS_BYTES;
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
      mbar_arrive(bar_taps_full + 8 * tb);
    }
  } else if (warp == SSM_WARP_MMA) {
    // ---- MMA warp.  Waits run warp-uniform (lane 0 polls, __syncwarp releases the warp); the MMAs of one k-step are issued inside an
    //      `if (elect_one_sync())` block with every descriptor a compile-time offset from warp-uniform values: ptxas then keeps the
    //      whole issue sequence on the uniform datapath (UTCHMMA back to back, 2-4 uniform ALU instructions in between).  A per-lane
    //      predicate on the instruction instead makes it build each descriptor in vector registers and move it over with an
    //      ELECT / R2UR / BRA.U.ANY loop: measured 88 cycles per MMA, more than the MMA itself (profiles/r02_notes.md). ----
    const uint32_t w1_16 = (sbase + SSM_OFF_W1) >> 4, w2_16 = (sbase + SSM_OFF_W2) >> 4, a16 = (sbase + SSM_OFF_A) >> 4;
    constexpr uint32_t n16 = SSM_NBLK_MAX;                             // row pitch of the parked weight image (the MMA's N is in idesc)
    unsigned long long dacc[4] = {0, 0, 0, 0};
    long long tl = p.dbg ? clock64() : 0;
    auto issue_mlp = [&](int it) {
      const int tb = it & 1;
      if (lane == 0) {
        mbar_wait(bar_taps_full + 8 * tb, ((uint32_t)it >> 1) & 1u);
        mbar_wait(bar_mlp_empty, ((uint32_t)it & 1u) ^ 1u);           // the converters have drained the previous tile's m1 accumulator
      }
      __syncwarp();
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t t16 = (sbase + SSM_OFF_TAPS + (uint32_t)tb * SSM_TAPS_BYTES) >> 4;
        const uint64_t bdesc = desc_noswz(w1_16, SSM_K, 8);
        umma_f16(tmem_base + SSM_TMEM_MLP, desc_noswz(t16, SSM_TAPS_ROWS, 8), bdesc, p.idesc_mlp, 0u);
        umma_f16(tmem_base + SSM_TMEM_MLP + SSM_K, desc_noswz(t16 + 128, SSM_TAPS_ROWS, 8), bdesc, p.idesc_mlp, 0u);
        umma_commit(bar_taps_empty + 8 * tb);
        umma_commit(bar_mlp_full);
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
            const uint32_t wk = w2_16 + (uint32_t)(2 * ks) * n16;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              if ((p.exp & 8) && t > 0) break;
              const uint32_t shift16 = (uint32_t)((t / 3) * HALO_W + (t % 3));
              umma_f16(d, desc_noswz(ak + shift16, PLANE_BYTES / 16, HALO_W), desc_noswz(wk + (uint32_t)(t * SSM_KG) * n16, n16, 8),
                       p.idesc_emb, (ks > 0 || t > 0) ? 1u : 0u);
            }
          }
          umma_commit(bar_a_empty + 8 * g);                           // these four planes may be overwritten with the next tile's values
          if (g == SSM_GROUPS - 1) umma_commit(bar_acc_full + 8 * b);
        }
        __syncwarp();
        ITG_SACC(3, tl);
        if (g == 0 && it + 1 < n_my) { issue_mlp(it + 1); ITG_SACC(0, tl); }     // 18 embed MMAs are queued while this waits for the converters
      }
    }
    if (p.dbg && blockIdx.x == 0 && lane == 0) for (int i = 0; i < 4; ++i) p.dbg[i] = dacc[i];
  } else if (warp >= SSM_WARP_CVT) {
    // ---- converters: m1 accumulator (TMEM) -> ReLU -> operand type -> A planes.  Warps 4-7: halo pixels 0..127, warps 8-9: 128..179 ----
    const int rb = (warp - SSM_WARP_CVT) >> 2, q = warp & 3;
    const int hp = rb * 128 + q * 32 + lane;
    const bool hp_ok = hp < HALO_PX;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(SSM_TMEM_MLP + rb * SSM_K);
    const uint32_t dst = sbase + SSM_OFF_A + (uint32_t)hp * 16u;
    unsigned long long dacc[3] = {0, 0, 0};
    long long tl = p.dbg ? clock64() : 0;
    const int hy = hp / HALO_W, hx = hp - hy * HALO_W;
    int tile = slot;
    for (int it = 0; it < n_my; ++it, tile += nslots) {
      // zero padding of the non-local Generator (--padding_mode zeros: conv3x3(..., p=1) of layers.py:213-224): the hidden map is zero
      // OUTSIDE the image, i.e. on the outermost ring of the (h+2) x (w+2) valid-conv grid, instead of relu(conv(zero-extended map))
      bool ring = false;
      if (p.zero_ring) {
        const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
        const int my = ty * TILE_H + hy, mx = tx * TILE_W + hx;
        ring = my == 0 || mx == 0 || my >= p.h + 1 || mx >= p.w + 1;
      }
      if (lane == 0) mbar_wait(bar_mlp_full, (uint32_t)it & 1u);
      __syncwarp();
      ITG_SACC(0, tl);
      tc_fence_after();
      // groups of 32 channels (two k-steps, four planes): one TMEM load, one proxy fence and one arrival per group; the next group's
      // load is in flight while this one is packed and stored (a converter warp working chunk by chunk was the critical path)
      uint32_t r[32];
      tmem_ld32_issue(trow, r);
      tmem_ld_wait(r);
#pragma unroll 1
      for (int g = 0; g < SSM_GROUPS; ++g) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = ring ? 0u : pack2<T>(fmaxf(__uint_as_float(r[2 * i]), 0.f), fmaxf(__uint_as_float(r[2 * i + 1]), 0.f));
        if (g + 1 < SSM_GROUPS) tmem_ld32_issue(trow + (uint32_t)(32 * (g + 1)), r);
        if (lane == 0) mbar_wait(bar_a_empty + 8 * g, ((uint32_t)it & 1u) ^ 1u);       // the previous tile's MMAs have read these planes
        __syncwarp();
        ITG_SACC(1, tl);
        if (hp_ok && !(p.exp & 2)) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            sts128(dst + (uint32_t)((4 * g + q4) * PLANE_BYTES), w[4 * q4], w[4 * q4 + 1], w[4 * q4 + 2], w[4 * q4 + 3]);
        }
        if (!(p.exp & 1)) fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a_full + 8 * g);
        if (g + 1 < SSM_GROUPS) tmem_ld_wait(r);
        if (g + 2 == SSM_GROUPS) {                                     // the last group is in registers: the m1 accumulator may be overwritten
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_mlp_empty);
        }
        ITG_SACC(2, tl);
      }
    }
    if (p.dbg && blockIdx.x == 0 && warp == SSM_WARP_CVT && lane == 0) for (int i = 0; i < 3; ++i) p.dbg[4 + i] = dacc[i];
  } else {
    // ---- epilogue (two groups of four warps; group g drains the 16-column chunks c = g, g + 2 of every tile):
    //      embed accumulator -> (1 + gamma) * bn0(x) + beta -> activation -> framed store.  x does not depend on the accumulator: its
    //      16 bytes per chunk are fetched before sleeping on the MMA barrier; bias / mean / rstd of the column block come from shared memory ----
    const int eg = warp >> 2, q = warp & 3;
    const int row = q * 32 + lane;
    const EpiParams& ep = p.ep;
    const uint32_t vb = sbase + SSM_VEC_BIAS, vm = sbase + SSM_VEC_MEAN, vr = sbase + SSM_VEC_RSTD;
    const int n0 = nb * p.n_blk;
    unsigned long long dacc[2] = {0, 0};
    long long tl = p.dbg ? clock64() : 0;
    int tile = slot;
    for (int it = 0; it < n_my; ++it, tile += nslots) {
      const int b = it & 1;
      const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
      const int y = ty * TILE_H + (row >> 3), x = tx * TILE_W + (row & 7);
      const bool valid = (y < p.h) && (x < p.w);
      bool live[2], ok[2];
      uint4 xr[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = eg + 2 * j, n = n0 + 16 * c;
        live[j] = 16 * c < p.n_blk && n < p.n_pad;                    // warp-uniform: the chunk exists
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
      for (int j = 0; j < 2; ++j) {
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
        for (int h = 0; h < 4; ++h) {                                  // 4 columns = (gamma, beta) of two channels
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
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * b);
      ITG_SACC(1, tl);
    }
    if (p.dbg && blockIdx.x == 0 && warp == 0 && lane == 0) for (int i = 0; i < 2; ++i) p.dbg[8 + i] = dacc[i];
  }

  tc_fence_before();
  __syncthreads();
  if (warp == SSM_WARP_PROD) tmem_dealloc(tmem_base, 512);
}

}  // namespace itg
