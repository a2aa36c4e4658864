// Persistent "halo-tile" implicit-GEMM convolution for the thin layers of the Generator (K <= 64 channels per
// tap, N <= 64 GEMM columns: blocks 4-6 and the final conv of the 241 config, 52.6 % of its FLOPs).
//
// Those layers are bound by data movement and per-pixel epilogue work, not by the tensor pipe, so the kernel is
// organised around bytes and instruction counts:
//   * the weights of ALL taps (<= 72 KB) are parked in shared memory once per CTA for the whole launch;
//   * each 16 x 8 output tile loads its (16+2) x (8+2) input neighbourhood ONCE, as K/8 planes of
//     [pixel][8 channels] -- the nine taps of the 3x3 window (or the 4 x 4 phase taps of the folded nearest-2x
//     up-sampling) are nine shifted views of that one tile: the local-padding halo gather is a 16-byte-granular
//     offset in the MMA's shared-memory descriptor (no-swizzle K-major layout, 8 x 16 B core matrices: rows =
//     8 consecutive pixels of a tile row, SBO = one halo-tile row, LBO = one plane).  The planes are filled by
//     a producer warp with 16-byte cp.async (zero-fill outside the buffer), up to three tiles in flight;
//   * CTAs are persistent (grid = SM count) and run four independent pipelines {producer warp, MMA warp, four
//     epilogue warps}, each over its own tiles with a private ring of input stages and private TMEM accumulator
//     buffers, so the epilogue of tile i overlaps the loads and MMAs of the following tiles and every mbarrier has
//     one arriving and one waiting party (see TILE_PIPES below);
//   * the MMA warps run warp-uniform (one elected lane issues, predicated, no branch), the conv mode and the
//     epilogue variant are template parameters: measured on B200, the single-thread issue loop of the first
//     version cost ~320 cycles per tcgen05.mma and was the bottleneck (profiles/r01_notes.md).
#pragma once
#include <type_traits>
#include "conv_umma.cuh"

namespace itg {

constexpr int TILE_W = 8, TILE_H = 16;                       // output tile = 128 GEMM rows
constexpr int HALO_W = TILE_W + 2, HALO_H = TILE_H + 2;      // input neighbourhood
constexpr int HALO_PX = HALO_W * HALO_H;                     // 180
constexpr int PLANE_BYTES = HALO_PX * 16;                    // one 8-channel plane of the halo tile: 2880 B
#ifndef ITG_TILE_PLANE_PAD
#define ITG_TILE_PLANE_PAD 16
#endif
constexpr int TILE_PLANE = PLANE_BYTES + ITG_TILE_PLANE_PAD;  // plane pitch of the cp.async-filled tiles: 2896 B = 16 (mod 128), so that the kg lanes that copy
                                                             // the 16-byte chunks of one pixel to kg planes hit different bank groups (tools/ldgsts_probe.cu:
                                                             // 9 vs 17 cycles per warp instruction)
// Warp roles: the CTA runs TILE_PIPES independent pipelines.  Pipeline m = producer warp (4 + 4 * PIPES + m), MMA warp m
// (one per scheduler) and epilogue group m (warps 4 + 4m .. 7 + 4m; TMEM lane quarter = warp % 4); it owns the tiles
// it = m (mod pipes) of this CTA, a private ring of input stages and private TMEM accumulator buffers.  Every mbarrier
// therefore has exactly one arriving and one waiting party that walk its phases in order -- parity waits are only
// safe under that condition (a warp that shares a barrier with a faster sibling can alias a phase; an ncu
// instrumentation pass found exactly that in an earlier round-robin version).
constexpr int TILE_PIPES = 4;
constexpr int TILE_EPI_GROUPS = TILE_PIPES;
constexpr int TILE_MMA_WARPS = TILE_PIPES;
constexpr int TILE_WARPS = 4 + 4 * TILE_PIPES + TILE_PIPES;
constexpr int TILE_THREADS = 32 * TILE_WARPS;
constexpr int TILE_HDR_BYTES = 2048;       // barriers + per-channel epilogue vectors (bias | scale | shift | scale*bias+shift)
constexpr int TILE_MAX_STAGES = 24;
constexpr int TILE_MAX_NBUF = 8;            // TMEM accumulator buffers
constexpr int TILE_VEC_OFF = 576;           // header: full[24] | empty[24] | tfull[8] | tempty[8] | tmem slot | vectors (1 KB)

struct TileParams {
  int m_h, m_w;            // M-grid size (input interior)
  int tiles_x, ntiles;
  const void* in;          // framed grid tensor (buffer origin)
  int in_c, in_pitch;      // storage channels, pixels per buffer row
  int buf_h, buf_w;        // buffer extent in pixels (interior + frame)
  int in_cg_off;           // first 8-channel group of the input slice
  int kg;                  // 8-channel planes per tile (k_pad / 8): 2, 4 or 8
  int n;                   // GEMM columns (n_pad), multiple of 16, <= 64
  int n_src, k_src;        // row pitch / K pitch of the [tap][n_pad][k_pad] weight tensor in global memory
  int taps_w;              // taps in the weight tensor: 9 | 1 | 16
  int w_bytes;             // bytes of the shared-memory weight image
  int stage_bytes;         // bytes of one input stage (kg planes)
  int stages, ahead;       // input stages in total (= pipes * ring); tiles a producer keeps in flight before publishing (< ring)
  int pipes, ring;         // active pipelines (<= TILE_PIPES, <= nbuf) and input stages per pipeline
  int nbuf;                // TMEM accumulator buffers (power of two, multiple of pipes)
  uint32_t tmem_cols;
  uint32_t idesc;
  const void* w;
  unsigned long long* dbg;  // optional [grid][16] cycle counters (ITG_TILE_DBG=1), NULL in production
  int exp;                  // developer experiments (ITG_TILE_EXP bit mask, read with ITG_TILE_DBG only; WRONG RESULTS, timing only):
                            // 1 producer issues no loads, 2 epilogue does not store, 4 MMA warp issues one tap only
  EpiParams ep;
};

#define ITG_ACC(slot, tvar) do { if (p.dbg) { const long long now_ = clock64(); dbg_acc[slot] += (unsigned long long)(now_ - tvar); tvar = now_; } } while (0)

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
#ifndef ITG_CPASYNC_MODE
#define ITG_CPASYNC_MODE "ca"          // through L1: measured 2 % faster per step than .cg; a pixel-major lane mapping was 20 % slower
#endif
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;
  asm volatile("cp.async." ITG_CPASYNC_MODE ".shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async." ITG_CPASYNC_MODE ".shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_dyn(int n) {        // wait until at most n groups are pending
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
  }
}
// K-major operand without swizzle (8 x 16 B core matrices).  lo word: start address >> 4 | LBO >> 4 << 16
// (LBO = bytes between the two K halves of one MMA); hi word: SBO >> 4 (bytes between 8-row groups) | version.
__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr16, uint32_t lbo16, uint32_t sbo16) {
  const uint32_t lo = (addr16 & 0x3FFFu) | (lbo16 << 16);
  const uint32_t hi = sbo16 | (1u << 14);
  return ((uint64_t)hi << 32) | lo;
}

template <int MODE> struct TileMode;
template <> struct TileMode<ITG_CONV3X3> { static constexpr int NPHASE = 1, NTAPS = 9; };
template <> struct TileMode<ITG_CONV1X1> { static constexpr int NPHASE = 1, NTAPS = 1; };
template <> struct TileMode<ITG_UPCONV> { static constexpr int NPHASE = 4, NTAPS = 4; };

// All MMAs of one tile: NPHASE accumulators x NTAPS taps x KSTEPS 16-channel steps, fully unrolled.
template <int MODE, int KSTEPS>
__device__ __forceinline__ void issue_tile(uint32_t d0, uint32_t a16, uint32_t w16, uint32_t n16, uint32_t kg, uint32_t idesc) {
#pragma unroll
  for (int q = 0; q < TileMode<MODE>::NPHASE; ++q) {
#pragma unroll
    for (int t = 0; t < TileMode<MODE>::NTAPS; ++t) {
      int dy, dx, wt;
      tap_offsets(MODE, q, t, dy, dx, wt);
      const uint32_t shift16 = (uint32_t)((1 + dy) * HALO_W + (1 + dx));
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const uint64_t adesc = desc_noswz(a16 + (uint32_t)(2 * ks) * (TILE_PLANE / 16) + shift16, TILE_PLANE / 16, HALO_W);
        const uint64_t bdesc = desc_noswz(w16 + ((uint32_t)wt * kg + (uint32_t)(2 * ks)) * n16, n16, 8);
        umma_f16(d0 + (uint32_t)q * n16, adesc, bdesc, idesc, (t > 0 || ks > 0) ? 1u : 0u);
      }
    }
  }
}

template <typename T, int F, int MODE>
__global__ void __launch_bounds__(TILE_THREADS, 1)
conv_tile_kernel(const TileParams p) {
  constexpr int NPHASE = TileMode<MODE>::NPHASE;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t bar_full = sbase;                    // [stages <= 24]
  const uint32_t bar_empty = sbase + 192;             // [stages <= 24]
  const uint32_t bar_tfull = sbase + 384;             // [nbuf <= 8] accumulator buffer complete
  const uint32_t bar_tempty = sbase + 448;            // [nbuf <= 8] accumulator buffer drained
  const uint32_t tmem_slot = sbase + 512;
  float* vec = reinterpret_cast<float*>(smem_raw + (sbase - smem_u32(smem_raw)) + TILE_VEC_OFF);   // bias | scale | shift | folded, 64 floats each
  const uint32_t w_smem = sbase + TILE_HDR_BYTES;
  const uint32_t a_smem = w_smem + (uint32_t)p.w_bytes;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(bar_full + 8 * i, 32);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < p.nbuf; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 4);               // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, p.tmem_cols);

  // ---- park the weights of every tap in shared memory: image [tap][k-group][n][8 channels] ----
  {
    const T* wg = reinterpret_cast<const T*>(p.w);
    const int chunks = p.taps_w * p.kg * p.n;                                  // 16-byte chunks
    for (int i = threadIdx.x; i < chunks; i += TILE_THREADS) {
      const int nn = i % p.n, j = (i / p.n) % p.kg, t = i / (p.n * p.kg);
      const uint4 v = *reinterpret_cast<const uint4*>(wg + ((size_t)t * p.n_src + nn) * p.k_src + j * 8);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(w_smem + (uint32_t)i * 16u), "r"(v.x), "r"(v.y),
                   "r"(v.z), "r"(v.w)
                   : "memory");
    }
    fence_proxy_async();                                                       // generic writes -> async proxy (UMMA) reads
    if (threadIdx.x < 64) {
      const int i = threadIdx.x;
      vec[i] = (p.ep.bias != nullptr && i < p.n) ? p.ep.bias[i] : 0.f;
      vec[64 + i] = (p.ep.scale != nullptr && i < p.n) ? p.ep.scale[i] : 1.f;
      vec[128 + i] = (p.ep.shift != nullptr && i < p.n) ? p.ep.shift[i] : 0.f;
      vec[192 + i] = fmaf(vec[64 + i], vec[i], vec[128 + i]);       // BN folded over the bias: scale * (acc + bias) + shift
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();                                          // activations of the previous launch are complete and visible from here on

  // Loop bookkeeping without integer division: ring positions advance by a fixed step, tile coordinates by a
  // precomputed (dy, dx) with carry.  mbarrier waits are done by lane 0 only (a 32-lane try_wait on one barrier
  // measured ~300 cycles even when already complete) followed by __syncwarp.
  if (warp >= 4 + 4 * TILE_EPI_GROUPS) {                                        // ---- producers: one warp per pipeline, all 32 lanes on one tile ----
    const int pw = warp - (4 + 4 * TILE_EPI_GROUPS);
    const int kg_log2 = 31 - __clz(p.kg);
    const int cg_total = p.in_c >> 3;
    const T* in = reinterpret_cast<const T*>(p.in);
    // 32 % kg == 0: a lane always serves the same 8-channel plane j and walks the halo pixels with a fixed stride
    const int j = lane & (p.kg - 1), px0 = lane >> kg_log2, px_step = 32 >> kg_log2;
    const bool cg_ok = (p.in_cg_off + j) < cg_total;
    const uint32_t ch_off = (uint32_t)((p.in_cg_off + j) * 8);
    const uint32_t plane_off = (uint32_t)(j * TILE_PLANE);
    // per-lane offsets of its halo pixels (<= 45 at kg = 2 ... 12 at kg = 8): computed on the fly, two multiplies each
    unsigned long long dbg_acc[4] = {0, 0, 0, 0};
    long long tl = p.dbg ? clock64() : 0;
    const int step = p.pipes * (int)gridDim.x;
    const int sdy = step / p.tiles_x, sdx = step - sdy * p.tiles_x;
    int tile = pw < p.pipes ? blockIdx.x + pw * (int)gridDim.x : p.ntiles;
    int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
    const int s0 = pw * p.ring;            // first stage of this pipeline's private ring
    int s = s0;                            // stage of tile k
    uint32_t ph = 0;
    int s_pub = s0;                        // stage of the oldest tile that is loaded (or loading) but not yet published
    int k = 0, k_pub = 0;                  // tiles issued / published
    auto publish_upto = [&](int k_end) {   // tiles [k_pub, k_end) have landed: hand them to the MMA warp
      fence_proxy_async();
      for (; k_pub < k_end; ++k_pub) {
        mbar_arrive(bar_full + 8 * s_pub);
        if (++s_pub == s0 + p.ring) s_pub = s0;
      }
    };
    for (; tile < p.ntiles; tile += step, ++k) {
      const int y0 = ty * TILE_H, x0 = tx * TILE_W;                            // halo origin in buffer pixels
      // A free stage for tile k means the MMAs of tile k - ring have finished.  If that is not yet the case, everything loaded so far is
      // published BEFORE sleeping: with a two-stage ring the old order (publish tile k-1 only after tile k's loads were issued) chained
      // MMA(k) behind MMA(k-1) plus a whole load-issue pass -- the K = 64 layers ran at a third of their MMA rate with the loads idle.
      uint32_t ready = 0;
      if (lane == 0) ready = mbar_try_wait(bar_empty + 8 * s, ph ^ 1u) ? 1u : 0u;
      ready = __shfl_sync(0xffffffffu, ready, 0);
      if (!ready) {
        if (k_pub < k) {
          cp_async_wait_dyn(0);
          publish_upto(k);
        }
        if (lane == 0) mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        __syncwarp();
      }
      ITG_ACC(0, tl);
      const uint32_t dst0 = a_smem + (uint32_t)(s * p.stage_bytes) + plane_off;
      const T* src0 = in + ((size_t)y0 * p.in_pitch + x0) * (size_t)p.in_c + ch_off;
      const bool inside = cg_ok && (y0 + HALO_H <= p.buf_h) && (x0 + HALO_W <= p.buf_w);
      if (p.exp & 1) {
#ifndef ITG_NO_1X1_INTERIOR
      } else if (MODE == ITG_CONV1X1 && inside) {
        // a 1x1 conv reads no neighbours: only the 16 x 8 interior of the halo tile is fetched (29 % fewer bytes and copies; the ring keeps
        // whatever an earlier tile left there and is never addressed by the single tap)
#pragma unroll 4
        for (int px = px0; px < TILE_W * TILE_H; px += px_step) {
          const int hy = (px >> 3) + 1, hx = (px & 7) + 1;
          cp_async16(dst0 + (uint32_t)((hy * HALO_W + hx) * 16), src0 + (uint32_t)((hy * p.in_pitch + hx) * p.in_c));
        }
#endif
      } else if (inside) {
#pragma unroll 4
        for (int px = px0; px < HALO_PX; px += px_step) {
          const int hy = (px * 205) >> 11, hx = px - hy * HALO_W;
          cp_async16(dst0 + (uint32_t)(px * 16), src0 + (uint32_t)((hy * p.in_pitch + hx) * p.in_c));
        }
      } else {
        for (int px = px0; px < HALO_PX; px += px_step) {
          const int hy = (px * 205) >> 11, hx = px - hy * HALO_W;
          const bool valid = cg_ok && (y0 + hy < p.buf_h) && (x0 + hx < p.buf_w);
          cp_async16_zfill(dst0 + (uint32_t)(px * 16), valid ? src0 + (uint32_t)((hy * p.in_pitch + hx) * p.in_c) : in, valid);
        }
      }
      cp_async_commit();
      ITG_ACC(1, tl);
      if (k + 1 - k_pub > p.ahead) {                     // keep at most `ahead` tiles in flight: the older ones have landed, publish them
        cp_async_wait_dyn(p.ahead);
        ITG_ACC(2, tl);
        publish_upto(k + 1 - p.ahead);
        ITG_ACC(3, tl);
      }
      if (++s == s0 + p.ring) { s = s0; ph ^= 1u; }
      tx += sdx; ty += sdy; if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
    }
    cp_async_wait_dyn(0);
    publish_upto(k);
    if (p.dbg && pw == 0 && lane == 0) for (int i = 0; i < 4; ++i) p.dbg[blockIdx.x * 16 + i] = dbg_acc[i];
  } else if (warp < TILE_MMA_WARPS) {                                          // ---- MMA warps (uniform; one lane issues), tiles it = mw (mod NM) ----
    const int mw = warp;
    unsigned long long dbg_acc[4] = {0, 0, 0, 0};
    long long tl = p.dbg ? clock64() : 0;
    const int ksteps = p.kg >> 1;
    const uint32_t w16 = w_smem >> 4, n16 = (uint32_t)p.n;                     // weight image: 16 B per (k-group, n)
    const int nbuf_log2 = 31 - __clz(p.nbuf);
    const int nm = p.pipes;
    const int s0 = mw * p.ring;
    int s = s0;
    uint32_t ph = 0;
    int it = mw;
    for (int tile = mw < nm ? blockIdx.x + mw * (int)gridDim.x : p.ntiles; tile < p.ntiles; tile += nm * gridDim.x, it += nm) {
      const int b = it & (p.nbuf - 1);
      const uint32_t bph = (uint32_t)(it >> nbuf_log2) & 1u;
      if (lane == 0) {
        mbar_wait(bar_tempty + 8 * b, bph ^ 1u);                               // epilogue has drained this accumulator
        ITG_ACC(0, tl);
        mbar_wait(bar_full + 8 * s, ph);                                       // the halo tile has landed
      }
      __syncwarp();
      ITG_ACC(1, tl);
      tc_fence_after();
      const uint32_t a16 = (a_smem + (uint32_t)(s * p.stage_bytes)) >> 4;
      if (elect_one_sync()) {                                                    // one elected lane issues the whole tile from a branch that ptxas
        const uint32_t d0 = tmem_base + (uint32_t)(b * NPHASE * p.n);            // recognises as single-threaded: descriptors stay on the uniform datapath
        if (p.exp & 4) issue_tile<ITG_CONV1X1, 1>(d0, a16, w16, n16, (uint32_t)p.kg, p.idesc);
        else if (ksteps == 1) issue_tile<MODE, 1>(d0, a16, w16, n16, (uint32_t)p.kg, p.idesc);
        else if (ksteps == 2) issue_tile<MODE, 2>(d0, a16, w16, n16, (uint32_t)p.kg, p.idesc);
        else issue_tile<MODE, 4>(d0, a16, w16, n16, (uint32_t)p.kg, p.idesc);
        umma_commit(bar_empty + 8 * s);                                        // input stage may be refilled
        umma_commit(bar_tfull + 8 * b);                                        // accumulators of this tile complete
      }
      __syncwarp();
      ITG_ACC(2, tl);
      if (++s == s0 + p.ring) { s = s0; ph ^= 1u; }
    }
    if (p.dbg && lane == 0 && mw == 0) for (int i = 0; i < 4; ++i) p.dbg[blockIdx.x * 16 + 4 + i] = dbg_acc[i];
  } else if (warp >= 4 && warp < 4 + 4 * TILE_EPI_GROUPS) {                    // ---- epilogue ----
    // group g drains the tiles of pipeline g (accumulator buffer it % nbuf)
    const int g = (warp - 4) >> 2;
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    const EpiParams& ep = p.ep;
    const uint32_t vec_smem = sbase + TILE_VEC_OFF;     // bias | scale | shift copies (see load_vec8)
    // Fast path (tile not touching the image border, specialised epilogue): addresses are formed once per pixel, the
    // per-channel vectors come from shared memory, no frame logic; everything else takes the general path below.
    constexpr bool FASTF = (F & EF_GENERIC) == 0;
    const bool fast = FASTF && ((F & EF_IMG) != 0 || ep.out_c <= p.n);
    unsigned long long dbg_acc[4] = {0, 0, 0, 0};
    long long tl = p.dbg ? clock64() : 0;
    const int nbuf_log2 = 31 - __clz(p.nbuf);
    const int ngroups = p.pipes;
    const int step = ngroups * (int)gridDim.x;
    const int sdy = step / p.tiles_x, sdx = step - sdy * p.tiles_x;
    int tile = g < ngroups ? blockIdx.x + g * (int)gridDim.x : p.ntiles;
    int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
    int it = g;
    constexpr bool PRE = (F & EF_GENERIC) == 0 && (F & EF_RES) != 0 && NPHASE == 1;
    // pixels of 16 k channels are 32-byte aligned: 256-bit loads / stores, one per 16 channels
    const bool wide = sizeof(T) == 2 && (ep.out_c & 15) == 0 && (!(F & EF_RES) || (ep.res_c & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(ep.out_raw) | reinterpret_cast<uintptr_t>(ep.out_act) | reinterpret_cast<uintptr_t>(ep.res)) & 31) == 0;
    for (; tile < p.ntiles; tile += step, it += ngroups) {
      const int b = it & (p.nbuf - 1);
      const uint32_t bph = (uint32_t)(it >> nbuf_log2) & 1u;
      const int y = ty * TILE_H + (row >> 3), x = tx * TILE_W + (row & 7);
      const bool valid = (y < p.m_h) && (x < p.m_w) && !(p.exp & 2);
      // residual rows do not depend on the accumulators: fetch them before sleeping on the MMA barrier (fetching them a
      // whole tile ahead was measured slower: the extra live registers spill under the 80-register cap)
      uint4 pre[8];
      if (PRE && valid) {
        const T* rp = reinterpret_cast<const T*>(ep.res) + grid_off(y >> ep.res_shift, x >> ep.res_shift, ep.res_w, ep.res_c, 0);
        if (wide) {                                                              // 16 channels (32 B) per load
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i * 16 < p.n && i * 16 < ep.out_c) ldg256(rp + i * 16, pre[2 * i], pre[2 * i + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i * 8 < p.n && i * 8 < ep.out_c) pre[i] = *reinterpret_cast<const uint4*>(rp + i * 8);
        }
      }
      if (lane == 0) mbar_wait(bar_tfull + 8 * b, bph);
      __syncwarp();
      ITG_ACC(0, tl);
      tc_fence_after();
      // interior tile: every pixel valid and none of its outputs on the image border (no frame writes needed)
      const bool interior = fast && ty > 0 && tx > 0 && (ty + 1) * TILE_H < p.m_h && (tx + 1) * TILE_W < p.m_w && !(p.exp & 2);
      if (interior) {
#pragma unroll
        for (int q = 0; q < NPHASE; ++q) {
          int oy = y, ox = x;
          if (MODE == ITG_UPCONV) { oy = 2 * y + (q >> 1); ox = 2 * x + (q & 1); }
          const uint32_t trow = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)((b * NPHASE + q) * p.n);
          if (F & EF_IMG) {                                                      // final conv: tanh -> planar fp32 image
            float v[16];
            tmem_ld16(trow, v);
            if (ep.img_layout == ITG_IMG_MERGED) {
              float* o = ep.out_img + (size_t)oy * ep.out_w + ox;
              const size_t plane = (size_t)ep.out_h * ep.out_w;
              float bb[8];
              const float4 b0 = lds_f4(vec_smem), b1 = lds_f4(vec_smem + 16);
              bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
#pragma unroll
              for (int c = 0; c < 8; ++c)
                if (c < ep.img_c) o[c * plane] = tanh_fast(v[c] + bb[c]);
            } else {
              float a[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) a[i] = v[i];
              epilogue8<T, F>(ep, oy, ox, 0, a, nullptr, vec_smem);
            }
            continue;
          }
          const size_t off = grid_off(oy, ox, ep.out_w, ep.out_c, 0);
          const T* rp = (F & EF_RES) ? reinterpret_cast<const T*>(ep.res) + grid_off(oy >> ep.res_shift, ox >> ep.res_shift, ep.res_w, ep.res_c, 0) : nullptr;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const int c0 = cc * 16;
            if (c0 >= p.n) break;
            float v[16];
            tmem_ld16(trow + (uint32_t)c0, v);
            if (wide && c0 < ep.out_c) {                                       // 16 channels: one 256-bit store per output tensor
              uint32_t rw[8], aw[8];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint32_t o4 = vec_smem + (uint32_t)((c0 + 8 * h) * 4);
                float b8[8], x8[8];
                {
                  const float4 ba = lds_f4(o4), bb = lds_f4(o4 + 16);
                  b8[0] = ba.x; b8[1] = ba.y; b8[2] = ba.z; b8[3] = ba.w; b8[4] = bb.x; b8[5] = bb.y; b8[6] = bb.z; b8[7] = bb.w;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) x8[i] = v[8 * h + i];
                if (F & EF_RES) {
                  uint4 rr;
                  if (PRE) rr = pre[2 * cc + h];
                  else rr = *reinterpret_cast<const uint4*>(rp + c0 + 8 * h);
                  Vec8<T> t0 = *reinterpret_cast<const Vec8<T>*>(&rr);
#pragma unroll
                  for (int i = 0; i < 8; ++i) x8[i] += Op<T>::to_f(t0.v[i]);
                }
                if (F & EF_RAW) {
#pragma unroll
                  for (int i = 0; i < 4; ++i) rw[4 * h + i] = pack2<T>(x8[2 * i] + b8[2 * i], x8[2 * i + 1] + b8[2 * i + 1]);
                }
                if (F & EF_ACT) {
                  const float4 sa = lds_f4(o4 + 256), sb = lds_f4(o4 + 272), ta = lds_f4(o4 + 768), tb = lds_f4(o4 + 784);
                  const float s8[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                  const float t8[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
                  float w8[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) w8[i] = fmaf(s8[i], x8[i], t8[i]);
                  if (!ep.act_linear) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) w8[i] = act_fn(w8[i], ep.leak);
                  }
#pragma unroll
                  for (int i = 0; i < 4; ++i) aw[4 * h + i] = pack2<T>(w8[2 * i], w8[2 * i + 1]);
                }
              }
              if (F & EF_RAW) stg256(reinterpret_cast<T*>(ep.out_raw) + off + c0, rw);
              if (F & EF_ACT) stg256(reinterpret_cast<T*>(ep.out_act) + off + c0, aw);
              continue;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {                                      // 8 channels at a time keeps the live registers low
              const int ch = c0 + 8 * h;
              if (ch >= ep.out_c) continue;
              float b8[8], s8[8], t8[8], x8[8];
              {
                const uint32_t o4 = vec_smem + (uint32_t)(ch * 4);
                const float4 ba = lds_f4(o4), bb = lds_f4(o4 + 16);
                b8[0] = ba.x; b8[1] = ba.y; b8[2] = ba.z; b8[3] = ba.w; b8[4] = bb.x; b8[5] = bb.y; b8[6] = bb.z; b8[7] = bb.w;
                if (F & EF_ACT) {
                  const float4 sa = lds_f4(o4 + 256), sb = lds_f4(o4 + 272), ta = lds_f4(o4 + 768), tb = lds_f4(o4 + 784);
                  s8[0] = sa.x; s8[1] = sa.y; s8[2] = sa.z; s8[3] = sa.w; s8[4] = sb.x; s8[5] = sb.y; s8[6] = sb.z; s8[7] = sb.w;
                  t8[0] = ta.x; t8[1] = ta.y; t8[2] = ta.z; t8[3] = ta.w; t8[4] = tb.x; t8[5] = tb.y; t8[6] = tb.z; t8[7] = tb.w;
                }
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) x8[i] = v[8 * h + i];
              if (F & EF_RES) {
                float r8[8];
                if (PRE) {
                  Vec8<T> t0 = *reinterpret_cast<const Vec8<T>*>(&pre[2 * cc + h]);
#pragma unroll
                  for (int i = 0; i < 8; ++i) r8[i] = Op<T>::to_f(t0.v[i]);
                } else {
                  load8(rp + ch, r8);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) x8[i] += r8[i];
              }
              if (F & EF_RAW) {
                float w8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w8[i] = x8[i] + b8[i];
                store8(reinterpret_cast<T*>(ep.out_raw) + off + ch, w8);
              }
              if (F & EF_ACT) {
                float w8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w8[i] = fmaf(s8[i], x8[i], t8[i]);
                if (!ep.act_linear) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) w8[i] = act_fn(w8[i], ep.leak);
                }
                store8(reinterpret_cast<T*>(ep.out_act) + off + ch, w8);
              }
            }
          }
        }
      } else
#pragma unroll
      for (int q = 0; q < NPHASE; ++q) {
        int oy = y, ox = x;
        if (MODE == ITG_UPCONV) { oy = 2 * y + (q >> 1); ox = 2 * x + (q & 1); }
        const uint32_t trow = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)((b * NPHASE + q) * p.n);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {                                       // n <= 64: compile-time indices keep pre[] in registers
          const int c0 = cc * 16;
          if (c0 < p.n) {
            float v[16];
            tmem_ld16(trow + (uint32_t)c0, v);
            if (valid) {
              float a[8], c[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) { a[i] = v[i]; c[i] = v[8 + i]; }
              epilogue8<T, F>(ep, oy, ox, c0, a, PRE ? &pre[2 * cc] : nullptr, vec_smem);
              epilogue8<T, F>(ep, oy, ox, c0 + 8, c, PRE ? &pre[2 * cc + 1] : nullptr, vec_smem);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
      ITG_ACC(1, tl);
      tx += sdx; ty += sdy; if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
    }
    if (p.dbg && ew == 0 && lane == 0) for (int i = 0; i < 2; ++i) p.dbg[blockIdx.x * 16 + 8 + g * 2 + i] = dbg_acc[i];
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace itg
