// Implicit-GEMM local-padding convolution on the 5th-gen tensor cores (sm_100a): TMA -> shared memory ->
// tcgen05.mma (accumulators in TMEM) -> tcgen05.ld -> fused epilogue.
//
// GEMM view:  D[m][n] = sum_{tap} sum_{k} A_tap[m][k] * W[tap][n][k]
//   m : 128 M-grid pixels of a TH x TW tile (TH*TW = 128) of the merged grid tensor
//   k : input channels (chunks of KC = 16/32/64 channels = one 32/64/128-byte swizzle row)
//   n : output channels (GEMM columns), n_blk <= 256 per CTA
// A_tap is the tile shifted by the tap offset (dy,dx): because the grid tensor carries its 1-pixel frame
// (outer padding / sequential halo / neighbour-GPU halo), the shifted tile is one 3-D TMA box
// {KC, TW, TH} at (c0, x0+dx+1, y0+dy+1) -- the local-padding "halo gather" is the TMA coordinate, and no
// padded patch is ever materialised.  TMA zero-fills what lies beyond the buffer (partial tiles, channel tail).
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4..7 = epilogue (warp w reads TMEM lanes 32*(w%4) .. +31).
#pragma once
#include <cuda.h>
#include "itg_common.cuh"

namespace itg {

struct UmmaParams {
  int m_h, m_w;          // M-grid size
  int in_c_off;
  int mode;
  int tw_log2, tiles_x;
  int n_pad, n_blk, nblocks;
  int nwork;             // work items = M tiles x phases x N blocks
  int cluster;           // CTAs per cluster (1, 2, 4 or 8).  > 1: the cluster's CTAs are the N blocks of one (tile, phase)
                         // item and share its activation tiles through TMA multicast; nblocks == cluster
  int nitems;            // M tiles x phases
  unsigned long long* dbg;  // optional per-role cycle counters of CTA 0 (ITG_TILE_DBG=1)
  int kc, nchunks, ksteps_last;
  int stages;
  int teams, groups;     // active epilogue groups (2 or 4) in 1 or 2 teams.  teams = 1: every active group drains every work item (latency:
                         // few items per CTA); teams = 2: team t drains the items li = t (mod 2), i.e. accumulator buffer t (throughput)
  int a_bytes, b_bytes;  // per-stage operand bytes (also the TMA transaction size)
  int a_stride, b_stride;  // 1024-aligned strides inside a stage
  uint32_t sbo_enc;      // (8 * swizzle bytes) >> 4
  uint32_t layout_type;  // UMMA smem-descriptor layout type: 2 = SW128, 4 = SW64, 6 = SW32
  uint32_t tmem_cols;
  uint32_t idesc;
  EpiParams ep;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a broken pipeline traps (the launch fails) instead of hanging the device.
#ifndef ITG_MBAR_TIMEOUT_CYCLES
#define ITG_MBAR_TIMEOUT_CYCLES 4000000000LL      // ~2 s
#endif
// (A suspend-time hint on try_wait -- letting the hardware park the waiting thread for up to 100 us -- was measured 6 % SLOWER on every
// workload: the wake-up latency costs more than the polling instructions it saves.)
// The slow path is a real function call: inlined at every wait site it put a clock read, a printf call and a trap (~40 instructions) into
// each hot loop, and the kernels' code outgrew the instruction caches (ncu: `no_instruction` was the top stall of the SSM kernel).
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  // bounded: a broken pipeline traps instead of hanging the GPU.  Both a cycle budget and a poll budget must be exhausted:
  // a profiler that freezes the SM for seconds (ncu PM-sampling passes) advances the clock but not the polls.
  const long long t0 = clock64();
  unsigned polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++polls > (1u << 20) && clock64() - t0 > ITG_MBAR_TIMEOUT_CYCLES) {
      printf("itg: mbarrier wait timed out (block %d,%d,%d thread %d bar 0x%x parity %u)\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  if (mbar_try_wait(bar, parity)) return;          // (one hardware-suspended retry before paying for the call)
  mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int kDummy = 0>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16 / bf16 operands, fp32 accumulate), issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// tcgen05.mma executed by the lane whose `leader` flag is set, WITHOUT a branch: the surrounding code stays in uniform
// control flow, so the descriptors are computed on the uniform datapath instead of being moved lane -> uniform
// register (R2UR + ELECT loops) for every instruction.
__device__ __forceinline__ void umma_f16_pred(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pred(uint32_t bar, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar), "r"(leader)
      : "memory");
}

// commit that arrives on the same barrier offset in every CTA of `mask` (stage release of a multicast operand ring)
__device__ __forceinline__ void umma_commit_mc_pred(uint32_t bar, uint16_t mask, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(bar), "h"(mask), "r"(leader)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major operand descriptor: rows of `swizzle` bytes, 8-row groups SBO apart (canonical TMA swizzled layout)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo_enc, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(sbo_enc & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                  // descriptor version (sm_100)
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
// Persistent CTAs (grid = min(work items, SM count)).  A work item is one 128-pixel tile x n_blk GEMM columns
// (x one of the 4 phases of the folded up-sampling conv); a CTA walks its items with the operand ring never draining
// between items, and with two TMEM accumulator buffers so that the epilogue of item i (two groups of four warps,
// alternating items) overlaps the TMA loads and MMAs of item i+1.
#define ITG_UACC(slot, tvar) do { if (p.dbg) { const long long now_ = clock64(); dacc[slot] += (unsigned long long)(now_ - tvar); tvar = now_; } } while (0)
#ifndef ITG_UMMA_EPI_GROUPS
#define ITG_UMMA_EPI_GROUPS 4
#endif
constexpr int UMMA_EPI_GROUPS = ITG_UMMA_EPI_GROUPS;      // groups of four epilogue warps; group g drains the 16-column chunks c = g (mod G) of EVERY item
constexpr int UMMA_THREADS = 32 * (4 + 4 * UMMA_EPI_GROUPS);     // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4.. epilogue
constexpr int UMMA_BAR_BYTES = 1024;
constexpr int UMMA_MAX_STAGES = 8;

__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <typename T, int F>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const UmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t bar_full = sbase;                 // [stages] x 8 B
  const uint32_t bar_empty = sbase + 128;          // [stages] x 8 B
  const uint32_t bar_tfull = sbase + 256;          // [2] accumulator buffer complete
  const uint32_t bar_tempty = sbase + 272;         // [2] accumulator buffer drained
  const uint32_t tmem_slot = sbase + 288;
  const uint32_t stage0 = sbase + UMMA_BAR_BYTES;
  const uint32_t stage_bytes = (uint32_t)(p.a_stride + p.b_stride);

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, (uint32_t)p.cluster);   // every CTA of the cluster releases every stage (multicast commit)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 4 * p.groups / p.teams);   // one arrival per epilogue warp of the team that drains this buffer
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();           // peers' barriers are initialised before anybody multicasts into them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();                                          // the previous launch's outputs (our activations) are complete from here on

  const int tw = 1 << p.tw_log2, th = 128 >> p.tw_log2;
  const int ntaps = (p.mode == ITG_CONV3X3) ? 9 : (p.mode == ITG_CONV1X1 ? 1 : 4);
  const int nphase = (p.mode == ITG_UPCONV) ? 4 : 1;
  const uint32_t buf_cols = p.tmem_cols >> 1;      // TMEM columns per accumulator buffer

  // work item w -> (tile, phase, n block).  cluster == 1: n blocks of one tile run on neighbouring CTAs and share its A
  // tiles in L2.  cluster > 1: a cluster walks the (tile, phase) items in lockstep, CTA rank r owns n block r and the
  // activation tile of every k iteration is fetched ONCE (round-robin issuer) and multicast into all CTAs of the cluster.
  const int crank = p.cluster > 1 ? (int)cluster_ctarank() : 0;
  const int w_first = p.cluster > 1 ? (int)blockIdx.x / p.cluster : (int)blockIdx.x;
  const int w_stride = p.cluster > 1 ? (int)gridDim.x / p.cluster : (int)gridDim.x;
  const int w_limit = p.cluster > 1 ? p.nitems : p.nwork;
  const uint16_t cmask = (uint16_t)((1u << p.cluster) - 1u);
  auto decode = [&](int w, int& tile, int& phase, int& n0) {
    int rest = w, nb = crank;
    if (p.cluster == 1) { nb = w % p.nblocks; rest = w / p.nblocks; }
    n0 = nb * p.n_blk;
    phase = rest % nphase;
    tile = rest / nphase;
  };

  if (warp == 0) {
    if (lane == 0) {                                                   // ---- TMA producer ----
      int s = 0, issuer = 0;
      uint32_t ph = 0;
      unsigned long long dacc[2] = {0, 0};
      long long tl = p.dbg ? clock64() : 0;
      for (int w = w_first; w < w_limit; w += w_stride) {
        int tile, phase, n0;
        decode(w, tile, phase, n0);
        const int y0 = (tile / p.tiles_x) * th, x0 = (tile % p.tiles_x) * tw;
        for (int t = 0; t < ntaps; ++t) {
          int dy, dx, wt;
          tap_offsets(p.mode, phase, t, dy, dx, wt);
          for (int c = 0; c < p.nchunks; ++c) {
            mbar_wait(bar_empty + 8 * s, ph ^ 1u);           // cluster > 1: ALL CTAs of the cluster have consumed this stage
            ITG_UACC(0, tl);
            mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)(p.a_bytes + p.b_bytes));
            const uint32_t sa = stage0 + s * stage_bytes;
            if (p.cluster == 1) {
              tma_load_3d(sa, &tm_a, bar_full + 8 * s, p.in_c_off + c * p.kc, x0 + dx + 1, y0 + dy + 1);
            } else {
              if (issuer == crank)
                tma_load_3d_mc(sa, &tm_a, bar_full + 8 * s, p.in_c_off + c * p.kc, x0 + dx + 1, y0 + dy + 1, cmask);
              if (++issuer == p.cluster) issuer = 0;
            }
            tma_load_2d(sa + p.a_stride, &tm_b, bar_full + 8 * s, c * p.kc, wt * p.n_pad + n0);
            if (++s == p.stages) { s = 0; ph ^= 1u; }
            ITG_UACC(1, tl);
          }
        }
      }
      if (p.dbg && blockIdx.x == 0) { p.dbg[0] = dacc[0]; p.dbg[1] = dacc[1]; }
    }
    __syncwarp();
  } else if (warp == 1) {                                              // ---- MMA warp: uniform control flow, one elected lane issues ----
    int s = 0, li = 0;
    uint32_t ph = 0;
    unsigned long long dacc[3] = {0, 0, 0};
    long long tl = p.dbg ? clock64() : 0;
    for (int w = w_first; w < w_limit; w += w_stride, ++li) {
      const int b = li & 1;
      const uint32_t bph = (uint32_t)(li >> 1) & 1u;
      if (lane == 0) mbar_wait(bar_tempty + 8 * b, bph ^ 1u);          // the epilogue has drained this accumulator buffer
      __syncwarp();
      ITG_UACC(0, tl);
      tc_fence_after();
      const uint32_t dcol = tmem_base + (uint32_t)b * buf_cols;
      int it = 0;
      for (int t = 0; t < ntaps; ++t) {
        for (int c = 0; c < p.nchunks; ++c, ++it) {
          if (lane == 0) mbar_wait(bar_full + 8 * s, ph);
          __syncwarp();
          ITG_UACC(1, tl);
          tc_fence_after();
          const uint32_t sa = stage0 + s * stage_bytes;
          const uint64_t adesc = make_smem_desc(sa, p.sbo_enc, p.layout_type);
          const uint64_t bdesc = make_smem_desc(sa + p.a_stride, p.sbo_enc, p.layout_type);
          const int nk = (c == p.nchunks - 1) ? p.ksteps_last : (p.kc >> 4);
          if (elect_one_sync()) {                 // single-threaded branch: ptxas keeps the descriptors on the uniform datapath (no ELECT / R2UR loop)
            umma_f16(dcol, adesc, bdesc, p.idesc, it > 0 ? 1u : 0u);                        // +32 B (16 channels) per K step
            if (nk > 1) umma_f16(dcol, adesc + 2, bdesc + 2, p.idesc, 1u);
            if (nk > 2) umma_f16(dcol, adesc + 4, bdesc + 4, p.idesc, 1u);
            if (nk > 3) umma_f16(dcol, adesc + 6, bdesc + 6, p.idesc, 1u);
            if (p.cluster == 1) umma_commit(bar_empty + 8 * s);                             // frees the stage when these MMAs have read it
            else umma_commit_mc_pred(bar_empty + 8 * s, cmask, 1u);                         // ... in every CTA of the cluster
            if (t == ntaps - 1 && c == p.nchunks - 1) umma_commit(bar_tfull + 8 * b);       // accumulators of this item complete
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1u; }
          ITG_UACC(2, tl);
        }
      }
    }
    if (p.dbg && blockIdx.x == 0 && lane == 0) { p.dbg[2] = dacc[0]; p.dbg[3] = dacc[1]; p.dbg[4] = dacc[2]; }
  } else if (warp >= 4) {                                              // ---- epilogue: both groups drain every item, group g takes the
    const int g = (warp - 4) >> 2;                                     //      16-column chunks c = g (mod 2): half the latency per item ----
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    // team t = groups of which it consists drain the items li = t (mod teams) (accumulator buffer li & 1); inside a team
    // group gi of gpt takes the 16-column chunks c = gi (mod gpt)
    const int gpt = p.groups / p.teams;
    const int team = g / gpt, gi = g - team * gpt;
    int li = team;
    unsigned long long dacc[2] = {0, 0};
    long long tl = p.dbg ? clock64() : 0;
    const int cstep = 16 * gpt;
    for (int w = g < p.groups ? w_first + team * w_stride : w_limit; w < w_limit; w += p.teams * w_stride, li += p.teams) {
      int tile, phase, n0;
      decode(w, tile, phase, n0);
      const int b = li & 1;
      const uint32_t bph = (uint32_t)(li >> 1) & 1u;
      const int y = (tile / p.tiles_x) * th + (row >> p.tw_log2), x = (tile % p.tiles_x) * tw + (row & (tw - 1));
      const bool valid = (y < p.m_h) && (x < p.m_w);
      int oy = y, ox = x;
      if (p.mode == ITG_UPCONV) { oy = 2 * y + (phase >> 1); ox = 2 * x + (phase & 1); }
      // interior tile (all pixels valid, none on the image border) with a specialised epilogue: addresses are formed once
      // per pixel and no frame logic runs; border tiles and the generic / SSM / image epilogues take the general path
      const int ty0 = (tile / p.tiles_x) * th, tx0 = (tile % p.tiles_x) * tw;
      const bool interior = (F & (EF_GENERIC | EF_IMG)) == 0 && ty0 > 0 && tx0 > 0 && ty0 + th < p.m_h && tx0 + tw < p.m_w;
      // the residual does not depend on the accumulators: the chunk's 32 bytes are fetched one chunk ahead, the first
      // one before sleeping on the MMA barrier
      const T* rp = nullptr;
      uint4 rn0 = make_uint4(0, 0, 0, 0), rn1 = rn0;
      if ((F & EF_RES) != 0 && interior) {
        rp = reinterpret_cast<const T*>(p.ep.res) + grid_off(oy >> p.ep.res_shift, ox >> p.ep.res_shift, p.ep.res_w, p.ep.res_c, 0);
        const int ch = n0 + 16 * gi;
        if (16 * gi < p.n_blk && ch < p.ep.out_c) {
          rn0 = *reinterpret_cast<const uint4*>(rp + ch);
          if (ch + 8 < p.ep.out_c) rn1 = *reinterpret_cast<const uint4*>(rp + ch + 8);
        }
      }
      if (lane == 0) mbar_wait(bar_tfull + 8 * b, bph);
      __syncwarp();
      ITG_UACC(0, tl);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)b * buf_cols;
      if (interior) {
        const EpiParams& ep = p.ep;
        const size_t off = grid_off(oy, ox, ep.out_w, ep.out_c, 0);
        for (int c0 = 16 * gi; c0 < p.n_blk; c0 += cstep) {
          uint4 rc[2] = {rn0, rn1};
          if (F & EF_RES) {
            const int chn = n0 + c0 + cstep;
            if (c0 + cstep < p.n_blk && chn < ep.out_c) {
              rn0 = *reinterpret_cast<const uint4*>(rp + chn);
              if (chn + 8 < ep.out_c) rn1 = *reinterpret_cast<const uint4*>(rp + chn + 8);
            }
          }
          float v[16];
          tmem_ld16(trow + (uint32_t)c0, v);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ch = n0 + c0 + 8 * h;
            if (ch >= ep.out_c) continue;
            float x8[8];
            const float4 ba = *reinterpret_cast<const float4*>(ep.bias + ch), bb = *reinterpret_cast<const float4*>(ep.bias + ch + 4);
            x8[0] = v[8 * h] + ba.x; x8[1] = v[8 * h + 1] + ba.y; x8[2] = v[8 * h + 2] + ba.z; x8[3] = v[8 * h + 3] + ba.w;
            x8[4] = v[8 * h + 4] + bb.x; x8[5] = v[8 * h + 5] + bb.y; x8[6] = v[8 * h + 6] + bb.z; x8[7] = v[8 * h + 7] + bb.w;
            if (F & EF_RES) {
              const Vec8<T> t0 = *reinterpret_cast<const Vec8<T>*>(&rc[h]);
#pragma unroll
              for (int i = 0; i < 8; ++i) x8[i] += Op<T>::to_f(t0.v[i]);
            }
            if (F & EF_RAW) store8(reinterpret_cast<T*>(ep.out_raw) + off + ch, x8);
            if (F & EF_ACT) {
              float w8[8];
              if (ep.scale != nullptr) {
                const float4 sa = *reinterpret_cast<const float4*>(ep.scale + ch), sb = *reinterpret_cast<const float4*>(ep.scale + ch + 4);
                const float4 ta = *reinterpret_cast<const float4*>(ep.shift + ch), tb = *reinterpret_cast<const float4*>(ep.shift + ch + 4);
                w8[0] = fmaf(sa.x, x8[0], ta.x); w8[1] = fmaf(sa.y, x8[1], ta.y); w8[2] = fmaf(sa.z, x8[2], ta.z); w8[3] = fmaf(sa.w, x8[3], ta.w);
                w8[4] = fmaf(sb.x, x8[4], tb.x); w8[5] = fmaf(sb.y, x8[5], tb.y); w8[6] = fmaf(sb.z, x8[6], tb.z); w8[7] = fmaf(sb.w, x8[7], tb.w);
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) w8[i] = x8[i];
              }
              if (!ep.act_linear) {
#pragma unroll
                for (int i = 0; i < 8; ++i) w8[i] = act_fn(w8[i], ep.leak);
              }
              store8(reinterpret_cast<T*>(ep.out_act) + off + ch, w8);
            }
          }
        }
      } else
      for (int c0 = 16 * gi; c0 < p.n_blk; c0 += cstep) {
        float v[16];
        tmem_ld16(trow + (uint32_t)c0, v);
        if (valid) {
          if ((F & EF_GENERIC) != 0 && p.ep.mod_x != nullptr) {
            epilogue_ssm16<T>(p.ep, oy, ox, n0 + c0, v);
          } else {
            float a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = v[i]; b[i] = v[8 + i]; }
            epilogue8<T, F>(p.ep, oy, ox, n0 + c0, a);
            epilogue8<T, F>(p.ep, oy, ox, n0 + c0 + 8, b);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(bar_tempty + 8 * b);
      ITG_UACC(1, tl);
    }
    if (p.dbg && blockIdx.x == 0 && ew == 0 && lane == 0) { p.dbg[5 + 2 * g] = dacc[0]; p.dbg[6 + 2 * g] = dacc[1]; }
  }

  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();           // nobody exits while a peer may still multicast into / arrive on this CTA
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace itg
