// Per-patch self-attention block (Attention.forward, models/layers.py:246-258) on the tensor cores, 16-bit operands.
//
// One CTA = one 16 x 16 patch (256 query pixels, 64 pooled keys), 8 warps; warp w owns the two pixel rows 2w, 2w+1
// of the patch, i.e. 32 queries AND eight complete 2x2 pooling windows, so everything except the pooled keys / values
// stays in that warp's registers (flash-attention style chaining of m16n8k16 accumulators into the next A operand):
//
//   [theta | phi | g] = X Wqkv^T + b      256 x 96 x C      (theta 16 cols, phi 16, g 64; zero padded).  The scores are
//                                         exponentiated, so theta / phi keep ~fp32 accuracy: their weights, and then
//                                         theta / phi themselves, are carried as hi + lo pairs of the 16-bit type
//                                         (W_hi X + W_lo X;  S = th_hi ph_hi + th_hi ph_lo + th_lo ph_hi)
//   phi_p, g_p = maxpool2x2(phi), maxpool2x2(g)             in-register max + one shuffle, parked in shared memory
//   S = theta phi_p^T                     256 x 64 x 16     no 1/sqrt(d) scaling in the reference
//   P = softmax_rows(S)                   fp32, in registers
//   O1 = P g_p                            256 x 64 x 64
//   O2 = O1 Wo^T + bo                     256 x C x 64
//   out = gamma * O2 + x                  written in place over the staged x tile, then streamed out coalesced together
//                                         with act(bn(out)) (+ its replicate / zero frame) for the next conv.
//
// The attention block is 1 % of the path's FLOPs: mma.sync (HMMA) is used on purpose -- a 256 x 64 score tile per patch
// does not amortise a TMEM / tcgen05 pipeline; what matters is that nothing round-trips through HBM.
#pragma once
#include "itg_common.cuh"

namespace itg {

struct AttnMmaParams {
  const void* x;
  int th, tw, C, xc;
  const float *w_theta, *b_theta, *w_phi, *b_phi, *w_g, *b_g, *w_o, *b_o, *gamma;
  void* out_raw;
  void* out_act;
  const float* scale;
  const float* shift;
  float leak;
  int border;
};

constexpr int AM_PATCH = 16, AM_NPX = 256, AM_NPOOL = 64;
constexpr int AM_QKV = 128;                // 16 theta + 16 phi + 64 g columns + 32 low-order halves of the theta / phi weights
constexpr int AM_KMAX = 128;               // max channels
constexpr int AM_XP = AM_KMAX + 8;         // row pitch (elements) of the x / Wqkv tiles: +16 B keeps LDS.32 fragment loads conflict-free
constexpr int AM_OP = 64 + 8;              // row pitch of Wo / g_p^T tiles (K = 64)
constexpr int AM_PP = 16 + 8;              // row pitch of phi_p
constexpr int AM_SMEM = (AM_NPX * AM_XP + AM_QKV * AM_XP + AM_KMAX * AM_OP + 2 * AM_NPOOL * AM_PP + 64 * AM_OP) * 2 + (AM_QKV + AM_KMAX) * 4;

template <typename T> struct Pack2;
template <> struct Pack2<__half> {
  static __device__ __forceinline__ uint32_t pack(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
};
template <> struct Pack2<__nv_bfloat16> {
  static __device__ __forceinline__ uint32_t pack(float a, float b) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
};

template <typename T>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__half>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename T>
__global__ void __launch_bounds__(256, 1) attention_mma_kernel(const AttnMmaParams p) {
  extern __shared__ __align__(16) uint8_t am_smem[];
  T* Xs = reinterpret_cast<T*>(am_smem);                 // [256][AM_XP]   x tile, later out_raw in place
  T* Wq = Xs + AM_NPX * AM_XP;                           // [96][AM_XP]    theta | phi | g weights
  T* Wo = Wq + AM_QKV * AM_XP;                           // [128][AM_OP]   output 1x1 weights (K = C/2 padded to 64)
  T* Pp = Wo + AM_KMAX * AM_OP;                          // [2][64][AM_PP] pooled phi  [hi | lo][key j][channel]
  T* Gt = Pp + 2 * AM_NPOOL * AM_PP;                     // [64][AM_OP]    pooled g, transposed [channel][key j]
  float* bq = reinterpret_cast<float*>(Gt + 64 * AM_OP); // [96]
  float* bo = bq + AM_QKV;                               // [128]

  const int C = p.C, C8 = C >> 3, C2 = C >> 1, xc = p.xc;
  const int KT = (C + 15) >> 4;                          // k16 steps over the channels
  const int NT_OUT = (xc + 7) >> 3;                      // n8 tiles of the output (storage channels)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gr = lane >> 2, gc = (lane & 3) * 2;         // fragment row / column-pair of this lane
  const int H = p.th * AM_PATCH, W = p.tw * AM_PATCH;
  const T* xg = reinterpret_cast<const T*>(p.x);

  // ---------------- stage weights and biases once per CTA (CTAs are persistent over patches) ----------------
  pdl_launch_dependents();
  {
    const int kpad = KT * 16;
    for (int i = threadIdx.x; i < AM_QKV * kpad; i += 256) {
      const int n = i / kpad, k = i % kpad;
      float v = 0.f;
      const int nn = n >= 96 ? n - 96 : n;               // rows 96..127: low-order halves of rows 0..31
      if (k < C) {
        if (nn < 16) { if (nn < C8) v = p.w_theta[nn * C + k]; }
        else if (nn < 32) { if (nn - 16 < C8) v = p.w_phi[(nn - 16) * C + k]; }
        else if (nn - 32 < C2) v = p.w_g[(nn - 32) * C + k];
      }
      const T hi = Op<T>::from_f(v);
      Wq[n * AM_XP + k] = n >= 96 ? Op<T>::from_f(v - Op<T>::to_f(hi)) : hi;
    }
    for (int i = threadIdx.x; i < NT_OUT * 8 * 64; i += 256) {
      const int n = i >> 6, k = i & 63;
      Wo[n * AM_OP + k] = Op<T>::from_f((n < C && k < C2) ? p.w_o[n * C2 + k] : 0.f);
    }
    for (int i = threadIdx.x; i < AM_QKV; i += 256) {
      float v = 0.f;
      if (i < 16) { if (i < C8) v = p.b_theta[i]; }
      else if (i < 32) { if (i - 16 < C8) v = p.b_phi[i - 16]; }
      else if (i < 96 && i - 32 < C2) v = p.b_g[i - 32];
      bq[i] = v;
    }
    for (int i = threadIdx.x; i < AM_KMAX; i += 256) bo[i] = i < C ? p.b_o[i] : 0.f;
  }

  const int row0 = warp * 32;                            // this warp's first pixel
  pdl_wait();                                            // x (the previous launch's output) is complete from here on
  for (int pid = blockIdx.x; pid < p.th * p.tw; pid += gridDim.x) {
  const int pr = pid / p.tw, pc = pid % p.tw;
  __syncthreads();                                       // weights staged / previous patch fully streamed out
  // ---------------- stage the x tile of this patch ----------------
  {
    const int groups = KT * 2;                           // 8-channel groups per row incl. zero padding
    for (int i = threadIdx.x; i < AM_NPX * groups; i += 256) {
      const int px = i / groups, g8 = i % groups;
      const int y = pr * AM_PATCH + (px >> 4), x = pc * AM_PATCH + (px & 15);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (g8 * 8 < xc) v = *reinterpret_cast<const uint4*>(xg + grid_off(y, x, W, xc, g8 * 8));
      *reinterpret_cast<uint4*>(Xs + px * AM_XP + g8 * 8) = v;
    }
  }
  __syncthreads();

  // ---------------- 1. [theta | phi | g] = X Wqkv^T, 2. 2x2 max pooling ----------------
  // Pooling: m = 0 / 1 are the patch rows 2w / 2w+1 (same px); fragment rows gr, gr^1 are px pairs -> one shuffle
  // across lanes ^4; lanes with even gr then own window jx = gr/2 (c0,c1) and jx + 4 (c2,c3) of window row jy = w.
  uint32_t th_hi[2][4], th_lo[2][4];                     // theta as A operand of the score GEMM, hi + lo
  {
    // pass A: theta (tiles 0,1), phi (2,3) with hi weights, plus the same four tiles with the lo weights (rows 96..127)
    float acc[2][8][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
    for (int kt = 0; kt < KT; ++kt) {
      uint32_t a[2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const T* xr = Xs + (row0 + m * 16 + gr) * AM_XP + kt * 16 + gc;
        a[m][0] = *reinterpret_cast<const uint32_t*>(xr);
        a[m][1] = *reinterpret_cast<const uint32_t*>(xr + 8 * AM_XP);
        a[m][2] = *reinterpret_cast<const uint32_t*>(xr + 8);
        a[m][3] = *reinterpret_cast<const uint32_t*>(xr + 8 * AM_XP + 8);
      }
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int wrow = (n < 4 ? n * 8 : 96 + (n - 4) * 8) + gr;
        const T* wr = Wq + wrow * AM_XP + kt * 16 + gc;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wr), b1 = *reinterpret_cast<const uint32_t*>(wr + 8);
        mma16816<T>(acc[0][n], a[0], b0, b1);
        mma16816<T>(acc[1][n], a[1], b0, b1);
      }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {                        // full-precision theta / phi = hi part + lo part + bias
      const float b0 = bq[n * 8 + gc], b1 = bq[n * 8 + gc + 1];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        acc[m][n][0] += acc[m][n + 4][0] + b0; acc[m][n][1] += acc[m][n + 4][1] + b1;
        acc[m][n][2] += acc[m][n + 4][2] + b0; acc[m][n][3] += acc[m][n + 4][3] + b1;
      }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {                      // A regs: (tile 0: c0c1, c2c3), (tile 1: c0c1, c2c3)
        const float v0 = acc[m][r >> 1][(r & 1) * 2], v1 = acc[m][r >> 1][(r & 1) * 2 + 1];
        const float h0 = Op<T>::to_f(Op<T>::from_f(v0)), h1 = Op<T>::to_f(Op<T>::from_f(v1));
        th_hi[m][r] = Pack2<T>::pack(h0, h1);
        th_lo[m][r] = Pack2<T>::pack(v0 - h0, v1 - h1);
      }
    }
#pragma unroll
    for (int n = 2; n < 4; ++n) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i] = fmaxf(acc[0][n][i], acc[1][n][i]);
        v[i] = fmaxf(v[i], __shfl_xor_sync(0xffffffffu, v[i], 4));
      }
      if ((gr & 1) == 0) {
        const int j0 = warp * 8 + (gr >> 1), j1 = j0 + 4;
        const int c = (n - 2) * 8 + gc;
        float h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = Op<T>::to_f(Op<T>::from_f(v[i]));
        *reinterpret_cast<uint32_t*>(Pp + j0 * AM_PP + c) = Pack2<T>::pack(h[0], h[1]);
        *reinterpret_cast<uint32_t*>(Pp + j1 * AM_PP + c) = Pack2<T>::pack(h[2], h[3]);
        *reinterpret_cast<uint32_t*>(Pp + (AM_NPOOL + j0) * AM_PP + c) = Pack2<T>::pack(v[0] - h[0], v[1] - h[1]);
        *reinterpret_cast<uint32_t*>(Pp + (AM_NPOOL + j1) * AM_PP + c) = Pack2<T>::pack(v[2] - h[2], v[3] - h[3]);
      }
    }
  }
  {
    // pass B: g (weight rows 32..95) and its pooling
    float acc[2][8][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
    for (int kt = 0; kt < KT; ++kt) {
      uint32_t a[2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const T* xr = Xs + (row0 + m * 16 + gr) * AM_XP + kt * 16 + gc;
        a[m][0] = *reinterpret_cast<const uint32_t*>(xr);
        a[m][1] = *reinterpret_cast<const uint32_t*>(xr + 8 * AM_XP);
        a[m][2] = *reinterpret_cast<const uint32_t*>(xr + 8);
        a[m][3] = *reinterpret_cast<const uint32_t*>(xr + 8 * AM_XP + 8);
      }
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const T* wr = Wq + (32 + n * 8 + gr) * AM_XP + kt * 16 + gc;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wr), b1 = *reinterpret_cast<const uint32_t*>(wr + 8);
        mma16816<T>(acc[0][n], a[0], b0, b1);
        mma16816<T>(acc[1][n], a[1], b0, b1);
      }
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const float b0 = bq[32 + n * 8 + gc], b1 = bq[32 + n * 8 + gc + 1];
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i] = fmaxf(acc[0][n][i], acc[1][n][i]) + ((i & 1) ? b1 : b0);      // max(a + b, c + b) = max(a, c) + b
        v[i] = fmaxf(v[i], __shfl_xor_sync(0xffffffffu, v[i], 4));
      }
      if ((gr & 1) == 0) {
        const int j0 = warp * 8 + (gr >> 1), j1 = j0 + 4;
        const int c = n * 8 + gc;
        Gt[c * AM_OP + j0] = Op<T>::from_f(v[0]);
        Gt[(c + 1) * AM_OP + j0] = Op<T>::from_f(v[1]);
        Gt[c * AM_OP + j1] = Op<T>::from_f(v[2]);
        Gt[(c + 1) * AM_OP + j1] = Op<T>::from_f(v[3]);
      }
    }
  }
  __syncthreads();

  // ---------------- 3. S = theta phi_p^T, 4. row softmax ----------------
  float s[2][8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const T* pr_ = Pp + (n * 8 + gr) * AM_PP + gc;
    const uint32_t b0 = *reinterpret_cast<const uint32_t*>(pr_), b1 = *reinterpret_cast<const uint32_t*>(pr_ + 8);
    const uint32_t l0 = *reinterpret_cast<const uint32_t*>(pr_ + AM_NPOOL * AM_PP), l1 = *reinterpret_cast<const uint32_t*>(pr_ + AM_NPOOL * AM_PP + 8);
#pragma unroll
    for (int m = 0; m < 2; ++m) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s[m][n][i] = 0.f;
      mma16816<T>(s[m][n], th_lo[m], b0, b1);             // small terms first
      mma16816<T>(s[m][n], th_hi[m], l0, l1);
      mma16816<T>(s[m][n], th_hi[m], b0, b1);
    }
  }
  uint32_t pa[2][4][4];                                  // P as A operand: [m][k16 tile over keys][4]
#pragma unroll
  for (int m = 0; m < 2; ++m) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {                        // h = 0: fragment row gr (c0,c1); h = 1: row gr + 8 (c2,c3)
      float mx = -INFINITY;
#pragma unroll
      for (int n = 0; n < 8; ++n) mx = fmaxf(mx, fmaxf(s[m][n][2 * h], s[m][n][2 * h + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        s[m][n][2 * h] = __expf(s[m][n][2 * h] - mx);
        s[m][n][2 * h + 1] = __expf(s[m][n][2 * h + 1] - mx);
        sum += s[m][n][2 * h] + s[m][n][2 * h + 1];
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = 1.f / sum;
#pragma unroll
      for (int n = 0; n < 8; ++n) { s[m][n][2 * h] *= inv; s[m][n][2 * h + 1] *= inv; }
    }
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      pa[m][kt][0] = Pack2<T>::pack(s[m][2 * kt][0], s[m][2 * kt][1]);
      pa[m][kt][1] = Pack2<T>::pack(s[m][2 * kt][2], s[m][2 * kt][3]);
      pa[m][kt][2] = Pack2<T>::pack(s[m][2 * kt + 1][0], s[m][2 * kt + 1][1]);
      pa[m][kt][3] = Pack2<T>::pack(s[m][2 * kt + 1][2], s[m][2 * kt + 1][3]);
    }
  }

  // ---------------- 5. O1 = P g_p ----------------
  uint32_t oa[2][4][4];                                  // O1 as A operand of the output conv: [m][k16 tile over C/2][4]
  {
    float o1[2][8][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) o1[m][n][i] = 0.f;
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const T* gp = Gt + (n * 8 + gr) * AM_OP + kt * 16 + gc;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(gp), b1 = *reinterpret_cast<const uint32_t*>(gp + 8);
        mma16816<T>(o1[0][n], pa[0][kt], b0, b1);
        mma16816<T>(o1[1][n], pa[1][kt], b0, b1);
      }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        oa[m][kt][0] = Pack2<T>::pack(o1[m][2 * kt][0], o1[m][2 * kt][1]);
        oa[m][kt][1] = Pack2<T>::pack(o1[m][2 * kt][2], o1[m][2 * kt][3]);
        oa[m][kt][2] = Pack2<T>::pack(o1[m][2 * kt + 1][0], o1[m][2 * kt + 1][1]);
        oa[m][kt][3] = Pack2<T>::pack(o1[m][2 * kt + 1][2], o1[m][2 * kt + 1][3]);
      }
  }

  // ---------------- 6. O2 = O1 Wo^T + bo;  out = gamma * O2 + x, in place over the x tile ----------------
  const float gamma = p.gamma[0];
  for (int n = 0; n < NT_OUT; ++n) {
    float o2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      const T* wr = Wo + (n * 8 + gr) * AM_OP + kt * 16 + gc;
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wr), b1 = *reinterpret_cast<const uint32_t*>(wr + 8);
      mma16816<T>(o2[0], oa[0][kt], b0, b1);
      mma16816<T>(o2[1], oa[1][kt], b0, b1);
    }
    const float b0 = bo[n * 8 + gc], b1 = bo[n * 8 + gc + 1];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        T* xr = Xs + (row0 + m * 16 + gr + 8 * h) * AM_XP + n * 8 + gc;
        const float x0 = Op<T>::to_f(xr[0]), x1 = Op<T>::to_f(xr[1]);
        *reinterpret_cast<uint32_t*>(xr) = Pack2<T>::pack(fmaf(gamma, o2[m][2 * h] + b0, x0), fmaf(gamma, o2[m][2 * h + 1] + b1, x1));
      }
    }
  }
  __syncthreads();

  // ---------------- 7. stream out: out_raw and act(bn(out)) with its frame, 16 B per thread ----------------
  {
    const int groups = NT_OUT;
    for (int i = threadIdx.x; i < AM_NPX * groups; i += 256) {
      const int px = i / groups, g8 = i % groups;
      const int y = pr * AM_PATCH + (px >> 4), x = pc * AM_PATCH + (px & 15);
      float v[8];
      load8(Xs + px * AM_XP + g8 * 8, v);
      if (p.out_raw != nullptr) store8(reinterpret_cast<T*>(p.out_raw) + grid_off(y, x, W, xc, g8 * 8), v);
      if (p.out_act != nullptr) {
        float a8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float sc = p.scale ? p.scale[g8 * 8 + j] : 1.f;
          const float sh = p.shift ? p.shift[g8 * 8 + j] : 0.f;
          a8[j] = act_fn(fmaf(sc, v[j], sh), p.leak);
        }
        store8_framed(reinterpret_cast<T*>(p.out_act), y, x, H, W, xc, g8 * 8, a8, p.border);
      }
    }
  }
  }  // patch loop
}

}  // namespace itg
