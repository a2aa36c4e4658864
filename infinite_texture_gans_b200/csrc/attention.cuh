// Per-patch self-attention block (Attention.forward, models/layers.py:246-258) fused into one kernel:
//   theta = Wt x + bt          (npx x C/8)
//   phi   = maxpool2x2(Wp x + bp)   (npx/4 x C/8)
//   g     = maxpool2x2(Wg x + bg)   (npx/4 x C/2)
//   beta  = softmax_rows(theta phi^T)        -- no 1/sqrt(d) scaling in the reference
//   out   = gamma * (Wo (beta g) + bo) + x
// plus the producer-side BN + activation of the next block (out_act) and its replicate frame.
// One CTA per patch, one thread per pixel; scores / softmax / (beta g) stay in registers of the pixel's
// thread, the pooled keys and values live in shared memory.  fp32 math on CUDA cores (1% of the path's FLOPs).
#pragma once
#include "itg_common.cuh"

namespace itg {

struct AttnParams {
  const void* x;
  int th, tw, patch, C, xc;
  const float *w_theta, *b_theta, *w_phi, *b_phi, *w_g, *b_g, *w_o, *b_o, *gamma;
  void* out_raw;
  void* out_act;
  const float* scale;
  const float* shift;
  float leak;
  int border;
};

constexpr int ATT_C8 = 16;   // max C/8
constexpr int ATT_C2 = 64;   // max C/2

// NPOOL = (patch/2)^2 pooled positions (64 for the 16x16 patches of base_res 4)
template <typename T, int NPOOL>
__global__ void __launch_bounds__(NPOOL * 4) attention_kernel(const AttnParams p) {
  extern __shared__ float att_smem[];
  const int npx = NPOOL * 4;
  const int C = p.C, C8 = C / 8, C2 = C / 2;
  float* s_phi = att_smem;                       // [npx][ATT_C8]  full-res phi, then pooled in place region below
  float* s_g = s_phi + npx * ATT_C8;             // [npx][ATT_C2]
  float* s_phip = s_g + npx * ATT_C2;            // [NPOOL][ATT_C8]
  float* s_gp = s_phip + NPOOL * ATT_C8;         // [NPOOL][ATT_C2]

  const int patch = p.patch, half = patch / 2;
  const int pid = blockIdx.x;
  const int pr = pid / p.tw, pc = pid % p.tw;
  const int t = threadIdx.x;
  const int py = t / patch, px = t % patch;
  const int H = p.th * patch, W = p.tw * patch;
  const int y = pr * patch + py, x = pc * patch + px;
  const T* xin = reinterpret_cast<const T*>(p.x) + grid_off(y, x, W, p.xc, 0);

  // ---- 1x1 convs: theta / phi (pass 1) and g (pass 2) ----
  float theta[ATT_C8], phi[ATT_C8];
#pragma unroll
  for (int i = 0; i < ATT_C8; ++i) { theta[i] = (i < C8) ? p.b_theta[i] : 0.f; phi[i] = (i < C8) ? p.b_phi[i] : 0.f; }
  for (int k = 0; k < C; k += 8) {
    float xv[8];
    load8(xin + k, xv);
#pragma unroll
    for (int i = 0; i < ATT_C8; ++i) {
      if (i < C8) {
        const float4 a0 = *reinterpret_cast<const float4*>(p.w_theta + (size_t)i * C + k);
        const float4 a1 = *reinterpret_cast<const float4*>(p.w_theta + (size_t)i * C + k + 4);
        const float4 c0 = *reinterpret_cast<const float4*>(p.w_phi + (size_t)i * C + k);
        const float4 c1 = *reinterpret_cast<const float4*>(p.w_phi + (size_t)i * C + k + 4);
        theta[i] += xv[0] * a0.x + xv[1] * a0.y + xv[2] * a0.z + xv[3] * a0.w + xv[4] * a1.x + xv[5] * a1.y + xv[6] * a1.z + xv[7] * a1.w;
        phi[i] += xv[0] * c0.x + xv[1] * c0.y + xv[2] * c0.z + xv[3] * c0.w + xv[4] * c1.x + xv[5] * c1.y + xv[6] * c1.z + xv[7] * c1.w;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < ATT_C8; ++i) s_phi[t * ATT_C8 + i] = phi[i];
  {
    float g[ATT_C2];
#pragma unroll
    for (int i = 0; i < ATT_C2; ++i) g[i] = (i < C2) ? p.b_g[i] : 0.f;
    for (int k = 0; k < C; k += 8) {
      float xv[8];
      load8(xin + k, xv);
#pragma unroll
      for (int i = 0; i < ATT_C2; ++i) {
        if (i < C2) {
          const float4 a0 = *reinterpret_cast<const float4*>(p.w_g + (size_t)i * C + k);
          const float4 a1 = *reinterpret_cast<const float4*>(p.w_g + (size_t)i * C + k + 4);
          g[i] += xv[0] * a0.x + xv[1] * a0.y + xv[2] * a0.z + xv[3] * a0.w + xv[4] * a1.x + xv[5] * a1.y + xv[6] * a1.z + xv[7] * a1.w;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < ATT_C2; ++i) s_g[t * ATT_C2 + i] = g[i];
  }
  __syncthreads();

  // ---- 2x2 max pooling of phi and g (F.max_pool2d, layers.py:249-250) ----
  for (int e = t; e < NPOOL * (ATT_C8 + ATT_C2); e += npx) {
    const int j = e / (ATT_C8 + ATT_C2), c = e % (ATT_C8 + ATT_C2);
    const int jy = j / half, jx = j % half;
    const int p00 = (2 * jy) * patch + 2 * jx;
    if (c < ATT_C8) {
      const float* s = s_phi + c;
      s_phip[j * ATT_C8 + c] = fmaxf(fmaxf(s[p00 * ATT_C8], s[(p00 + 1) * ATT_C8]),
                                     fmaxf(s[(p00 + patch) * ATT_C8], s[(p00 + patch + 1) * ATT_C8]));
    } else {
      const int cc = c - ATT_C8;
      const float* s = s_g + cc;
      s_gp[j * ATT_C2 + cc] = fmaxf(fmaxf(s[p00 * ATT_C2], s[(p00 + 1) * ATT_C2]),
                                    fmaxf(s[(p00 + patch) * ATT_C2], s[(p00 + patch + 1) * ATT_C2]));
    }
  }
  __syncthreads();

  // ---- scores, softmax, beta * g : all in this pixel's thread ----
  float sc[NPOOL];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < NPOOL; ++j) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < ATT_C8; ++i) a = fmaf(theta[i], s_phip[j * ATT_C8 + i], a);   // padded entries are zero
    sc[j] = a;
    mx = fmaxf(mx, a);
  }
  float den = 0.f;
#pragma unroll
  for (int j = 0; j < NPOOL; ++j) { sc[j] = expf(sc[j] - mx); den += sc[j]; }
  const float inv = 1.f / den;
  float og[ATT_C2];
#pragma unroll
  for (int i = 0; i < ATT_C2; ++i) og[i] = 0.f;
#pragma unroll 4
  for (int j = 0; j < NPOOL; ++j) {
    const float b = sc[j] * inv;
#pragma unroll
    for (int i = 0; i < ATT_C2; ++i) og[i] = fmaf(b, s_gp[j * ATT_C2 + i], og[i]);
  }

  // ---- output 1x1 conv, residual, and the fused BN + activation of the consumer ----
  const float gamma = p.gamma[0];
  for (int c = 0; c < p.xc; c += 8) {
    float xv[8], o[8];
    load8(xin + c, xv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int cc = c + i;
      float a = 0.f;
      if (cc < C) {
        a = p.b_o[cc];
        const float* wr = p.w_o + (size_t)cc * C2;
#pragma unroll
        for (int q = 0; q < ATT_C2; ++q)
          if (q < C2) a = fmaf(wr[q], og[q], a);
      }
      o[i] = (cc < C) ? fmaf(gamma, a, xv[i]) : 0.f;
    }
    if (p.out_raw != nullptr) store8(reinterpret_cast<T*>(p.out_raw) + grid_off(y, x, W, p.xc, c), o);
    if (p.out_act != nullptr) {
      float a8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float s = p.scale ? p.scale[c + i] : 1.f;
        const float sh = p.shift ? p.shift[c + i] : 0.f;
        a8[i] = act_fn(fmaf(s, o[i], sh), p.leak);
      }
      store8_framed(reinterpret_cast<T*>(p.out_act), y, x, H, W, p.xc, c, a8, p.border);
    }
  }
}

}  // namespace itg
