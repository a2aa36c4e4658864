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

// ------------------------------------------------------------------------------------------------
// small data-movement kernels
// ------------------------------------------------------------------------------------------------
// fp32 planar (C,H,W) -> channels-last (H,W,dst_c) in T, zero-filled channel tail.  One thread per (pixel, 8 ch).
template <typename T>
__global__ void pack_nchw_kernel(const float* __restrict__ src, int C, int H, int W, T* __restrict__ dst, int dst_c) {
  const size_t groups = (size_t)dst_c / 8;
  const size_t total = (size_t)H * W * groups;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i % ((size_t)H * W);      // pixel fastest: coalesced reads of each channel plane
    const int g = (int)(i / ((size_t)H * W));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      v[j] = (c < C) ? src[(size_t)c * H * W + pix] : 0.f;
    }
    store8(dst + pix * dst_c + (size_t)g * 8, v);
  }
}

// SSM noise map -> 3x3 tap stack: src fp32 (Hm, Wm) single channel; dst framed grid tensor with interior
// (Hm-2) x (Wm-2) and dst_c >= 9 channels: dst(y, x)[t] = src[y + t/3][x + t%3], channels >= 9 zero.
// The 1 -> 128 `mlp_shared` conv (layers.py:220) then runs as a K=16 1x1 GEMM on the tensor cores.
template <typename T>
__global__ void pack_map_taps_kernel(const float* __restrict__ src, int Hm, int Wm, T* __restrict__ dst, int dst_c) {
  const int h = Hm - 2, w = Wm - 2;
  const size_t groups = (size_t)dst_c / 8;
  const size_t total = (size_t)h * w * groups;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const size_t pix = i / groups;
    const int x = (int)(pix % w), y = (int)(pix / w);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = g * 8 + j;
      v[j] = (t < 9) ? src[(size_t)(y + t / 3) * Wm + x + t % 3] : 0.f;
    }
    store8(dst + grid_off(y, x, w, dst_c, g * 8), v);
  }
}

template <typename T>
__global__ void copy_rect_kernel(const T* __restrict__ src, int src_pitch, int sy, int sx, T* __restrict__ dst,
                                 int dst_pitch, int dy, int dx, int h, int w, int c) {
  const size_t groups = (size_t)c / 8;
  const size_t total = (size_t)h * w * groups;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const size_t pix = i / groups;
    const int xx = (int)(pix % w), yy = (int)(pix / w);
    const Vec8<T>* s = reinterpret_cast<const Vec8<T>*>(src + ((size_t)(sy + yy) * src_pitch + sx + xx) * c) + g;
    Vec8<T>* d = reinterpret_cast<Vec8<T>*>(dst + ((size_t)(dy + yy) * dst_pitch + dx + xx) * c) + g;
    *d = *s;
  }
}

// ------------------------------------------------------------------------------------------------
// row-band multi-GPU halo exchange over peer-mapped memory (NVLink P2P), no host involvement
// ------------------------------------------------------------------------------------------------
// One launch per conv2d_lp input: blocks 0 / 1 PUSH this rank's first / last interior pixel row (frame columns
// included) into the up / down neighbour's inbox and then publish the step number in the neighbour's flag
// (system-scope release); blocks 2 / 3 wait (bounded spin, system-scope acquire) for the up / down neighbour's flag to
// reach the step number and PULL the inbox row into this rank's top / bottom frame row.  Inboxes and flags are one per
// halo point, so a neighbour can never overwrite a row that has not been consumed.
struct HaloXchgParams {
  void* grid;            // (h+2) x (w+2) x c framed grid tensor of this rank
  int h, w, c;
  void* up_inbox;        // peer pointers (NULL at the first / last band): neighbour's bottom / top inbox row
  void* down_inbox;
  int* up_flag;          // peer pointers: neighbour's "bottom arrived" / "top arrived" flags
  int* down_flag;
  const void* top_inbox; // local inbox rows written by the neighbours
  const void* bot_inbox;
  int* top_flag;         // local flags
  int* bot_flag;
  const int* step;       // device-resident step counter (advanced once per Generator pass)
  int roles;             // bit 0 push up, bit 1 push down, bit 2 pull top, bit 3 pull bottom
};

__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <typename T>
__global__ void __launch_bounds__(1024) halo_xchg_kernel(const HaloXchgParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int role = blockIdx.x;                       // 0 push up, 1 push down, 2 pull top, 3 pull bottom
  if (!((p.roles >> role) & 1)) return;
  const size_t row_elems = (size_t)(p.w + 2) * p.c;
  const size_t chunks = row_elems / 8;               // 16-byte chunks (c is a multiple of 8)
  T* g = reinterpret_cast<T*>(p.grid);
  const int step = *p.step;
  if (role < 2) {
    T* dst = reinterpret_cast<T*>(role == 0 ? p.up_inbox : p.down_inbox);
    int* flag = role == 0 ? p.up_flag : p.down_flag;
    if (dst == nullptr) return;
    const T* src = g + (size_t)(role == 0 ? 1 : p.h) * row_elems;
    for (size_t i = threadIdx.x; i < chunks; i += blockDim.x)
      reinterpret_cast<Vec8<T>*>(dst)[i] = reinterpret_cast<const Vec8<T>*>(src)[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(flag, step);
  } else {
    const T* src = reinterpret_cast<const T*>(role == 2 ? p.top_inbox : p.bot_inbox);
    int* flag = role == 2 ? p.top_flag : p.bot_flag;
    if (src == nullptr) return;
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      while (ld_acquire_sys(flag) < step) {
        if (clock64() - t0 > 8000000000LL) {         // ~4 s: a neighbour died; fail the launch instead of hanging the GPU
          printf("itg: halo exchange timed out waiting for step %d (flag %d)\n", step, ld_acquire_sys(flag));
          __trap();
        }
      }
    }
    __syncthreads();
    __threadfence_system();
    T* dst = g + (size_t)(role == 2 ? 0 : p.h + 1) * row_elems;
    for (size_t i = threadIdx.x; i < chunks; i += blockDim.x)
      reinterpret_cast<Vec8<T>*>(dst)[i] = reinterpret_cast<const Vec8<T>*>(src)[i];
  }
}

__global__ void step_advance_kernel(int* step) { *step += 1; }

// test_sample.py:78 + torchvision save_image: uint8 = trunc(clamp((x * 0.5 + 0.5) * 255 + 0.5, 0, 255)), every step rounded to fp32
__global__ void image_to_u8_kernel(const float* __restrict__ img, int c, int h, int w, long long row_pitch, long long plane_pitch,
                                   uint8_t* __restrict__ out) {
  const size_t total = (size_t)h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / w), x = (int)(i - (size_t)y * w);
    const float* px = img + (size_t)y * row_pitch + x;
    uint8_t* o = out + i * c;
    for (int k = 0; k < c; ++k) {
      float v = __fadd_rn(__fmul_rn(px[(size_t)k * plane_pitch], 0.5f), 0.5f);
      v = __fadd_rn(__fmul_rn(v, 255.f), 0.5f);
      v = fminf(fmaxf(v, 0.f), 255.f);
      o[k] = (uint8_t)(int)v;                              // .to(uint8) truncates
    }
  }
}

// F.pad(x, (1,1,1,1), mode) on the frame of a grid tensor; sides: bit0 top, bit1 bottom, bit2 left, bit3 right
template <typename T>
__global__ void fill_frame_kernel(T* __restrict__ t, int h, int w, int c, int border, int sides) {
  const size_t groups = (size_t)c / 8;
  const int per = 2 * (w + 2) + 2 * h;           // frame pixels: top row, bottom row (with corners), left, right
  const size_t total = (size_t)per * groups;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    int q = (int)(i / groups);
    int fy, fx;
    if (q < w + 2) { fy = -1; fx = q - 1; }
    else if (q < 2 * (w + 2)) { fy = h; fx = q - (w + 2) - 1; }
    else if (q < 2 * (w + 2) + h) { fy = q - 2 * (w + 2); fx = -1; }
    else { fy = q - 2 * (w + 2) - h; fx = w; }
    const bool top = fy < 0, bot = fy >= h, lef = fx < 0, rig = fx >= w;
    // a frame pixel is written if every side it lies on is enabled
    if ((top && !(sides & 1)) || (bot && !(sides & 2)) || (lef && !(sides & 4)) || (rig && !(sides & 8))) continue;
    const int cy = min(max(fy, 0), h - 1), cx = min(max(fx, 0), w - 1);
    Vec8<T>* d = reinterpret_cast<Vec8<T>*>(t + grid_off(fy, fx, w, c, 0)) + g;
    if (border == ITG_BORDER_REPLICATE) {
      *d = *(reinterpret_cast<const Vec8<T>*>(t + grid_off(cy, cx, w, c, 0)) + g);
    } else {
      Vec8<T> z;
#pragma unroll
      for (int j = 0; j < 8; ++j) z.v[j] = Op<T>::from_f(0.f);
      *d = z;
    }
  }
}

}  // namespace itg
