// Shared device helpers: operand types, grid-tensor addressing, and the fused conv epilogue that both conv
// implementations (tcgen05 implicit GEMM and the CUDA-core direct conv) call with identical semantics.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/itg.h"

namespace itg {

// Programmatic dependent launch (PDL): kernels launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while their predecessor in the stream is still draining.  pdl_launch_dependents() lets the successor's CTAs
// be scheduled as soon as SM resources free up; pdl_wait() blocks until the predecessor grid has completed and its
// memory is visible -- everything a kernel does before pdl_wait() (barrier init, TMEM allocation, parking weights
// in shared memory) must not touch activations.  Both are no-ops for launches without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// operand types
// ---------------------------------------------------------------------------------------------------
template <typename T> struct Op;
template <> struct Op<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct Op<__half> {
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct Op<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// 8 consecutive channels of one pixel
template <typename T> struct Vec8 { T v[8]; };
template <> struct __align__(16) Vec8<__half> { __half v[8]; };
template <> struct __align__(16) Vec8<__nv_bfloat16> { __nv_bfloat16 v[8]; };
template <> struct __align__(16) Vec8<float> { float v[8]; };

// 256-bit global accesses (sm_100: LDG/STG.E.256): one 16-channel pixel of a 16-bit grid tensor per instruction
__device__ __forceinline__ void ldg256(const void* p, uint4& lo, uint4& hi) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t (&w)[8]) {
  asm volatile("st.global.v8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
               "r"(w[5]), "r"(w[6]), "r"(w[7]), "l"(p)
               : "memory");
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<float>(float a, float) { return __float_as_uint(a); }   // (unused: 16-bit tensors only)

template <typename T>
__device__ __forceinline__ void load8(const T* __restrict__ p, float (&f)[8]) {
  if constexpr (sizeof(T) == 2) {
    Vec8<T> t = *reinterpret_cast<const Vec8<T>*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = Op<T>::to_f(t.v[i]);
  } else {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
}

template <typename T>
__device__ __forceinline__ void store8(T* __restrict__ p, const float (&f)[8]) {
  if constexpr (sizeof(T) == 2) {
    Vec8<T> t;
#pragma unroll
    for (int i = 0; i < 8; ++i) t.v[i] = Op<T>::from_f(f[i]);
    *reinterpret_cast<Vec8<T>*>(p) = t;
  } else {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
}

// element offset of interior pixel (y, x), channel c of a framed grid tensor with interior width w
__device__ __forceinline__ size_t grid_off(int y, int x, int w, int c_store, int c) {
  return ((size_t)(y + 1) * (size_t)(w + 2) + (size_t)(x + 1)) * (size_t)c_store + (size_t)c;
}
// same, for a window into a wider buffer (`pitch` pixels per buffer row)
__device__ __forceinline__ size_t grid_off_pitch(int y, int x, int pitch, int c_store, int c) {
  return ((size_t)(y + 1) * (size_t)pitch + (size_t)(x + 1)) * (size_t)c_store + (size_t)c;
}

// Store 8 channels of interior pixel (y,x) and, for replicate outer padding, the frame pixels that mirror it
// (F.pad(..., 'replicate') of layers.py:82 applied once to the whole merged image).
template <typename T>
__device__ __forceinline__ void store8_framed(T* __restrict__ base, int y, int x, int h, int w, int c_store, int c,
                                              const float (&f)[8], int border) {
  store8(base + grid_off(y, x, w, c_store, c), f);
  if (border != ITG_BORDER_NONE) {
    const bool ey = (y == 0) | (y == h - 1), ex = (x == 0) | (x == w - 1);
    if (ey | ex) {
      const int fy = (y == 0) ? -1 : ((y == h - 1) ? h : y);
      const int fx = (x == 0) ? -1 : ((x == w - 1) ? w : x);
      float z[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) z[i] = (border == ITG_BORDER_REPLICATE) ? f[i] : 0.f;   // constant = zeros
      if (ey) store8(base + grid_off(fy, x, w, c_store, c), z);
      if (ex) store8(base + grid_off(y, fx, w, c_store, c), z);
      if (ey && ex) store8(base + grid_off(fy, fx, w, c_store, c), z);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// epilogue
// ---------------------------------------------------------------------------------------------------
struct EpiParams {
  int out_h, out_w, out_c;
  int n_pad;
  const float* bias;
  int res_kind, res_shift, res_c, res_h, res_w;
  const void* res;
  const void* mod_x;
  int mod_c, mod_shift, mod_h, mod_w;
  const float* mod_mean;
  const float* mod_rstd;
  void* out_raw;
  void* out_act;
  const float* scale;
  const float* shift;
  float leak;
  int act_linear;
  float* out_f32;
  float* out_img;
  int img_c, img_layout, patch;
  int border;
};

__device__ __forceinline__ float act_fn(float v, float leak) { return v >= 0.f ? v : v * leak; }

__device__ __forceinline__ float tanh_fast(float x) {     // MUFU.TANH, abs error ~5e-4: 16-bit paths only
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Compile-time description of what an epilogue instance does, so that the thin-layer kernel (where the
// per-pixel epilogue IS the cost) carries no dead branches.  EF_GENERIC = decide everything at run time.
enum : int { EF_RES = 1, EF_RAW = 2, EF_ACT = 4, EF_IMG = 8, EF_GENERIC = 1 << 20 };

// Epilogue for 8 consecutive GEMM columns [n, n+8) of output pixel (oy, ox).  acc = raw fp32 accumulators.
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// per-channel vector v[n .. n+8): from global memory, or (vec_smem != 0) from the CTA's shared-memory copy at
// vec_smem + slot * 256 bytes (slot 0 bias, 1 scale, 2 shift; 64 floats each)
__device__ __forceinline__ void load_vec8(const float* g, uint32_t vec_smem, int slot, int n, float4& a, float4& b) {
  if (vec_smem != 0) {
    a = lds_f4(vec_smem + (uint32_t)(slot * 256 + n * 4));
    b = lds_f4(vec_smem + (uint32_t)(slot * 256 + n * 4 + 16));
  } else {
    a = *reinterpret_cast<const float4*>(g + n);
    b = *reinterpret_cast<const float4*>(g + n + 4);
  }
}

// `pre`: the residual's 8 channels already fetched by the caller (16-bit operand types only), or nullptr.
// `vec_smem`: shared-memory address of the bias | scale | shift copy (tile kernel), 0 = read them from global memory.
template <typename T, int F = EF_GENERIC>
__device__ __forceinline__ void epilogue8(const EpiParams& ep, int oy, int ox, int n, float (&acc)[8],
                                          const uint4* pre = nullptr, uint32_t vec_smem = 0) {
  constexpr bool G = (F & EF_GENERIC) != 0;
  float v[8];
  if (ep.bias != nullptr) {
    float4 b0, b1;
    load_vec8(ep.bias, vec_smem, 0, n, b0, b1);
    v[0] = acc[0] + b0.x; v[1] = acc[1] + b0.y; v[2] = acc[2] + b0.z; v[3] = acc[3] + b0.w;
    v[4] = acc[4] + b1.x; v[5] = acc[5] + b1.y; v[6] = acc[6] + b1.z; v[7] = acc[7] + b1.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = acc[i];
  }

  if (G ? (ep.out_img != nullptr) : ((F & EF_IMG) != 0)) {     // final conv: tanh -> fp32 planar image
    if (n == 0) {
      for (int c = 0; c < ep.img_c && c < 8; ++c) {
        size_t o;
        if (ep.img_layout == ITG_IMG_PATCHES) {
          const int P = ep.patch, pw = ep.out_w / P;
          const int py = oy / P, px = ox / P;
          o = ((((size_t)(py * pw + px) * ep.img_c + c) * P) + (oy - py * P)) * P + (ox - px * P);
        } else {
          o = ((size_t)c * ep.out_h + oy) * (size_t)ep.out_w + ox;
        }
        ep.out_img[o] = (sizeof(T) == 2) ? tanh_fast(v[c]) : tanhf(v[c]);
      }
    }
    return;
  }

  if (G && ep.mod_x != nullptr) {                    // SSM: 8 columns = 4 channels (gamma, beta interleaved)
    const int c0 = n >> 1;
    if (c0 >= ep.out_c) return;
    const T* xp = reinterpret_cast<const T*>(ep.mod_x) +
                  grid_off(oy >> ep.mod_shift, ox >> ep.mod_shift, ep.mod_w, ep.mod_c, c0);
    float o4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float xh = (Op<T>::to_f(xp[i]) - ep.mod_mean[c0 + i]) * ep.mod_rstd[c0 + i];
      float y = (1.f + v[2 * i]) * xh + v[2 * i + 1];
      o4[i] = ep.act_linear ? y : act_fn(y, ep.leak);
    }
    // 4-channel store (+ replicate frame): done element-wise through an 8-wide helper would over-write; store 4
    T* ob = reinterpret_cast<T*>(ep.out_act);
    auto put4 = [&](int yy, int xx) {
      T* p = ob + grid_off(yy, xx, ep.out_w, ep.out_c, c0);
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = Op<T>::from_f(o4[i]);
    };
    put4(oy, ox);
    if (ep.border != ITG_BORDER_NONE) {
      const int h = ep.out_h, w = ep.out_w;
      const int fy = (oy == 0) ? -1 : ((oy == h - 1) ? h : oy);
      const int fx = (ox == 0) ? -1 : ((ox == w - 1) ? w : ox);
      const bool ey = (oy == 0) | (oy == h - 1), ex = (ox == 0) | (ox == w - 1);
      if (ep.border == ITG_BORDER_CONSTANT) {
#pragma unroll
        for (int i = 0; i < 4; ++i) o4[i] = 0.f;
      }
      if (ey) put4(fy, ox);
      if (ex) put4(oy, fx);
      if (ey && ex) put4(fy, fx);
    }
    return;
  }

  if (n >= ep.out_c) return;                          // padded GEMM columns beyond the stored channels

  if (G ? (ep.res_kind == ITG_RES_GRID) : ((F & EF_RES) != 0)) {
    float r[8];
    if (pre != nullptr && sizeof(T) == 2) {
      Vec8<T> t = *reinterpret_cast<const Vec8<T>*>(pre);
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = Op<T>::to_f(t.v[i]);
    } else {
      load8(reinterpret_cast<const T*>(ep.res) +
                grid_off(oy >> ep.res_shift, ox >> ep.res_shift, ep.res_w, ep.res_c, n), r);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += r[i];
  } else if (G && ep.res_kind == ITG_RES_F32) {
    const float* rp = reinterpret_cast<const float*>(ep.res) +
                      ((size_t)(oy >> ep.res_shift) * ep.res_w + (ox >> ep.res_shift)) * (size_t)ep.res_c + n;
    float r[8];
    load8(rp, r);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += r[i];
  }

  if (G ? (ep.out_raw != nullptr) : ((F & EF_RAW) != 0))
    store8(reinterpret_cast<T*>(ep.out_raw) + grid_off(oy, ox, ep.out_w, ep.out_c, n), v);
  if (G && ep.out_f32 != nullptr)
    store8(ep.out_f32 + ((size_t)oy * ep.out_w + ox) * (size_t)ep.out_c + n, v);
  if (G ? (ep.out_act != nullptr) : ((F & EF_ACT) != 0)) {
    float a[8];
    if (ep.scale != nullptr) {                       // scale and shift always come as a pair (BN eval fold)
      float4 s0, s1, t0, t1;
      load_vec8(ep.scale, vec_smem, 1, n, s0, s1);
      load_vec8(ep.shift, vec_smem, 2, n, t0, t1);
      a[0] = fmaf(s0.x, v[0], t0.x); a[1] = fmaf(s0.y, v[1], t0.y); a[2] = fmaf(s0.z, v[2], t0.z); a[3] = fmaf(s0.w, v[3], t0.w);
      a[4] = fmaf(s1.x, v[4], t1.x); a[5] = fmaf(s1.y, v[5], t1.y); a[6] = fmaf(s1.z, v[6], t1.z); a[7] = fmaf(s1.w, v[7], t1.w);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = v[i];
    }
    if (!ep.act_linear) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = act_fn(a[i], ep.leak);
    }
    store8_framed(reinterpret_cast<T*>(ep.out_act), oy, ox, ep.out_h, ep.out_w, ep.out_c, n, a, ep.border);
  }
}

// SSM epilogue for 16 consecutive GEMM columns [n, n+16) = 8 channels (gamma_c, beta_c interleaved) of pixel (oy, ox):
//   y_c = (1 + gamma_c) * (x_c - mean_c) * rstd_c + beta_c   [-> activation], one 16-byte load of x and one 16-byte store.
// StochasticSpatialModulation.forward, models/layers.py:228-234.
template <typename T>
__device__ __forceinline__ void epilogue_ssm16(const EpiParams& ep, int oy, int ox, int n, const float (&acc)[16]) {
  const int c0 = n >> 1;
  if (c0 >= ep.out_c) return;
  float g[16];
  if (ep.bias != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = *reinterpret_cast<const float4*>(ep.bias + n + 4 * q);
      g[4 * q] = acc[4 * q] + b.x; g[4 * q + 1] = acc[4 * q + 1] + b.y; g[4 * q + 2] = acc[4 * q + 2] + b.z; g[4 * q + 3] = acc[4 * q + 3] + b.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) g[i] = acc[i];
  }
  float x[8];
  load8(reinterpret_cast<const T*>(ep.mod_x) + grid_off(oy >> ep.mod_shift, ox >> ep.mod_shift, ep.mod_w, ep.mod_c, c0), x);
  const float4 m0 = *reinterpret_cast<const float4*>(ep.mod_mean + c0), m1 = *reinterpret_cast<const float4*>(ep.mod_mean + c0 + 4);
  const float4 r0 = *reinterpret_cast<const float4*>(ep.mod_rstd + c0), r1 = *reinterpret_cast<const float4*>(ep.mod_rstd + c0 + 4);
  const float mean[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
  const float rstd[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
  float y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float xh = (x[i] - mean[i]) * rstd[i];
    const float v = (1.f + g[2 * i]) * xh + g[2 * i + 1];
    y[i] = ep.act_linear ? v : act_fn(v, ep.leak);
  }
  store8_framed(reinterpret_cast<T*>(ep.out_act), oy, ox, ep.out_h, ep.out_w, ep.out_c, c0, y, ep.border);
}

// tap geometry shared by both conv kernels
struct TapGeom {
  int ntaps;        // taps per phase
  int nphase;       // 1 or 4
};

__device__ __forceinline__ void tap_offsets(int mode, int phase, int t, int& dy, int& dx, int& wt) {
  if (mode == ITG_CONV3X3) { dy = t / 3 - 1; dx = t % 3 - 1; wt = t; }
  else if (mode == ITG_CONV1X1) { dy = 0; dx = 0; wt = 0; }
  else { const int a = phase >> 1, b = phase & 1, i = t >> 1, j = t & 1; dy = a - 1 + i; dx = b - 1 + j; wt = phase * 4 + t; }
}

}  // namespace itg
