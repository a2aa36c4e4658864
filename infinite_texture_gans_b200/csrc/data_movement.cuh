// Data movement around the conv / attention kernels: host noise -> grid tensors (utils.py:228, 246), halo moves of the sequential
// protocol (models/layers.py:103-143), the row-band multi-GPU halo exchange over peer-mapped memory, F.pad of the frame
// (models/layers.py:82) and the 8-bit output stage of test_sample.py:75-79.  All HBM-bound, 16 bytes per thread.
#pragma once
#include "itg_common.cuh"

namespace itg {

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
  long long timeout_cycles;   // how long a pull block waits for its neighbour before failing the launch
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
        if (clock64() - t0 > p.timeout_cycles) {     // a neighbour died (default ~60 s, ITG_HALO_TIMEOUT_S): fail the launch instead of hanging the GPU
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


// ------------------------------------------------------------------------------------------------
// counter-based noise: a window of a standard-normal FIELD, generated where it is consumed
// ------------------------------------------------------------------------------------------------
// utils.build_z / build_maps (utils.py:221-256) draw the full-grid noise with torch.randn on the host and ship it to the device; for
// a 65536^2 texture that is 2.2 GB of latents and seconds of host time per pass.  Here element e = (c * Hf + y) * Wf + x of field
// `stream` (0 = z, 1 + i = noise map of level i) is a pure function of (seed, stream, e): Philox4x32-10 (Salmon et al., SC'11; the
// Random123 reference known-answer vectors are checked in tests/) keyed by the seed, counter = (e / 4, stream), Box-Muller on the two
// pairs of outputs.  Any rank can therefore generate exactly its own band (or sub-image) of the same field, with no host work, no
// PCIe traffic and no exchange; it matches torch.randn in distribution only.
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__global__ void noise_normal_kernel(float* __restrict__ dst, int C, int h, int w, int y0, int x0, int Hf, int Wf, uint32_t seed_lo,
                                    uint32_t seed_hi, uint32_t stream) {
  const size_t total = (size_t)C * h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w), y = (int)((i / w) % h), c = (int)(i / ((size_t)w * h));
    const unsigned long long e = ((unsigned long long)c * Hf + (unsigned long long)(y0 + y)) * Wf + (unsigned long long)(x0 + x);
    const unsigned long long g = e >> 2;
    uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32), stream, 0u};
    philox4x32_10(ctr, seed_lo, seed_hi);
    const int pair = (int)(e & 2);                                    // outputs (0,1) serve elements 0,1 of the group, (2,3) elements 2,3
    const float u1 = ((float)ctr[pair] + 1.0f) * 2.3283064365386963e-10f;        // (0, 1]
    const float u2 = (float)ctr[pair + 1] * 2.3283064365386963e-10f;             // [0, 1]
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    dst[i] = rad * ((e & 1) ? sn : cs);
  }
}

}  // namespace itg
