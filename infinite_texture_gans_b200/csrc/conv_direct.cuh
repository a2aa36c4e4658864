// CUDA-core direct convolution on grid tensors (any operand dtype, fp32 accumulate).
//
// This is the exact (fp32) mode of the path and the on-device cross-check of the tcgen05 kernel: it takes the
// same itg_conv_desc, the same packed weights [tap][n_pad][k_pad] and runs the same epilogue, so the two
// implementations can be compared launch by launch on identical operands.
//
// One CTA = 128 output-grid pixels (a TH x TW tile) x NB GEMM columns.  Each thread owns one pixel and NB
// accumulators; weights of the current (tap, k-chunk) are staged in shared memory and read as broadcasts.
#pragma once
#include "itg_common.cuh"

namespace itg {

struct DirectParams {
  const void* in;
  int in_h, in_w, in_pitch, in_c, in_c_off, k;
  const void* w;
  int n_pad, k_pad;
  int mode;
  int tw_log2;          // tile width = 1 << tw_log2, tile height = 128 >> tw_log2
  int tiles_x;
  EpiParams ep;
};

constexpr int DIRECT_NB = 16;   // GEMM columns per CTA
constexpr int DIRECT_KC = 32;   // channels per staged weight chunk

template <typename T>
__global__ void __launch_bounds__(128) conv_direct_kernel(const DirectParams p) {
  __shared__ float w_s[DIRECT_NB][DIRECT_KC];

  const int tile = blockIdx.x;
  const int n0 = blockIdx.y * DIRECT_NB;
  const int phase = blockIdx.z;
  const int tw = 1 << p.tw_log2;
  const int tx = threadIdx.x & (tw - 1), ty = threadIdx.x >> p.tw_log2;
  const int y = (tile / p.tiles_x) * (128 >> p.tw_log2) + ty;
  const int x = (tile % p.tiles_x) * tw + tx;
  const bool valid = (y < p.in_h) && (x < p.in_w);

  const int ntaps = (p.mode == ITG_CONV3X3) ? 9 : (p.mode == ITG_CONV1X1 ? 1 : 4);
  const T* in = reinterpret_cast<const T*>(p.in);
  const T* w = reinterpret_cast<const T*>(p.w);

  float acc[DIRECT_NB];
#pragma unroll
  for (int i = 0; i < DIRECT_NB; ++i) acc[i] = 0.f;

  for (int t = 0; t < ntaps; ++t) {
    int dy, dx, wt;
    tap_offsets(p.mode, phase, t, dy, dx, wt);
    const T* a_ptr = in + grid_off_pitch(valid ? y + dy : 0, valid ? x + dx : 0, p.in_pitch, p.in_c, p.in_c_off);
    const T* w_tap = w + ((size_t)wt * p.n_pad + n0) * (size_t)p.k_pad;
    for (int k0 = 0; k0 < p.k; k0 += DIRECT_KC) {
      __syncthreads();
      for (int i = threadIdx.x; i < DIRECT_NB * DIRECT_KC; i += 128) {
        const int nn = i / DIRECT_KC, kk = i % DIRECT_KC;
        w_s[nn][kk] = (k0 + kk < p.k_pad) ? Op<T>::to_f(w_tap[(size_t)nn * p.k_pad + k0 + kk]) : 0.f;
      }
      __syncthreads();
      if (valid) {
#pragma unroll
        for (int kk = 0; kk < DIRECT_KC; kk += 8) {
          if (k0 + kk < p.k) {      // k and channel offsets are multiples of 8; padded channels hold zeros
            float a[8];
            load8(a_ptr + k0 + kk, a);
#pragma unroll
            for (int nn = 0; nn < DIRECT_NB; ++nn) {
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[nn] = fmaf(a[j], w_s[nn][kk + j], acc[nn]);
            }
          }
        }
      }
    }
  }

  if (!valid) return;
  int oy = y, ox = x;
  if (p.mode == ITG_UPCONV) { oy = 2 * y + (phase >> 1); ox = 2 * x + (phase & 1); }
  if (p.ep.mod_x != nullptr) {
    epilogue_ssm16<T>(p.ep, oy, ox, n0, acc);
    return;
  }
#pragma unroll
  for (int g = 0; g < DIRECT_NB; g += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = acc[g + i];
    epilogue8<T>(p.ep, oy, ox, n0 + g, v);
  }
}

}  // namespace itg
