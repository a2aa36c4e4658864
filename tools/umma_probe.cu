// Micro-benchmark of tcgen05.mma issue cost / latency / throughput for the thin-layer shapes (developer tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_probe tools/umma_probe.cu -lcuda
// One CTA per SM; warp 1 issues `batch` MMAs (M=128, N, K=16), commits, waits; repeated `reps` times.
// Variants: operand layout (no-swizzle 8x16B core matrices with halo-tile strides, or SW128), dependent
// accumulation into one TMEM tile vs round-robin over `nacc` independent tiles, wait per batch or only at the end.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "../infinite_texture_gans_b200/csrc/conv_tile.cuh"

using namespace itg;

struct ProbeParams { int n, batch, reps, nacc, layout, wait_each, a_stride16; uint32_t idesc; unsigned long long* out; };

__global__ void __launch_bounds__(128, 1) probe(const ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar = sbase, slot = sbase + 16, a_smem = sbase + 1024, b_smem = a_smem + 64 * 1024;
  for (uint32_t i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a_smem + i * 4), "r"(0x3c003c00u));
  fence_proxy_async();
  if (warp == 0 && lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 2) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  if (warp == 1) {
    const bool leader = elect_one_sync();
    uint32_t phase = 0;
    long long t0 = clock64();
    long long t_issue = 0;
    // descriptors precomputed; the batch loop is a plain unrolled issue of 4 MMAs per iteration
    uint64_t ad[4], bd[4];
    for (int i = 0; i < 4; ++i) {
      if (p.layout == 0) {
        ad[i] = desc_noswz((a_smem >> 4) + (uint32_t)i * (uint32_t)p.a_stride16, PLANE_BYTES / 16, HALO_W);
        bd[i] = desc_noswz((b_smem >> 4) + (uint32_t)i * 2u * (uint32_t)p.n, (uint32_t)p.n, 8);
      } else {
        ad[i] = make_smem_desc(a_smem + (uint32_t)i * 32u, 64, 2);
        bd[i] = make_smem_desc(b_smem + (uint32_t)i * 32u, 64, 2);
      }
    }
    const uint32_t acc1 = (p.nacc > 1) ? (uint32_t)p.n : 0u;
    for (int r = 0; r < p.reps; ++r) {
      const long long ti = clock64();
      for (int i = 0; i < p.batch; i += 4) {
        if (leader) {
          umma_f16(tmem_base, ad[0], bd[0], p.idesc, 1u);
          umma_f16(tmem_base + acc1, ad[1], bd[1], p.idesc, 1u);
          umma_f16(tmem_base, ad[2], bd[2], p.idesc, 1u);
          umma_f16(tmem_base + acc1, ad[3], bd[3], p.idesc, 1u);
        }
      }
      t_issue += clock64() - ti;
      if (p.wait_each || r == p.reps - 1) {
        if (leader) umma_commit(bar);
        __syncwarp();
        mbar_wait(bar, phase);
        phase ^= 1u;
      }
    }
    const long long t1 = clock64();
    if (lane == 0) { p.out[blockIdx.x * 2] = (unsigned long long)(t1 - t0); p.out[blockIdx.x * 2 + 1] = (unsigned long long)t_issue; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

int main() {
  unsigned long long* out;
  cudaMalloc(&out, 148 * 2 * sizeof(unsigned long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("layout n batch nacc wait_each | cycles/batch  cycles/mma  issue-cycles/mma\n");
  const int reps = 200;
  for (int layout = 0; layout < 2; ++layout)
    for (int n : {16, 32, 64, 128, 256})
      for (int nacc : {1, 2})
        for (int batch : {4, 8, 16, 64})
          for (int wait_each : {1, 0}) {
            if (nacc * n > 512) continue;
            if (layout == 1 && (nacc > 1) && n < 64) continue;
            ProbeParams p{n, batch, reps, nacc, layout, wait_each, 1, 0, out};
            p.idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
            probe<<<148, 128, 180 * 1024>>>(p);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            unsigned long long h[2];
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            printf("%s n=%3d batch=%2d nacc=%d wait_each=%d | %8.1f %8.1f %8.1f\n", layout ? "sw128" : "noswz", n, batch, nacc, wait_each,
                   (double)h[0] / reps, (double)h[0] / reps / batch, (double)h[1] / reps / batch);
          }
  return 0;
}
