// Developer probe: cost of a warp-level 16-byte cp.async (LDGSTS.128) as a function of the lane -> address mapping.
// Six warps per CTA, one CTA per SM, L2-resident source; prints cycles per warp instruction (per SM, all six warps issuing).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldgsts_probe tools/ldgsts_probe.cu && ./ldgsts_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void cpa(uint32_t dst, const void* src, int mode) {
  if (mode == 0) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  else if (mode == 1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(16) : "memory");      // zero-fill form
}

// map: 0 contiguous/contiguous, 1 four lanes per pixel (64 B of 208) -> planes, 2 lane = pixel (112 B stride) -> contiguous,
// 3 row-contiguous global -> padded planes (2896), 4 row-contiguous global -> planes (2880), 5 contiguous global + 16 B misaligned,
// 6 eight lanes per pixel (128 B of 208) -> padded planes, 7 eight lanes per pixel (128 B of 128: K = 64 storage) -> padded planes,
// 8 like 1 but padded planes, 9 contiguous global -> smem scattered 32 different 2896-planes
template <int MAP>
__global__ void __launch_bounds__(192, 1) probe(const uint8_t* __restrict__ g, size_t gbytes, int iters, int mode, int wrap, long long* out) {
  extern __shared__ uint8_t smem[];
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t base0 = ((size_t)blockIdx.x * 6 + warp) * ((gbytes / (148 * 6)) & ~(size_t)4095);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    size_t go; uint32_t so;
    const int ii = i & wrap;
    if (MAP == 0) { go = (size_t)ii * 512 + lane * 16; so = (uint32_t)((ii & 15) * 512 + lane * 16); }
    else if (MAP == 1 || MAP == 8) { const int px = ii * 8 + (lane >> 2); go = (size_t)px * 208 + (lane & 3) * 16; so = (uint32_t)((lane & 3) * (MAP == 8 ? 2896 : 2880) + (px % 180) * 16); }
    else if (MAP == 2) { const int px = ii * 32 + lane; go = (size_t)px * 112; so = (uint32_t)((px % 180) * 16); }
    else if (MAP == 3 || MAP == 4) { const int q = ii * 32 + lane; const int px = q / 13, c = q - px * 13; go = (size_t)q * 16; so = (uint32_t)(c * (MAP == 3 ? 2896 : 2880) + (px % 180) * 16); }
    else if (MAP == 5) { go = (size_t)ii * 512 + lane * 16 + 16; so = (uint32_t)((ii & 15) * 512 + lane * 16); }
    else if (MAP == 6) { const int px = ii * 4 + (lane >> 3); go = (size_t)px * 208 + (lane & 7) * 16; so = (uint32_t)((lane & 7) * 2896 + (px % 180) * 16); }
    else if (MAP == 7) { const int px = ii * 4 + (lane >> 3); go = (size_t)px * 128 + (lane & 7) * 16; so = (uint32_t)((lane & 7) * 2896 + (px % 180) * 16); }
    else { go = (size_t)ii * 512 + lane * 16; so = (uint32_t)(lane * 2896 + (ii % 180) * 16); }
    cpa(sb + warp * 0 + so, g + base0 + go, mode);
    if ((i & 7) == 7) { asm volatile("cp.async.commit_group;" ::: "memory"); asm volatile("cp.async.wait_group 2;" ::: "memory"); }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int MAP>
void run(const char* name, const uint8_t* g, size_t gbytes, long long* out) {
  for (int cfg = 0; cfg < 8; ++cfg) {
    const int mode = cfg % 3 == 2 ? 2 : cfg % 3, wrap = cfg < 3 ? 63 : 4095, smem = cfg < 6 ? 100 * 1024 : 219 * 1024;
    if (cfg == 7) continue;
    const int iters = 4096;
    cudaFuncSetAttribute(probe<MAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<MAP><<<148, 192, smem>>>(g, gbytes, iters, mode, wrap, out);
    probe<MAP><<<148, 192, smem>>>(g, gbytes, iters, mode, wrap, out);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < 148; ++i) s += (double)h[i];
    s /= 148;
    printf("%-62s %-9s %-9s smem %3d KB: %6.1f cycles per warp instruction at SM level, %5.1f B/cycle/SM  [%s]\n", name,
           mode == 0 ? ".ca" : (mode == 1 ? ".cg" : ".ca zfill"), wrap == 63 ? "L1-hit" : "streaming", smem / 1024, s / iters / 6, 6.0 * 512 * iters / s,
           cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  const size_t gbytes = 16ull << 30;      // every warp streams its own 4096 x (up to) 3584 B
  uint8_t* g; long long* out;
  cudaMalloc(&g, gbytes + 4096); cudaMemset(g, 1, gbytes + 4096); cudaMalloc(&out, 148 * sizeof(long long));
  run<0>("0 contiguous 512 B -> contiguous", g, gbytes, out);
//run<5>("5 contiguous 512 B (+16 B misaligned) -> contiguous", g, gbytes, out);
//run<9>("9 contiguous 512 B -> 32 padded planes", g, gbytes, out);
  run<1>("1 four lanes per pixel (64 of 208 B) -> 4 planes (2880)", g, gbytes, out);
//run<8>("8 four lanes per pixel (64 of 208 B) -> 4 padded planes (2896)", g, gbytes, out);
  run<2>("2 lane = pixel (16 of 112 B) -> contiguous", g, gbytes, out);
  run<3>("3 row-contiguous (13 chunks per pixel) -> 13 padded planes", g, gbytes, out);
//run<4>("4 row-contiguous (13 chunks per pixel) -> 13 planes (2880)", g, gbytes, out);
  run<6>("6 eight lanes per pixel (128 of 208 B) -> 8 padded planes", g, gbytes, out);
  run<7>("7 eight lanes per pixel (128 of 128 B) -> 8 padded planes", g, gbytes, out);
  return 0;
}
