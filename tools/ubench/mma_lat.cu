// mma.sync m16n8k16 bf16 latency: dependent chain in one warp, 1 or 3 interleaved chains, 1..16 warps per CTA.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
template <int CH>
__global__ void k(int iters, long long* out) {
  uint32_t a0 = threadIdx.x, a1 = 3, a2 = 7, a3 = 9, b0 = 5, b1 = 11;
  float c[CH][4];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  float s = 0; for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 1.2345f) out[1] = 1;
}
template <int CH> void run(int nthreads) {
  long long* d; cudaMalloc(&d, 16);
  const int iters = 2000;
  k<CH><<<1, nthreads>>>(iters, d); cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%2d warps, %d chains/warp: %.1f clk per iteration (= per dependent mma step), %s\n", nthreads / 32, CH, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main() {
  for (int nt : {32, 128, 256, 512}) { run<1>(nt); run<3>(nt); run<6>(nt); }
  return 0;
}
