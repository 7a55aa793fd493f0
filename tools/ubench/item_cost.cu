// Microbenchmark: cost of one background "item" of decoder_cw (4 chunk-tiles of A fragments + 4 B fragments + 12 HMMA) by where
// the A fragments come from: (0) shared memory, (1) global/L2 prefetched one item ahead into registers, (2) tensor memory is not
// covered here.  W worker warps per CTA (others blocked at a barrier), 112 CTAs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v;
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v;
}
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* clk, const uint4* stream, int workers, int hmma) {
  __shared__ __align__(128) unsigned char smem[49152 - 256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 12000; i += 512) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  __syncthreads();
  float acc[3][4] = {};
  long long t0 = 0, t1 = 0;
  if (warp < workers) {
    const uint32_t sb = smem_u32(smem) + lane * 16;
    const uint4* src = stream + ((size_t)(blockIdx.x * 16 + warp) * 32) * 64 + lane;   // 32 chunk-tiles of 1 KB per warp
    uint4 wa[8], wn[8];
    int kf = 0;
    auto prefetch = [&]() {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        int idx = kf + kk; if (idx >= 28) idx -= 28;
        wn[2 * kk] = ldg_stream(src + (size_t)idx * 64);
        wn[2 * kk + 1] = ldg_stream(src + (size_t)idx * 64 + 32);
      }
    };
    if (MODE == 1) prefetch();
    for (int it = 0; it < 33; ++it) {
      if (it == 1) t0 = clock64();
      uint4 xf[4];
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) wa[i] = lds128(sb + (uint32_t)((it * 8 + i) % 56) * 512u);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) wa[i] = wn[i];
        kf += 4; if (kf >= 28) kf -= 28;
        prefetch();
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) xf[i] = lds128(sb + 28672u + (uint32_t)((it * 4 + i) % 36) * 512u);
      if (hmma) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          mma16816(acc[0], wa[2 * i], xf[i].x, xf[i].y);
          mma16816(acc[1], wa[2 * i + 1], xf[i].x, xf[i].y);
          mma16816(acc[2], wa[2 * i], xf[i].z, xf[i].w);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[0][i] += __uint_as_float(wa[2 * i].x ^ xf[i].x); acc[1][i] += __uint_as_float(wa[2 * i + 1].y ^ xf[i].z); }
      }
    }
    t1 = clock64();
    if (MODE == 1) asm volatile("" :: "r"(wn[0].x), "r"(wn[7].w));
  }
  __syncthreads();
  out[blockIdx.x * 512 + tid] = acc[0][0] + acc[1][1] + acc[2][2] + acc[0][3];
  if (blockIdx.x == 0 && warp == 0 && lane == 0) *clk = t1 - t0;
}
int main() {
  float* out; long long* clk; uint4* stream;
  cudaMalloc(&out, 112 * 512 * 4); cudaMalloc(&clk, 8);
  cudaMalloc(&stream, (size_t)112 * 16 * 32 * 1024); cudaMemset(stream, 0x3c, (size_t)112 * 16 * 32 * 1024);
  for (int hmma = 1; hmma >= 0; --hmma)
    for (int mode = 0; mode < 2; ++mode)
      for (int workers : {1, 4, 12})
        for (int rep = 0; rep < 2; ++rep) {
          if (mode == 0) k<0><<<112, 512>>>(out, clk, stream, workers, hmma); else k<1><<<112, 512>>>(out, clk, stream, workers, hmma);
          long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
          if (rep) printf("%s A from %-22s workers=%2d: %lld clk per item  %s\n", hmma ? "HMMA" : "no-MMA", mode ? "L2 (prefetch, registers)" : "shared memory", workers, h / 32, cudaGetErrorString(cudaGetLastError()));
        }
  return 0;
}
