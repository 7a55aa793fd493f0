// Microbenchmark: the context-slice loop of the decoder (few dependent LDS + FFMA trips per warp) with inline-asm
// shared loads vs plain C++ shared loads, 1..16 warps active.  Prints clocks per call.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float lds_f(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float4 lds_f4(uint32_t a) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ float lds_f_nv(uint32_t a) { float v; asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float4 lds_f4_nv(uint32_t a) { float4 v; asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* clk, int S, int T_in, int active_warps) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  float* sc = reinterpret_cast<float*>(smem);                 // [T_in][S]
  float* msl = reinterpret_cast<float*>(smem + 8192);         // [T_in][S][16]
  for (int i = tid; i < T_in * S; i += 512) sc[i] = 0.001f * i;
  for (int i = tid; i < T_in * S * 16; i += 512) msl[i] = 0.01f * (i & 255);
  __syncthreads();
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  long long best = 1 << 30;
  float4 acc = make_float4(0, 0, 0, 0), acc2 = acc;
  float ssum = 0, ssum2 = 0;
  for (int rep = 0; rep < 8; ++rep) {
    __syncthreads();
    const long long t0 = clock64();
    if (warp < active_warps && g < S) {
      const int NW = 16;
      if (MODE == 2) {
        const float* pa = sc + warp * S + g;
        const float* ma = msl + (warp * S + g) * 16 + t * 4;
        for (int j = warp; j + NW < T_in; j += 2 * NW) {
          const float p0v = pa[0], p1v = pa[NW * S];
          const float4 m0 = *reinterpret_cast<const float4*>(ma), m1 = *reinterpret_cast<const float4*>(ma + NW * S * 16);
          acc.x = fmaf(p0v, m0.x, acc.x); acc.y = fmaf(p0v, m0.y, acc.y); acc.z = fmaf(p0v, m0.z, acc.z); acc.w = fmaf(p0v, m0.w, acc.w);
          acc2.x = fmaf(p1v, m1.x, acc2.x); acc2.y = fmaf(p1v, m1.y, acc2.y); acc2.z = fmaf(p1v, m1.z, acc2.z); acc2.w = fmaf(p1v, m1.w, acc2.w);
          ssum += p0v; ssum2 += p1v;
          pa += 2 * NW * S; ma += 2 * NW * S * 16;
        }
      } else {
        uint32_t pa = sbase + (uint32_t)(warp * S + g) * 4u;
        uint32_t ma = sbase + 8192 + (uint32_t)((warp * S + g) * 16 + t * 4) * 4u;
        const uint32_t dp = (uint32_t)NW * S * 4u, dm = (uint32_t)NW * S * 64u;
        for (int j = warp; j + NW < T_in; j += 2 * NW) {
          float p0v, p1v; float4 m0, m1;
          if (MODE == 0) { p0v = lds_f(pa); p1v = lds_f(pa + dp); m0 = lds_f4(ma); m1 = lds_f4(ma + dm); }
          else { p0v = lds_f_nv(pa); p1v = lds_f_nv(pa + dp); m0 = lds_f4_nv(ma); m1 = lds_f4_nv(ma + dm); }
          acc.x = fmaf(p0v, m0.x, acc.x); acc.y = fmaf(p0v, m0.y, acc.y); acc.z = fmaf(p0v, m0.z, acc.z); acc.w = fmaf(p0v, m0.w, acc.w);
          acc2.x = fmaf(p1v, m1.x, acc2.x); acc2.y = fmaf(p1v, m1.y, acc2.y); acc2.z = fmaf(p1v, m1.z, acc2.z); acc2.w = fmaf(p1v, m1.w, acc2.w);
          ssum += p0v; ssum2 += p1v;
          pa += 2 * dp; ma += 2 * dm;
        }
      }
    }
    const long long t1 = clock64();
    if (t1 - t0 < best) best = t1 - t0;
  }
  out[tid] = acc.x + acc.y + acc.z + acc.w + acc2.x + acc2.y + acc2.z + acc2.w + ssum + ssum2;
  if (tid == 0) *clk = best;
}
int main() {
  float* out; long long* clk; cudaMalloc(&out, 512 * 4); cudaMalloc(&clk, 8);
  const int smem = 8192 + 100 * 8 * 64 + 1024;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int S : {1, 5, 8})
    for (int aw : {1, 4, 16})
      for (int mode = 0; mode < 3; ++mode) {
        if (mode == 0) k<0><<<1, 512, smem>>>(out, clk, S, 100, aw);
        if (mode == 1) k<1><<<1, 512, smem>>>(out, clk, S, 100, aw);
        if (mode == 2) k<2><<<1, 512, smem>>>(out, clk, S, 100, aw);
        long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        printf("S=%d warps=%2d mode=%d (%s): %lld clk  %s\n", S, aw, mode, mode == 0 ? "asm volatile" : mode == 1 ? "asm" : "c++", h, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
