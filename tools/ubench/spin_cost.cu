// Microbenchmark: what do warps that wait on an mbarrier cost the warps that work?  Worker warps (one per SM sub-partition)
// run a dependent LDS + HMMA chain; the other warps of the CTA wait (a) in mbarrier.try_wait loops, (b) in test_wait +
// nanosleep loops, (c) blocked at a named barrier.  Prints the worker's clocks.
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
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* clk, int workers) {
  __shared__ __align__(128) unsigned char smem[32768];
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 8192; i += 512) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  const uint32_t mb = smem_u32(&mbar);
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(workers * 32) : "memory");
  __syncthreads();
  const bool worker = warp >= 16 - workers;
  float acc[3][4] = {};
  long long t0 = 0, t1 = 0;
  if (worker) {
    const uint32_t sb = smem_u32(smem) + lane * 16;
    t0 = clock64();
    for (int it = 0; it < 16; ++it) {            // 16 "items": 4 A loads (hi, lo), 4 B loads, 12 HMMA each
      uint4 wa[8], xf[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) wa[i] = lds128(sb + (uint32_t)(it * 8 + i) * 512u % 32768u);
#pragma unroll
      for (int i = 0; i < 4; ++i) xf[i] = lds128(sb + (uint32_t)(it * 4 + i) * 512u % 32768u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        mma16816(acc[0], wa[2 * i], xf[i].x, xf[i].y);
        mma16816(acc[1], wa[2 * i + 1], xf[i].x, xf[i].y);
        mma16816(acc[2], wa[2 * i], xf[i].z, xf[i].w);
      }
    }
    t1 = clock64();
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb) : "memory");
    if (MODE == 2) asm volatile("bar.sync 1, 512;" ::: "memory");
  } else {
    if (MODE == 0) {
      asm volatile("{\n.reg .pred P1;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n@P1 bra DN;\nbra LW;\nDN:\n}" ::"r"(mb), "r"(0), "r"(0x989680u) : "memory");
    } else if (MODE == 1) {
      uint32_t done = 0;
      while (!done) {
        asm volatile("{\n.reg .pred P1;\nmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(done) : "r"(mb), "r"(0) : "memory");
        if (!done) __nanosleep(100);
      }
    } else if (MODE == 2) {
      asm volatile("bar.sync 1, 512;" ::: "memory");
    } else {
      asm volatile("{\n.reg .pred P1;\nLW2:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DN2;\nbra LW2;\nDN2:\n}" ::"r"(mb), "r"(0) : "memory");
    }
  }
  __syncthreads();
  out[tid] = acc[0][0] + acc[1][1] + acc[2][2];
  if (warp == 15 && lane == 0) *clk = t1 - t0;
}
int main() {
  float* out; long long* clk; cudaMalloc(&out, 512 * 4); cudaMalloc(&clk, 8);
  const char* names[4] = {"try_wait + suspend hint (spin)", "test_wait + nanosleep(100)", "blocked at a named barrier", "try_wait, no hint (spin)"};
  for (int workers : {1, 4})
    for (int mode = 0; mode < 4; ++mode)
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, 512>>>(out, clk, workers);
        if (mode == 1) k<1><<<1, 512>>>(out, clk, workers);
        if (mode == 2) k<2><<<1, 512>>>(out, clk, workers);
        if (mode == 3) k<3><<<1, 512>>>(out, clk, workers);
        long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("workers=%d others: %-34s: %lld clk per 16 items (%lld per item)  %s\n", workers, names[mode], h, h / 16, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
