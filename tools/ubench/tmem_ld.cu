// Probe: tcgen05.ld (TMEM -> registers) latency and throughput as a weight store for mma.sync kernels.
// Each warp owns a 128-column slice of its 32-lane quarter (16 chunk-tiles of 8 words per lane).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

__global__ void __launch_bounds__(512, 1) probe(long long* out, uint32_t* chk, int active_warps, int nld, int with_lds) {
  __shared__ uint32_t tbase_s;
  __shared__ uint4 junk[512];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  junk[tid] = make_uint4(tid, 1, 2, 3);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase_s;
  // lane quarter = warp % 4 (hardware rule), column slice = (warp / 4) * 128
  const uint32_t my = tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 128u;
  uint32_t r[8];
  for (int c = 0; c < 16; ++c) {
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = (uint32_t)(warp * 1000000 + c * 1000 + lane * 8 + k);
    tmem_st8(my + c * 8, r);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp < active_warps) {
    // latency: one dependent load
    t0 = clock64();
    tmem_ld8(my, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc += r[0] + r[7];
    t1 = clock64();
    // throughput: nld loads (cycling over the 16 chunk-tiles), one wait at the end of every group of 4
    for (int i = 0; i < nld; i += 4) {
      uint32_t a[8], b[8], c[8], d[8];
      tmem_ld8(my + ((i + 0) & 15) * 8, a);
      tmem_ld8(my + ((i + 1) & 15) * 8, b);
      tmem_ld8(my + ((i + 2) & 15) * 8, c);
      tmem_ld8(my + ((i + 3) & 15) * 8, d);
      if (with_lds) { const uint4 j = junk[(tid + i) & 511]; acc += j.x; }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += a[1] + b[2] + c[3] + d[4];
    }
    t2 = clock64();
  }
  if (lane == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = t2 - t1; }
  // verify content of chunk 5
  tmem_ld8(my + 5 * 8, r);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t bad = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) bad |= (r[k] != (uint32_t)(warp * 1000000 + 5 * 1000 + lane * 8 + k));
  if (bad) atomicAdd(chk, 1u);
  if (acc == 0xdeadbeef) chk[1] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int main() {
  long long* d; uint32_t* chk;
  cudaMalloc(&d, 32 * sizeof(long long)); cudaMalloc(&chk, 8); cudaMemset(chk, 0, 8);
  for (int lds = 0; lds < 2; ++lds)
    for (int aw : {1, 4, 8, 16}) {
      const int nld = 256;
      probe<<<1, 512>>>(d, chk, aw, nld, lds);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[32]; uint32_t c[2];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(c, chk, 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int w = 0; w < aw; ++w) mx = h[2 * w + 1] > mx ? h[2 * w + 1] : mx;
      printf("warps=%2d lds=%d: single ld+wait %lld clk; %d x (x8 = 1 KB/warp) in %lld clk -> %.1f clk per ld per warp, %.1f B/clk/SM; mismatches %u\n",
             aw, lds, h[0], nld, mx, (double)mx / nld, (double)aw * nld * 1024 / mx, c[0]);
    }
  return 0;
}
