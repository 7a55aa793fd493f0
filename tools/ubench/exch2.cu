// Microbenchmark 2: what bounds one all-to-all activation exchange in a 16-CTA cluster?
// Variants of the decoder's pattern "warp p sends this CTA's block to peer p" (st.async + complete_tx).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o exch2 exch2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void st_async4(uint32_t ra, float a, float b, float c, float d, uint32_t rm) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(ra), "f"(a), "f"(b), "f"(c), "f"(d), "r"(rm) : "memory");
}
__device__ __forceinline__ bool try_wait(uint32_t mb, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(ok) : "r"(mb), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t parity) { while (!try_wait(mb, parity)) {} }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// NT threads; each CTA sends F4 float4 (F4*16 bytes) to every peer per round.
// DEP: the value sent depends on data received in the previous round (the real recurrence).
// STAGE: values pass through local smem + __syncthreads before the sending warps pick them up.
// POLL1: one lane per warp polls the mbarrier, the rest wait at __syncwarp.
// NBUF: number of alternating buffers/mbarriers (2 = what the decoder does per phase pair).
template <int NT, int F4, bool DEP, bool STAGE, bool POLL1>
__global__ void __launch_bounds__(NT, 1) k(int iters, long long* out) {
  __shared__ __align__(16) float buf[2][16 * F4 * 4];
  __shared__ __align__(16) float stg[F4 * 4];
  __shared__ __align__(8) uint64_t mb[2];
  uint32_t nct; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(nct));
  uint32_t q; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(q));
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  constexpr int NWARP = NT / 32;
  if (threadIdx.x < 2) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mb[threadIdx.x])), "r"(1));
  for (int i = threadIdx.x; i < 2 * 16 * F4 * 4; i += NT) (&buf[0][0])[i] = 0.f;
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  cluster_sync();
  long long t0 = clock64();
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    const int b = it & 1; const uint32_t par = (it >> 1) & 1;
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mb[b])), "r"(nct * F4 * 16) : "memory");
    const float v = DEP ? acc + it : (float)it;
    if (STAGE) {
      if (threadIdx.x < F4 * 4) stg[threadIdx.x] = v;
      __syncthreads();
    }
    for (int p = wp; p < (int)nct; p += NWARP) {
      if (lane < F4) {
        float4 d = STAGE ? *reinterpret_cast<const float4*>(&stg[lane * 4]) : make_float4(v, v, v, v);
        st_async4(mapa(smem_u32(&buf[b][(q * F4 + lane) * 4]), p), d.x, d.y, d.z, d.w, mapa(smem_u32(&mb[b]), p));
      }
    }
    if (POLL1) {
      if (lane == 0) mbar_wait(smem_u32(&mb[b]), par);
      __syncwarp();
    } else {
      mbar_wait(smem_u32(&mb[b]), par);
    }
    acc += buf[b][(threadIdx.x * 4) % (16 * F4 * 4)];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && q == 0 && blockIdx.x < 16) out[0] = t1 - t0;
  if (acc == 12345.678f) out[1] = 1;
  cluster_sync();
}
template <int NT, int F4, bool DEP, bool STAGE, bool POLL1> void run(int cs, const char* name) {
  long long* d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
  const int iters = 2000;
  auto kern = k<NT, F4, DEP, STAGE, POLL1>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(cs); cfg.blockDim = dim3(NT);
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, iters, d);
  cudaError_t e2 = cudaDeviceSynchronize();
  long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("cluster %2d NT=%3d %4d B/peer dep=%d stage=%d poll1=%d %-12s: %7.1f cycles/exchange (%s %s)\n", cs, NT, F4 * 16, (int)DEP, (int)STAGE, (int)POLL1, name, (double)h[0] / iters, cudaGetErrorString(e), cudaGetErrorString(e2));
  cudaFree(d);
}

// k2: ONE warp (warp 0) sends F4 float4 to each of the nct peers (nct instructions, lanes < F4);
// PERSRC: 16 mbarriers per buffer, one per source CTA; warp w waits only for source w (w < nct), else for all.
template <int NT, int F4, bool PERSRC, int NSEND>
__global__ void __launch_bounds__(NT, 1) k2(int iters, long long* out) {
  __shared__ __align__(16) float buf[2][16 * F4 * 4];
  __shared__ __align__(8) uint64_t mb[2][16];
  uint32_t nct; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(nct));
  uint32_t q; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(q));
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  if (threadIdx.x < 32) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mb[threadIdx.x / 16][threadIdx.x % 16])), "r"(1));
  for (int i = threadIdx.x; i < 2 * 16 * F4 * 4; i += NT) (&buf[0][0])[i] = 0.f;
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  cluster_sync();
  long long t0 = clock64();
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    const int b = it & 1; const uint32_t par = (it >> 1) & 1;
    if (PERSRC) {
      if (threadIdx.x < nct) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mb[b][threadIdx.x])), "r"(F4 * 16) : "memory");
    } else {
      if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mb[b][0])), "r"(nct * F4 * 16) : "memory");
    }
    const float v = acc + it;
    if (wp < NSEND && lane < F4) {
      for (int p = wp; p < (int)nct; p += NSEND)
        st_async4(mapa(smem_u32(&buf[b][(q * F4 + lane) * 4]), p), v, v, v, v, mapa(smem_u32(&mb[b][PERSRC ? q : 0]), p));
    }
    if (PERSRC) {
      for (int s = 0; s < (int)nct; ++s) mbar_wait(smem_u32(&mb[b][(s + wp) % nct]), par);
      acc += buf[b][((wp % nct) * F4 * 4 + lane) % (16 * F4 * 4)];
    } else {
      mbar_wait(smem_u32(&mb[b][0]), par);
      acc += buf[b][(threadIdx.x * 4) % (16 * F4 * 4)];
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && q == 0 && blockIdx.x < 16) out[0] = t1 - t0;
  if (acc == 12345.678f) out[1] = 1;
  cluster_sync();
}
template <int NT, int F4, bool PERSRC, int NSEND> void run2(int cs) {
  long long* d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
  const int iters = 2000;
  auto kern = k2<NT, F4, PERSRC, NSEND>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(cs); cfg.blockDim = dim3(NT);
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, iters, d);
  cudaError_t e2 = cudaDeviceSynchronize();
  long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("k2 cluster %2d NT=%3d %4d B/peer persrc=%d senders=%2d: %7.1f cycles/exchange (%s %s)\n", cs, NT, F4 * 16, (int)PERSRC, NSEND, (double)h[0] / iters, cudaGetErrorString(e), cudaGetErrorString(e2));
  cudaFree(d);
}

__global__ void __launch_bounds__(512, 1) ffma2_rate(int iters, long long* clk, float* sink) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  const float x = 1.0001f, y = 0.5f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2)
      asm volatile("{\n.reg .b64 ra, rb, rc;\nmov.b64 ra, {%2, %2};\nmov.b64 rb, {%3, %3};\nmov.b64 rc, {%0, %1};\nfma.rn.f32x2 rc, rc, ra, rb;\nmov.b64 {%0, %1}, rc;\n}" : "+f"(a[i]), "+f"(a[i + 1]) : "f"(x), "f"(y));
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[0] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 1.2345f) sink[0] = s;
}
int main() {
  {
    long long* clk; cudaMalloc(&clk, 8); float* sink; cudaMalloc(&sink, 4);
    ffma2_rate<<<1, 512>>>(4000, clk, sink); cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("FFMA2 (fma.rn.f32x2): %.1f FMA/clk/SM  %s\n", 512.0 * 16.0 * 4000 / h, cudaGetErrorString(cudaGetLastError()));
  }
  for (int cs : {16, 8}) {
    run2<512, 16, false, 1>(cs); run2<512, 16, true, 1>(cs); run2<512, 16, false, 4>(cs); run2<512, 16, true, 4>(cs);
    run2<512, 16, false, 16>(cs); run2<512, 16, true, 16>(cs);
    run2<512, 20, false, 1>(cs); run2<512, 20, true, 1>(cs); run2<512, 20, true, 2>(cs);
    run2<512, 32, false, 1>(cs); run2<512, 32, true, 1>(cs); run2<512, 32, true, 2>(cs);
    run2<512, 4, false, 1>(cs); run2<512, 4, true, 1>(cs);
    run2<256, 16, true, 1>(cs); run2<256, 20, true, 1>(cs);
  }
  for (int cs : {16}) {
    run<512, 4, true, false, false>(cs, "");
    run<512, 4, false, false, false>(cs, "");
    run<512, 4, true, false, true>(cs, "");
    run<512, 4, true, true, false>(cs, "");
    run<512, 4, true, true, true>(cs, "");
    run<512, 16, true, false, false>(cs, "");
    run<512, 16, false, false, false>(cs, "");
    run<512, 16, true, false, true>(cs, "");
    run<512, 16, true, true, true>(cs, "");
    run<512, 32, true, true, true>(cs, "");
    run<128, 4, true, false, false>(cs, "");
    run<128, 16, true, false, false>(cs, "");
    run<32, 4, true, false, false>(cs, "");
    run<32, 16, true, false, false>(cs, "");
    run<32, 1, true, false, false>(cs, "");
  }
  return 0;
}
