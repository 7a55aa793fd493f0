// Microbenchmark: latency of one all-to-all activation exchange inside a thread-block cluster,
// the communication pattern of the decoder step (st.async + mbarrier::complete_tx).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dsmem_exchange dsmem_exchange.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void st_async1(uint32_t ra, float v, uint32_t rm) { asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(ra), "f"(v), "r"(rm) : "memory"); }
__device__ __forceinline__ void st_async4(uint32_t ra, float v, uint32_t rm) { asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%1,%1,%1}, [%2];" ::"r"(ra), "f"(v), "r"(rm) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(mb), "r"(parity) : "memory");
}
// mode 0: every warp sends one 4 B value to every peer (v3 pattern, S=1)
// mode 1: every warp sends one 16 B value to every peer (v3, S=4)
// mode 2: one warp per peer sends 16 B x 4 lanes (v2 pattern: 64 B block per source CTA)
// mode 3: barrier.cluster arrive.release + wait.acquire only
// mode 4: like 0 but only warp 0 sends (16 messages per CTA per round)
// mode 5: 64 B block staged in local smem, warp 0 lane p bulk-copies it to peer p (cp.async.bulk smem->dsmem)
// mode 6: like 5 with a 256 B block
// mode 7: like 2 with 256 B per source CTA (16 lanes x 16 B per peer)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* out) {
  __shared__ __align__(16) float buf[2][16 * 16 * 4];
  __shared__ __align__(16) float stg[2][64];
  __shared__ __align__(8) uint64_t mb[2];
  uint32_t nct; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(nct));
  uint32_t q; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(q));
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  if (threadIdx.x < 2) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mb[threadIdx.x])), "r"(1));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  long long t0 = clock64();
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    const int b = it & 1; const uint32_t par = (it >> 1) & 1;
    if (MODE == 3) {
      asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
      continue;
    }
    uint32_t bytes = MODE == 0 ? 16 * nct * 4 : (MODE == 1 ? 16 * nct * 16 : (MODE == 2 ? nct * 64 : nct * 4));
    if (MODE == 5) bytes = nct * 64;
    if (MODE == 6 || MODE == 7) bytes = nct * 256;
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mb[b])), "r"(bytes) : "memory");
    if (MODE == 0 || MODE == 1) {
      if (lane < (int)nct) {
        uint32_t ra = mapa(smem_u32(&buf[b][(q * 16 + wp) * 4]), lane), rm = mapa(smem_u32(&mb[b]), lane);
        if (MODE == 0) st_async1(ra, acc + it, rm); else st_async4(ra, acc + it, rm);
      }
    } else if (MODE == 2) {
      if (wp < (int)nct && lane < 4) {
        uint32_t ra = mapa(smem_u32(&buf[b][(q * 16 + lane) * 4]), wp), rm = mapa(smem_u32(&mb[b]), wp);
        st_async4(ra, acc + it, rm);
      }
    } else if (MODE == 5 || MODE == 6) {
      constexpr int NB = MODE == 5 ? 64 : 256;
      if (threadIdx.x < NB / 4) stg[b][threadIdx.x] = acc + it;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (wp == 0 && lane < (int)nct) {
        uint32_t ra = mapa(smem_u32(&buf[b][q * (NB / 4)]), lane), rm = mapa(smem_u32(&mb[b]), lane);
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(ra), "r"(smem_u32(&stg[b][0])), "r"(NB), "r"(rm) : "memory");
      }
    } else if (MODE == 7) {
      if (wp < (int)nct && lane < 16) {
        uint32_t ra = mapa(smem_u32(&buf[b][(q * 16 + lane) * 4]), wp), rm = mapa(smem_u32(&mb[b]), wp);
        st_async4(ra, acc + it, rm);
      }
    } else if (MODE == 4) {
      if (wp == 0 && lane < (int)nct) {
        uint32_t ra = mapa(smem_u32(&buf[b][q * 4]), lane), rm = mapa(smem_u32(&mb[b]), lane);
        st_async1(ra, acc + it, rm);
      }
    }
    mbar_wait(smem_u32(&mb[b]), par);
    acc += buf[b][(lane * 16 + wp) * 4 % (16 * 16 * 4)];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && q == 0 && blockIdx.x < 16) out[0] = t1 - t0;
  if (acc == 12345.678f) out[1] = 1;
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int MODE> void run(int cs, const char* name) {
  long long* d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
  const int iters = 2000;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(cs); cfg.blockDim = dim3(512);
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k<MODE>, iters, d);
  cudaError_t e2 = cudaDeviceSynchronize();
  long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("cluster %2d  %-44s : %7.1f cycles/exchange  (%s %s)\n", cs, name, (double)h[0] / iters, cudaGetErrorString(e), cudaGetErrorString(e2));
  cudaFree(d);
}
int main() {
  for (int cs : {16, 8, 4, 2}) {
    run<3>(cs, "barrier.cluster arrive.release/wait.acquire");
    run<4>(cs, "1 warp/CTA: 4 B to each peer");
    run<2>(cs, "CS warps/CTA: 64 B to one peer each");
    run<0>(cs, "16 warps/CTA: 4 B to each peer");
    run<1>(cs, "16 warps/CTA: 16 B to each peer");
    run<7>(cs, "CS warps/CTA: 256 B to one peer each");
    run<5>(cs, "bulk smem->dsmem: 64 B to each peer");
    run<6>(cs, "bulk smem->dsmem: 256 B to each peer");
  }
  return 0;
}
