// Hardware probes that size the decoder design (developer aid, not product code):
//   (a) how many 8-/16-CTA clusters of 512 threads can be co-resident for a given smem size
//   (b) per-SM L2 -> SM streaming bandwidth (LDG.128 into registers, cp.async.bulk into smem)
//       as a function of the number of SMs streaming at once
//   (c) legacy mma.sync m16n8k16 bf16 issue rate per SM
//   (d) one-way latency of st.async + mbarrier try_wait between two CTAs of a cluster (ping-pong)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o probe2 probe2.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) {
  uint32_t o;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r));
  return o;
}
__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra "
      "LAB_WAIT;\nDONE:\n}" ::"r"(mb),
      "r"(parity)
      : "memory");
}

// ---------------- (a) occupancy ----------------
__global__ void __launch_bounds__(512, 1) occ_kernel(float* p) {
  extern __shared__ float sm[];
  if (p) p[0] = sm[threadIdx.x];
}
static void probe_occupancy() {
  for (int cs : {8, 16}) {
    for (int kb : {32, 100, 160, 200, 227}) {
      cudaFuncSetAttribute(occ_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);
      cudaFuncSetAttribute(occ_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(cs);
      cfg.blockDim = dim3(512);
      cfg.dynamicSmemBytes = kb * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, occ_kernel, &cfg);
      printf("occupancy: cluster %2d, 512 thr, %3d KB smem -> max active clusters %d (%s)\n", cs, kb, n,
             cudaGetErrorString(e));
    }
  }
}

// ---------------- (b) L2 streaming bandwidth per SM ----------------
// each CTA re-reads its own `bytes` slice `iters` times
template <int UNROLL>
__global__ void __launch_bounds__(512, 1) stream_ldg(const float4* __restrict__ base, size_t slice_f4, int iters,
                                                     long long* clk, float* sink) {
  const float4* p = base + (size_t)blockIdx.x * slice_f4;
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (size_t i = threadIdx.x; i + (UNROLL - 1) * 512 < slice_f4; i += UNROLL * 512) {
      float4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) v[u] = __ldcg(p + i + u * 512);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  if (acc == 1.2345f) sink[0] = acc;
}

// bulk async copies global -> smem through a ring of NSTAGE chunks of CHUNK bytes; warps only wait.
__global__ void __launch_bounds__(512, 1) stream_bulk(const char* __restrict__ base, size_t slice_bytes, int chunk,
                                                      int nstage, int iters, long long* clk, float* sink) {
  extern __shared__ __align__(128) char ring[];
  __shared__ __align__(8) uint64_t full[8];
  const char* p = base + (size_t)blockIdx.x * slice_bytes;
  const int nchunk = (int)(slice_bytes / chunk);
  if (threadIdx.x < nstage) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[threadIdx.x])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int total = nchunk * iters;
  float acc = 0.f;
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int c = 0; c < nstage && c < total; ++c) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[c])), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(ring + (size_t)c * chunk)),
                   "l"(p + (size_t)(c % nchunk) * chunk), "r"(chunk), "r"(smem_u32(&full[c]))
                   : "memory");
    }
  }
  for (int c = 0; c < total; ++c) {
    const int s = c % nstage;
    mbar_wait(smem_u32(&full[s]), (c / nstage) & 1);
    // touch one word per thread so the data is really consumed
    acc += *reinterpret_cast<const float*>(ring + (size_t)s * chunk + (threadIdx.x * 4) % chunk);
    __syncthreads();   // everyone done with stage s
    if (threadIdx.x == 0 && c + nstage < total) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(ring + (size_t)s * chunk)),
                   "l"(p + (size_t)((c + nstage) % nchunk) * chunk), "r"(chunk), "r"(smem_u32(&full[s]))
                   : "memory");
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  if (acc == 1.2345f) sink[0] = acc;
}

static void probe_stream() {
  const size_t slice = 384 * 1024;   // ~ one CTA's share of the decoder weights at cluster size 16
  const int maxg = 148;
  char* d;
  cudaMalloc(&d, slice * maxg);
  cudaMemset(d, 0, slice * maxg);
  long long* clk;
  cudaMalloc(&clk, maxg * sizeof(long long));
  float* sink;
  cudaMalloc(&sink, 4);
  const int iters = 20;
  std::vector<long long> h(maxg);
  for (int g : {1, 16, 64, 128, 148}) {
    stream_ldg<8><<<g, 512>>>((const float4*)d, slice / 16, 2, clk, sink);   // warm L2
    stream_ldg<8><<<g, 512>>>((const float4*)d, slice / 16, iters, clk, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h.data(), clk, g * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < g; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("L2 stream LDG.128 x8 : %3d CTAs x 384 KB : %.1f B/clk/SM (slowest CTA), %.0f clk per pass  %s\n", g,
           (double)slice * iters / mx, (double)mx / iters, cudaGetErrorString(cudaGetLastError()));
    stream_ldg<16><<<g, 512>>>((const float4*)d, slice / 16, iters, clk, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h.data(), clk, g * sizeof(long long), cudaMemcpyDeviceToHost);
    mx = 0;
    for (int i = 0; i < g; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("L2 stream LDG.128 x16: %3d CTAs x 384 KB : %.1f B/clk/SM (slowest CTA), %.0f clk per pass  %s\n", g,
           (double)slice * iters / mx, (double)mx / iters, cudaGetErrorString(cudaGetLastError()));
    for (int chunk : {8192, 16384, 32768}) {
      const int nstage = 4;
      cudaFuncSetAttribute(stream_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk * nstage);
      stream_bulk<<<g, 512, chunk * nstage>>>(d, slice, chunk, nstage, iters, clk, sink);
      cudaDeviceSynchronize();
      cudaMemcpy(h.data(), clk, g * sizeof(long long), cudaMemcpyDeviceToHost);
      mx = 0;
      for (int i = 0; i < g; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("L2 stream bulk %5d B x%d stages: %3d CTAs x 384 KB : %.1f B/clk/SM, %.0f clk per pass  %s\n", chunk,
             nstage, g, (double)slice * iters / mx, (double)mx / iters, cudaGetErrorString(cudaGetLastError()));
    }
  }
  cudaFree(d); cudaFree(clk); cudaFree(sink);
}

// ---------------- (c) mma.sync rate ----------------
__global__ void __launch_bounds__(512, 1) mma_rate(int iters, long long* clk, float* sink) {
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 5, b1 = 11;
  float c[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)   // 4 independent accumulators per warp
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 1.2345f) sink[0] = s;
}
__global__ void __launch_bounds__(512, 1) ffma_rate(int iters, long long* clk, float* sink) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  const float x = 1.0001f, y = 0.5f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 1.2345f) sink[0] = s;
}
static void probe_mma() {
  long long* clk; cudaMalloc(&clk, 8);
  float* sink; cudaMalloc(&sink, 4);
  const int iters = 4000;
  long long h = 0;
  mma_rate<<<1, 512>>>(iters, clk, sink);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  // 16 warps x 4 mma per iteration; 4 SMSPs
  printf("mma.sync m16n8k16 bf16: %.2f clk per mma per SMSP (%.0f MAC/clk/SM)  %s\n", (double)h / (iters * 16.0),
         2048.0 * iters * 64.0 / h, cudaGetErrorString(cudaGetLastError()));
  ffma_rate<<<1, 512>>>(iters, clk, sink);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("FFMA: %.1f FMA/clk/SM  %s\n", 512.0 * 16.0 * iters / h, cudaGetErrorString(cudaGetLastError()));
  cudaFree(clk); cudaFree(sink);
}

// ---------------- (d) ping-pong latency over st.async ----------------
__global__ void __launch_bounds__(32, 1) pingpong(int iters, long long* out) {
  __shared__ __align__(16) float buf[4];
  __shared__ __align__(8) uint64_t mb;
  uint32_t q;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(q));
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mb)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const uint32_t peer = q ^ 1u;
  const uint32_t ra = mapa(smem_u32(buf), peer), rm = mapa(smem_u32(&mb), peer);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 4;" ::"r"(smem_u32(&mb)) : "memory");
    if (q == 0) {
      if (threadIdx.x == 0)
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(ra), "f"((float)it), "r"(rm) : "memory");
      mbar_wait(smem_u32(&mb), it & 1);
    } else {
      mbar_wait(smem_u32(&mb), it & 1);
      if (threadIdx.x == 0)
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(ra), "f"(buf[0] + 1.f), "r"(rm) : "memory");
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && q == 0) out[0] = t1 - t0;
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
static void probe_pingpong() {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 4000;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2); cfg.blockDim = dim3(32);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, pingpong, iters, d);
  cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("st.async ping-pong: %.1f clk one-way (store -> remote try_wait returns)  %s\n", (double)h / (2.0 * iters), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  probe_occupancy();
  probe_pingpong();
  probe_mma();
  probe_stream();
  return 0;
}
