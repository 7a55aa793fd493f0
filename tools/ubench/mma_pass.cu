// Probe: one operand pass of the decoder (6 chunks: LDS.128 B fragment + 3 HMMAs hi*hi, lo*hi, hi*lo) in several schedules.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v;
}
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(const uint4* __restrict__ w, long long* out, float* sink, int active, int reps) {
  __shared__ uint4 x[6 * 8 * 4 + 64];
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid < 6 * 8 * 4 + 64) x[tid] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3c003c00u, 0x3c003c00u);
  uint4 wb[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) wb[i] = w[(warp * 12 + i) * 32 + lane];
  __syncthreads();
  const uint32_t xaddr = (uint32_t)__cvta_generic_to_shared(x) + (lane >> 2) * 64 + (lane & 3) * 16;
  const uint32_t csb = 5 * 64;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < active) {
    long long tot = 0;
    for (int r = 0; r < reps; ++r) {
      __syncwarp();
      const long long ta = clock64();
      float hh[4] = {0, 0, 0, 0}, hl[4] = {0, 0, 0, 0}, lh[4] = {0, 0, 0, 0};
      if (MODE == 0) {   // load, 3 mma, load, 3 mma ... (double-buffered in source)
        uint4 xa = lds128(xaddr), xb;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          if (i + 1 < 6) { if (i & 1) xa = lds128(xaddr + (i + 1) * csb); else xb = lds128(xaddr + (i + 1) * csb); }
          const uint4& xf = (i & 1) ? xb : xa;
          mma(hh, wb[2 * i], xf.x, xf.y); mma(lh, wb[2 * i + 1], xf.x, xf.y); mma(hl, wb[2 * i], xf.z, xf.w);
        }
      } else if (MODE == 1) {   // all six B fragments first
        uint4 xf[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) xf[i] = lds128(xaddr + i * csb);
#pragma unroll
        for (int i = 0; i < 6; ++i) { mma(hh, wb[2 * i], xf[i].x, xf[i].y); mma(lh, wb[2 * i + 1], xf[i].x, xf[i].y); mma(hl, wb[2 * i], xf[i].z, xf[i].w); }
      } else {   // two accumulator sets (even / odd chunks), all B fragments first
        float h2[4] = {0, 0, 0, 0}, l2[4] = {0, 0, 0, 0}, m2[4] = {0, 0, 0, 0};
        uint4 xf[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) xf[i] = lds128(xaddr + i * csb);
#pragma unroll
        for (int i = 0; i < 6; i += 2) {
          mma(hh, wb[2 * i], xf[i].x, xf[i].y); mma(lh, wb[2 * i + 1], xf[i].x, xf[i].y); mma(hl, wb[2 * i], xf[i].z, xf[i].w);
          mma(h2, wb[2 * i + 2], xf[i + 1].x, xf[i + 1].y); mma(l2, wb[2 * i + 3], xf[i + 1].x, xf[i + 1].y); mma(m2, wb[2 * i + 2], xf[i + 1].z, xf[i + 1].w);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) { hh[q] += h2[q]; hl[q] += l2[q]; lh[q] += m2[q]; }
      }
      const float rsum = (hh[0] + hl[0] + lh[0]) + (hh[1] + hl[1] + lh[1]) + (hh[2] + hl[2] + lh[2]) + (hh[3] + hl[3] + lh[3]);
      // the pass is complete when its results are: keep them, make the next pass's operands depend on them
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(xaddr + 6 * 8 * 64), "f"(rsum) : "memory");
      acc += rsum;
      wb[0].x ^= (__float_as_uint(rsum) & 1u);
      const long long tb = clock64();
      tot += tb - ta;
    }
    t0 = 0; t1 = tot;
  }
  if (lane == 0) out[warp] = t1 - t0;
  if (acc == 1.2345f) *sink = acc;
}
int main() {
  uint4* w; long long* out; float* sink;
  cudaMalloc(&w, 16 * 12 * 32 * 16); cudaMemset(w, 0x3c, 16 * 12 * 32 * 16); cudaMalloc(&out, 16 * 8); cudaMalloc(&sink, 4);
  const int reps = 200;
  for (int mode = 0; mode < 3; ++mode)
    for (int active : {1, 4, 8, 16}) {
      if (mode == 0) k<0><<<1, 512>>>(w, out, sink, active, reps);
      if (mode == 1) k<1><<<1, 512>>>(w, out, sink, active, reps);
      if (mode == 2) k<2><<<1, 512>>>(w, out, sink, active, reps);
      cudaDeviceSynchronize();
      long long h[16]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < active; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("mode %d, %2d warps: %.1f clk per 6-chunk pass (18 mma) -> %.1f clk per chunk  %s\n", mode, active, (double)mx / reps, (double)mx / reps / 6, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
