// HBM-bound helpers around the GEMMs: batch-norm batch statistics, the
// (affine +) max-pool of the conv bank, in-place affine, decode step count.
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.cuh"

namespace taco {

namespace {

// ---- batch statistics (training-mode BN forward, reference modules.py:101) ----
// grid.x = channel blocks of 32, grid.y = row chunks.  block (32, 8).
__global__ void __launch_bounds__(256)
bn_partial_kernel(const float* __restrict__ x, int64_t x_bs, int ldx, int col_off, int N, int T,
                  int C, double* __restrict__ acc) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int M = N * T;
  float s = 0.f, ss = 0.f;
  if (c < C) {
    for (int m = blockIdx.y * 8 + threadIdx.y; m < M; m += gridDim.y * 8) {
      const int n = m / T, t = m - n * T;
      const float v = __ldg(x + (int64_t)n * x_bs + (int64_t)t * ldx + col_off + c);
      s += v;
      ss = fmaf(v, v, ss);
    }
  }
  __shared__ float sh[2][8][32];
  sh[0][threadIdx.y][threadIdx.x] = s;
  sh[1][threadIdx.y][threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double ds = 0.0, dss = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { ds += sh[0][i][threadIdx.x]; dss += sh[1][i][threadIdx.x]; }
    atomicAdd(acc + c, ds);
    atomicAdd(acc + C + c, dss);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ acc, int C, double inv_count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = acc[c] * inv_count;
  double var = acc[C + c] * inv_count - mean * mean;   // biased variance (tf.nn.moments)
  if (var < 0.0) var = 0.0;
  const double sc = (double)gamma[c] / sqrt(var + (double)eps);
  scale[c] = (float)sc;
  shift[c] = (float)((double)beta[c] - mean * sc);
}

// SPLIT: the pooled tensor only feeds the next tcgen05 GEMM, so it is written directly as bf16 hi + lo
// (x ~= hi + lo, the GEMM's operand format, dense [N*T][C]) instead of fp32 followed by a separate split pass.
template <bool SPLIT>
__global__ void __launch_bounds__(256)
affine_maxpool_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int T, int C4,
                      const float* __restrict__ scale, const float* __restrict__ shift,
                      uint2* __restrict__ hi, uint2* __restrict__ lo) {
  const int64_t total = (int64_t)N * T * C4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const int64_t row = i / C4;
    const int t = (int)(row % T);
    float4 a = reinterpret_cast<const float4*>(x)[i];
    float4 b = (t + 1 < T) ? reinterpret_cast<const float4*>(x)[i + C4] : a;
    if (scale) {
      const float4 s = ldg_f4(scale + c4 * 4), h = ldg_f4(shift + c4 * 4);
      a.x = fmaf(a.x, s.x, h.x); a.y = fmaf(a.y, s.y, h.y); a.z = fmaf(a.z, s.z, h.z); a.w = fmaf(a.w, s.w, h.w);
      b.x = fmaf(b.x, s.x, h.x); b.y = fmaf(b.y, s.y, h.y); b.z = fmaf(b.z, s.z, h.z); b.w = fmaf(b.w, s.w, h.w);
    }
    const float4 v = make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
    if (!SPLIT) {
      reinterpret_cast<float4*>(y)[i] = v;
    } else {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(v.x - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v.y - __bfloat162float(h1));
      const __nv_bfloat16 l2 = __float2bfloat16_rn(v.z - __bfloat162float(h2)), l3 = __float2bfloat16_rn(v.w - __bfloat162float(h3));
      auto pk = [](__nv_bfloat16 p, __nv_bfloat16 q2) { return (uint32_t)__bfloat16_as_ushort(p) | ((uint32_t)__bfloat16_as_ushort(q2) << 16); };
      hi[i] = make_uint2(pk(h0, h1), pk(h2, h3));
      lo[i] = make_uint2(pk(l0, l1), pk(l2, l3));
    }
  }
}

__global__ void __launch_bounds__(256)
affine_inplace_kernel(float* __restrict__ x, int64_t x_bs, int ldx, int N, int T, int C,
                      const float* __restrict__ scale, const float* __restrict__ shift,
                      const float* __restrict__ res, int64_t res_bs, int ldres) {
  const int64_t total = (int64_t)N * T * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t row = i / C;
    const int t = (int)(row % T);
    const int n = (int)(row / T);
    float* px = x + (int64_t)n * x_bs + (int64_t)t * ldx + c;
    float v = fmaf(*px, __ldg(scale + c), __ldg(shift + c));
    if (res) v += __ldg(res + (int64_t)n * res_bs + (int64_t)t * ldres + c);
    *px = v;
  }
}

// TacoTestHelper.next_inputs + dynamic_decode bookkeeping (reference
// models/helpers.py:35, Appendix B.3): a sample finishes at the first step whose
// 80*r outputs are all exactly 0.0; the loop ends when every sample has
// finished or at max_iters.  The decoder kernel always runs max_steps steps
// (finished samples keep computing in the reference too: impute_finished=False),
// so the loop length is recovered afterwards.
__global__ void find_steps_rows_kernel(const float* __restrict__ dec_out, int N, int max_steps, int D,
                                       int* __restrict__ first_fin) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= N * max_steps) return;
  const float* row = dec_out + (int64_t)warp * D;
  bool nz = false;
  for (int c = lane; c < D; c += 32) nz |= (row[c] != 0.0f);
  if (!__any_sync(0xffffffffu, nz) && lane == 0) atomicMin(first_fin + warp / max_steps, warp % max_steps);
}
__global__ void find_steps_init_kernel(int* first_fin, int N, int max_steps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) first_fin[i] = max_steps - 1;
}
__global__ void find_steps_final_kernel(const int* first_fin, int N, int* steps_out) {
  int mx = 0;
  for (int i = threadIdx.x; i < N; i += 32) mx = max(mx, first_fin[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (threadIdx.x == 0) *steps_out = mx + 1;
}

}  // namespace

void launch_bn_batch_stats(const float* x, int64_t x_bs, int ldx, int col_off, int N, int T, int C,
                           const float* gamma, const float* beta, float eps, double* acc,
                           float* scale_out, float* shift_out, cudaStream_t st) {
  cudaMemsetAsync(acc, 0, sizeof(double) * 2 * C, st);
  const int M = N * T;
  int chunks = (M + 63) / 64;
  if (chunks > 296) chunks = 296;
  if (chunks < 1) chunks = 1;
  dim3 grid((C + 31) / 32, chunks), block(32, 8);
  bn_partial_kernel<<<grid, block, 0, st>>>(x, x_bs, ldx, col_off, N, T, C, acc);
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(acc, C, 1.0 / (double)M, gamma, beta, eps,
                                                      scale_out, shift_out);
}

void launch_affine_maxpool(const float* x, float* y, int N, int T, int C, const float* scale,
                           const float* shift, cudaStream_t st) {
  const int64_t total = (int64_t)N * T * (C / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  affine_maxpool_kernel<false><<<blocks, 256, 0, st>>>(x, y, N, T, C / 4, scale, shift, nullptr, nullptr);
}

void launch_affine_maxpool_split(const float* x, void* hi, void* lo, int N, int T, int C, const float* scale,
                                 const float* shift, cudaStream_t st) {
  const int64_t total = (int64_t)N * T * (C / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  affine_maxpool_kernel<true><<<blocks, 256, 0, st>>>(x, nullptr, N, T, C / 4, scale, shift,
                                                      reinterpret_cast<uint2*>(hi), reinterpret_cast<uint2*>(lo));
}

void launch_affine_inplace(float* x, int64_t x_bs, int ldx, int N, int T, int C, const float* scale,
                           const float* shift, const float* res, int64_t res_bs, int ldres,
                           cudaStream_t st) {
  const int64_t total = (int64_t)N * T * C;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  affine_inplace_kernel<<<blocks, 256, 0, st>>>(x, x_bs, ldx, N, T, C, scale, shift, res, res_bs,
                                                ldres);
}

void launch_find_steps(const float* dec_out, int N, int max_steps, int D, int* first_fin,
                       int* steps_out, cudaStream_t st) {
  find_steps_init_kernel<<<(N + 127) / 128, 128, 0, st>>>(first_fin, N, max_steps);
  const int64_t threads = (int64_t)N * max_steps * 32;
  find_steps_rows_kernel<<<(int)((threads + 255) / 256), 256, 0, st>>>(dec_out, N, max_steps, D,
                                                                      first_fin);
  find_steps_final_kernel<<<1, 32, 0, st>>>(first_fin, N, steps_out);
}

}  // namespace taco
