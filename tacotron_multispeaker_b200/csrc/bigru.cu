// K6b: recurrence of the CBHG bidirectional GRU as a persistent kernel.
//
// Replaces tf.nn.bidirectional_dynamic_rnn(GRUCell(128), GRUCell(128), x,
// sequence_length) at reference models/modules.py:68-74 (a tf.while_loop of
// T_in resp. T_out=1000 iterations per direction).  The input halves of the
// GRU kernels are hoisted into one dense GEMM (xproj, conv_gemm.cu); what is
// left per step and direction is
//     [r|u] = sigmoid(xg_t + h U_g)      U_g [128,256]
//     c     = tanh  (xc_t + (r*h) U_c)   U_c [128,128]   (reset BEFORE matmul: TF GRUCell)
//     h'    = u*h + (1-u)*c
// One CTA owns (direction, NS samples) for the whole sequence: the 49,152
// recurrent weights live in REGISTERS (96 per thread x 512 threads) for all
// steps, h lives in shared memory, nothing crosses CTAs, and the hoisted
// projections stream in through a 4-deep cp.async ring.
#include "common.cuh"
#include "kernels.cuh"

namespace taco {

namespace {
constexpr int H = 128;        // GRU width
constexpr int XW = 768;       // xproj row: [fw gates 256 | fw cand 128 | bw gates 256 | bw cand 128]
constexpr int RING = 4;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int Nw> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(Nw) : "memory");
}

template <int NS>
__global__ void __launch_bounds__(512, 1)
bigru_kernel(const float* __restrict__ xproj, const float* __restrict__ ug,
             const float* __restrict__ uc, const int32_t* __restrict__ lengths, int N, int T,
             float* __restrict__ out, int64_t out_bs) {
  const int tid = threadIdx.x;
  const int dir = blockIdx.y;
  const int n0 = blockIdx.x * NS;

  __shared__ __align__(16) float hs[NS][H];
  __shared__ __align__(16) float rh[NS][H];
  __shared__ float us[NS][H];
  __shared__ float pg[2][NS][2 * H];
  __shared__ float pc[4][NS][H];
  __shared__ __align__(16) float ring[RING][NS][3 * H];
  __shared__ int s_len[NS];

  // ---- recurrent weights -> registers (resident for the whole sequence) ----
  const int gcol = tid & 255, gk0 = (tid >> 8) * 64;
  const int ccol = tid & 127, ck0 = (tid >> 7) * 32;
  float wg[64], wc[32];
  {
    const float* Ug = ug + (size_t)dir * H * 2 * H;
    const float* Uc = uc + (size_t)dir * H * H;
#pragma unroll
    for (int i = 0; i < 64; ++i) wg[i] = __ldg(Ug + (gk0 + i) * (2 * H) + gcol);
#pragma unroll
    for (int i = 0; i < 32; ++i) wc[i] = __ldg(Uc + (ck0 + i) * H + ccol);
  }

  if (tid < NS) {
    int L = 0;
    if (n0 + tid < N) {
      L = lengths ? lengths[n0 + tid] : T;
      L = max(0, min(L, T));
    }
    s_len[tid] = L;
  }
  for (int i = tid; i < NS * H; i += 512) (&hs[0][0])[i] = 0.f;
  __syncthreads();
  int len[NS], maxlen = 0;
#pragma unroll
  for (int s = 0; s < NS; ++s) { len[s] = s_len[s]; maxlen = max(maxlen, len[s]); }

  // outputs are zero for t >= len (dynamic_rnn zero_output); this CTA's 128 columns.
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    if (n0 + s >= N) continue;
    float* o = out + (int64_t)(n0 + s) * out_bs + dir * H;
    for (int i = len[s] * H + tid; i < T * H; i += 512) o[(int64_t)(i >> 7) * (2 * H) + (i & 127)] = 0.f;
  }

  // cp.async producer: thread (s, q) copies float4 q of sample s' 384-float row.
  const int ps = tid / 96, pq = tid - ps * 96;
  auto issue = [&](int step) {
    const int L = ps < NS ? s_len[ps] : 0;
    if (step < L) {
      const int pos = dir == 0 ? step : L - 1 - step;
      const float* src = xproj + ((int64_t)(n0 + ps) * T + pos) * XW + dir * (3 * H) + pq * 4;
      cp_async16(&ring[step % RING][ps][pq * 4], src);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int p = 0; p < RING - 1; ++p) issue(p);

  for (int step = 0; step < maxlen; ++step) {
    issue(step + RING - 1);
    cp_async_wait<RING - 1>();   // the group of `step` has landed (this thread's part)

    // ---- gate partial sums: h[gk0..gk0+64) . U_g[:, gcol] ----
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        const float4 hv = *reinterpret_cast<const float4*>(&hs[s][gk0 + i]);
        a0 = fmaf(hv.x, wg[i], a0);
        a1 = fmaf(hv.y, wg[i + 1], a1);
        a2 = fmaf(hv.z, wg[i + 2], a2);
        a3 = fmaf(hv.w, wg[i + 3], a3);
      }
      pg[tid >> 8][s][gcol] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    const float* xr = &ring[step % RING][0][0];
    if (tid < 2 * H) {
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const float g = sigmoid_f(xr[s * 3 * H + gcol] + pg[0][s][gcol] + pg[1][s][gcol]);
        if (gcol < H) rh[s][gcol] = g * hs[s][gcol];
        else us[s][gcol - H] = g;
      }
    }
    __syncthreads();
    // ---- candidate partial sums: (r*h)[ck0..ck0+32) . U_c[:, ccol] ----
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 hv = *reinterpret_cast<const float4*>(&rh[s][ck0 + i]);
        a0 = fmaf(hv.x, wc[i], a0);
        a1 = fmaf(hv.y, wc[i + 1], a1);
        a2 = fmaf(hv.z, wc[i + 2], a2);
        a3 = fmaf(hv.w, wc[i + 3], a3);
      }
      pc[tid >> 7][s][ccol] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    if (tid < H) {
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (step < len[s]) {
          const float c = tanh_f(xr[s * 3 * H + 2 * H + ccol] +
                                 ((pc[0][s][ccol] + pc[1][s][ccol]) + (pc[2][s][ccol] + pc[3][s][ccol])));
          const float u = us[s][ccol], h = hs[s][ccol];
          const float hn = u * h + (1.0f - u) * c;
          hs[s][ccol] = hn;
          const int pos = dir == 0 ? step : len[s] - 1 - step;
          out[(int64_t)(n0 + s) * out_bs + (int64_t)pos * (2 * H) + dir * H + ccol] = hn;
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
}

}  // namespace

void launch_bigru(const float* xproj, const float* ug, const float* uc, const int32_t* lengths,
                  int N, int T, float* out, int64_t out_bs, cudaStream_t st) {
  if (N <= 0 || T <= 0) return;
  // Samples per CTA: keep 2*ceil(N/NS) CTAs within one wave of 148 SMs when possible.
  int NS = 1;
  while (NS < 4 && 2 * ((N + NS - 1) / NS) > 148) NS *= 2;
  dim3 grid((N + NS - 1) / NS, 2);
  if (NS == 1) bigru_kernel<1><<<grid, 512, 0, st>>>(xproj, ug, uc, lengths, N, T, out, out_bs);
  else if (NS == 2) bigru_kernel<2><<<grid, 512, 0, st>>>(xproj, ug, uc, lengths, N, T, out, out_bs);
  else bigru_kernel<4><<<grid, 512, 0, st>>>(xproj, ug, uc, lengths, N, T, out, out_bs);
}

}  // namespace taco
