// conv1d('same') / dense as an implicit GEMM, fp32 FFMA path.
//
// Replaces tf.layers.conv1d + activation + (inference) batch_normalization at
// reference models/modules.py:93-101, tf.layers.dense at modules.py:10,59-60,
// 79-89 and models/tacotron.py:101, and the BahdanauAttention memory layer
// (tacotron.py:68).  Rows of the GEMM are the flattened (n,t) positions; the
// K axis is (tap j, input channel c); the im2col matrix is never materialised:
// each A-tile load shifts the time index by (j - pad_left) and zero-fills rows
// that fall outside [0,T)  (pad_left=(k-1)/2 -- TF 'same' for stride 1).
//
// 128x128x16 tiles, 256 threads, 8x8 register micro-tile, double-buffered
// shared memory with register prefetch.  Epilogue: +bias -> activation ->
// per-channel affine (folded BN) -> +residual, or the highway gate
// H*T + x*(1-T) over interleaved (H,T) column pairs (modules.py:77-90).
#include "common.cuh"
#include "kernels.cuh"

namespace taco {

namespace {
constexpr int BK = 16, NT = 256, APAD = 4;

// BM = 64*HM rows x BN = 64*HN columns per CTA; a thread owns (4*HM) x (4*HN) outputs.
template <int HM, int HN>
__global__ void __launch_bounds__(NT, 2) conv_gemm_kernel(const ConvGemm p) {
  constexpr int BM = 64 * HM, BN = 64 * HN, RM = 4 * HM, RN = 4 * HN;
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int M = p.N * p.T;
  const int Ktot = p.k * p.Cin;
  const int pl = (p.k - 1) >> 1;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // A-load role: two (row, k-quad) slots per thread.
  int a_row[HM], a_kq[HM], a_t[HM];
  const float* a_base[HM];
  bool a_valid[HM];
#pragma unroll
  for (int l = 0; l < HM; ++l) {
    const int idx = tid + l * NT;
    a_row[l] = idx >> 2;
    a_kq[l] = idx & 3;
    const int m = m0 + a_row[l];
    a_valid[l] = m < M;
    const int n = a_valid[l] ? m / p.T : 0;
    a_t[l] = a_valid[l] ? m - n * p.T : 0;
    a_base[l] = p.x + (int64_t)n * p.x_bs;
  }
  // B-load role.
  int b_kk[HN], b_col[HN];
#pragma unroll
  for (int l = 0; l < HN; ++l) {
    const int idx = tid + l * NT;
    b_kk[l] = idx / (BN / 4);
    b_col[l] = (idx % (BN / 4)) << 2;
  }

  float4 ra[HM], rb[HN];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int l = 0; l < HM; ++l) {
      const int kk = k0 + a_kq[l] * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_valid[l] && kk < Ktot) {
        const int j = kk / p.Cin;
        const int c = kk - j * p.Cin;
        const int tt = a_t[l] + j - pl;
        if (tt >= 0 && tt < p.T) v = ldg_f4(a_base[l] + (int64_t)tt * p.ldx + c);
      }
      ra[l] = v;
    }
#pragma unroll
    for (int l = 0; l < HN; ++l) {
      const int kb = k0 + b_kk[l];
      const int col = n0 + b_col[l];
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kb < Ktot && col < p.ldw) w = ldg_f4(p.w + (int64_t)kb * p.ldw + col);
      rb[l] = w;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int l = 0; l < HM; ++l) {
      const int kq = a_kq[l] * 4, r = a_row[l];
      As[buf][kq + 0][r] = ra[l].x;
      As[buf][kq + 1][r] = ra[l].y;
      As[buf][kq + 2][r] = ra[l].z;
      As[buf][kq + 3][r] = ra[l].w;
    }
#pragma unroll
    for (int l = 0; l < HN; ++l) *reinterpret_cast<float4*>(&Bs[buf][b_kk[l]][b_col[l]]) = rb[l];
  };

  float acc[RM][RN];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;

  const int nk = (Ktot + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int it = 0; it < nk; ++it) {
    const int buf = it & 1;
    if (it + 1 < nk) load_tiles((it + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[RM], b[RN];
#pragma unroll
      for (int hh = 0; hh < HM; ++hh) {
        const float4 t4 = *reinterpret_cast<const float4*>(&As[buf][kk][hh * 64 + ty * 4]);
        a[hh * 4 + 0] = t4.x; a[hh * 4 + 1] = t4.y; a[hh * 4 + 2] = t4.z; a[hh * 4 + 3] = t4.w;
      }
#pragma unroll
      for (int hh = 0; hh < HN; ++hh) {
        const float4 t4 = *reinterpret_cast<const float4*>(&Bs[buf][kk][hh * 64 + tx * 4]);
        b[hh * 4 + 0] = t4.x; b[hh * 4 + 1] = t4.y; b[hh * 4 + 2] = t4.z; b[hh * 4 + 3] = t4.w;
      }
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (it + 1 < nk) {
      store_tiles(buf ^ 1);   // other buffer: last read two iterations ago, fenced by the sync below
      __syncthreads();
    }
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int m = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
    if (m >= M) continue;
    const int n = m / p.T, t = m - n * p.T;
    float* orow = p.out + (int64_t)n * p.out_bs + (int64_t)t * p.ldo + p.col_off;
    const float* rrow = p.res ? p.res + (int64_t)n * p.res_bs + (int64_t)t * p.ldres : nullptr;
#pragma unroll
    for (int jh = 0; jh < HN; ++jh) {
      const int cbase = n0 + jh * 64 + tx * 4;
      if (p.epi == EPI_PLAIN) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = cbase + j;
          float x = acc[i][jh * 4 + j];
          if (c < p.Cout) {
            if (p.bias) x += __ldg(p.bias + c);
            x = apply_act(x, p.act);
            if (p.scale) x = fmaf(x, __ldg(p.scale + c), __ldg(p.shift + c));
            if (rrow) x += __ldg(rrow + c);
          }
          v[j] = x;
        }
        if (cbase + 3 < p.Cout && ((reinterpret_cast<uintptr_t>(orow + cbase) & 15) == 0)) {
          *reinterpret_cast<float4*>(orow + cbase) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (cbase + j < p.Cout) orow[cbase + j] = v[j];
        }
      } else {  // EPI_HIGHWAY: columns (2c, 2c+1) = (H_c, T_c); carry input in `res`
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          const int c = cbase + j;
          if (c + 1 < p.Cout) {
            const int ch = c >> 1;
            const float H = fmaxf(acc[i][jh * 4 + j] + __ldg(p.bias + c), 0.f);
            const float Tg = sigmoid_f(acc[i][jh * 4 + j + 1] + __ldg(p.bias + c + 1));
            const float xin = __ldg(rrow + ch);
            orow[ch] = H * Tg + xin * (1.0f - Tg);
          }
        }
      }
    }
  }
}
}  // namespace

void launch_conv_gemm(const ConvGemm& p, cudaStream_t st) {
  const int M = p.N * p.T;
  if (M <= 0 || p.Cout <= 0) return;
  // Big tiles when they already fill the 148 SMs, else 64x64 tiles for more CTAs in flight.
  const int big = ((M + 127) / 128) * ((p.Cout + 127) / 128);
  if (big >= 148) {
    dim3 grid((M + 127) / 128, (p.Cout + 127) / 128);
    conv_gemm_kernel<2, 2><<<grid, NT, 0, st>>>(p);
  } else {
    dim3 grid((M + 63) / 64, (p.Cout + 63) / 64);
    conv_gemm_kernel<1, 1><<<grid, NT, 0, st>>>(p);
  }
}

}  // namespace taco
