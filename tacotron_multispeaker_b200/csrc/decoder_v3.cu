// K7 (v3): attention decoder loop, warp-owned hidden units, cluster of 16 CTAs.
//
// Same operator and same cluster/DSMEM exchange as decoder.cu (reference
// models/tacotron.py:66-94, models/rnn_wrappers.py, models/helpers.py) with
// the per-CTA work re-cut so that NOTHING inside a step needs a block barrier:
//
//  * a warp owns one hidden unit c = 16*q + w of every 256-wide layer (q = CTA
//    rank in the cluster, w = warp): it computes the r, u, candidate-x and
//    candidate-h dot products of that unit itself, so all GRU gating happens
//    in registers of one warp;
//  * the 32 lanes split K (lane l owns k = l, l+32, ...), S samples share every
//    weight load; the per-(column,sample) partial sums are combined by a
//    recursive-halving shuffle reduction inside the warp;
//  * results go to all 16 CTAs with st.async (+ mbarrier byte credit) straight
//    from the producing warp; consumers wait on the per-phase mbarrier;
//  * the recurrent halves of the three GRU gate matrices (they multiply the OLD
//    state) are resident in shared memory and evaluated while the previous
//    phase's results are still in flight; the streamed halves are double
//    buffered in registers and fetched from L2 one full phase ahead;
//  * softmax: producers push exp(score - B) with B = ||v||_1 >= |score| (tanh
//    is bounded), so no max pass is needed; the normaliser is summed together
//    with the context (exactly softmax after the division).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"

namespace taco {

namespace {
constexpr int CS3 = 16;        // CTAs per cluster
constexpr int NT3 = 512;       // threads per CTA
constexpr int NW3 = 16;        // warps per CTA = hidden units per CTA
constexpr int DH = 256, DP = 128;
constexpr int XIN_ROWS = 384;  // [frame(M<=128) | ctx(256)] padded to a multiple of 128 rows

// float4 per lane of the streamed weight blocks, in consumption order
//            P1 P2 P3 P4 P5 P8 P9 P10 P11 P12 P13
constexpr int F4_P1 = 3, F4_P2 = 2, F4_P3 = 3, F4_P4 = 2, F4_P5 = 4, F4_P8 = 2, F4_P9 = 6, F4_P10 = 2, F4_P11 = 6,
              F4_P12 = 2, F4_P13 = 4;
constexpr int OFF_P1 = 0, OFF_P2 = OFF_P1 + F4_P1, OFF_P3 = OFF_P2 + F4_P2, OFF_P4 = OFF_P3 + F4_P3,
              OFF_P5 = OFF_P4 + F4_P4, OFF_P8 = OFF_P5 + F4_P5, OFF_P9 = OFF_P8 + F4_P8, OFF_P10 = OFF_P9 + F4_P9,
              OFF_P11 = OFF_P10 + F4_P10, OFF_P12 = OFF_P11 + F4_P11, OFF_P13 = OFF_P12 + F4_P12,
              F4_STEP = OFF_P13 + F4_P13;                       // 36 float4 per lane per step
constexpr int F4_E = 4;                                         // resident block per GRU: 2 columns x K=256

__device__ __forceinline__ void mbar_init3(uint64_t* mb, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mb)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect3(uint64_t* mb, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mb)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait3(uint64_t* mb, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(mb)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void st_async4(uint32_t raddr, float a, float b, float c, float d, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%2,%3,%4}, [%5];"
               ::"r"(raddr), "f"(a), "f"(b), "f"(c), "f"(d), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void st_async2(uint32_t raddr, float a, float b, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1,%2}, [%3];"
               ::"r"(raddr), "f"(a), "f"(b), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void st_async1(uint32_t raddr, float a, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];"
               ::"r"(raddr), "f"(a), "r"(rmbar) : "memory");
}

template <int S> struct Log2;
template <> struct Log2<1> { static constexpr int v = 0; };
template <> struct Log2<2> { static constexpr int v = 1; };
template <> struct Log2<4> { static constexpr int v = 2; };
template <> struct Log2<8> { static constexpr int v = 3; };

// Streamed weight block of one phase: NF4 float4 per lane, consecutive lanes consecutive float4.
template <int NF4>
__device__ __forceinline__ void load_w(const float4* __restrict__ g, float4 (&w)[NF4]) {
#pragma unroll
  for (int i = 0; i < NF4; ++i) w[i] = __ldg(g + i * 32);
}

// acc[s*NCP + col] += sum over this lane's rows k = lane + 32*(4g+j) of xs[k][s] * W[col][k].
// w[g*NCOL + col] holds the 4 weights j = 0..3 of column `col`.  xs: shared [rows][S].
template <int S, int NCP, int NCOL, int KG, class WT>
__device__ __forceinline__ void dot_acc(const WT& w, const float* __restrict__ xs, int lane, float (&acc)[S * NCP]) {
#pragma unroll
  for (int g = 0; g < KG; ++g) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = lane + 32 * (4 * g + j);
      float x[S];
      if constexpr (S % 4 == 0) {
#pragma unroll
        for (int s4 = 0; s4 < S / 4; ++s4) {
          const float4 t4 = *reinterpret_cast<const float4*>(xs + (size_t)k * S + s4 * 4);
          x[s4 * 4] = t4.x; x[s4 * 4 + 1] = t4.y; x[s4 * 4 + 2] = t4.z; x[s4 * 4 + 3] = t4.w;
        }
      } else if constexpr (S == 2) {
        const float2 t2 = *reinterpret_cast<const float2*>(xs + (size_t)k * 2);
        x[0] = t2.x; x[1] = t2.y;
      } else {
        x[0] = xs[k];
      }
#pragma unroll
      for (int c = 0; c < NCOL; ++c) {
        const float4 w4 = w[g * NCOL + c];
        const float wj = j == 0 ? w4.x : (j == 1 ? w4.y : (j == 2 ? w4.z : w4.w));
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s * NCP + c] = fmaf(x[s], wj, acc[s * NCP + c]);
      }
    }
  }
}

// Sum acc[] over the 32 lanes.  Values are indexed s*NCP + col.  Recursive halving over the
// sample index (each round a lane keeps the half selected by one of its lane bits and adds
// the partner's partial sums for that half), then butterflies for the remaining lane bits.
// On return every lane holds in acc[0..NCP) the totals of sample s(lane) = lane >> (5 - log2 S).
template <int N, int OFF, int NCP>
struct WarpHalve {
  template <int V>
  static __device__ __forceinline__ void run(float (&v)[V], int lane) {
    if constexpr (OFF >= 1) {
      if constexpr (N > NCP) {
        constexpr int h = N / 2;
        const bool up = (lane & OFF) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
          const float send = up ? v[i] : v[i + h];
          const float keep = up ? v[i + h] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        WarpHalve<h, OFF / 2, NCP>::run(v, lane);
      } else {
#pragma unroll
        for (int i = 0; i < NCP; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], OFF);
        WarpHalve<N, OFF / 2, NCP>::run(v, lane);
      }
    }
  }
};
template <int S, int NCP>
__device__ __forceinline__ void warp_reduce(float (&acc)[S * NCP], int lane) {
  WarpHalve<S * NCP, 16, NCP>::run(acc, lane);
}

// Every lane holds the value of sample s(lane); return all S values in every lane.
template <int S>
__device__ __forceinline__ void gather_all(float v, float (&out)[S]) {
#pragma unroll
  for (int s = 0; s < S; ++s) out[s] = __shfl_sync(0xffffffffu, v, s << (5 - Log2<S>::v));
}

// Write the S values of one activation row ([row][S] layout, local address `dst`) into all 16 CTAs.
template <int S>
__device__ __forceinline__ void push_row(float* dst, const float (&vals)[S], uint64_t* mb, int lane) {
  const uint32_t d0 = smem_u32(dst), m0 = smem_u32(mb);
  const int peer = lane & 15;
  const uint32_t ra = mapa_u32(d0, (uint32_t)peer), rm = mapa_u32(m0, (uint32_t)peer);
  if constexpr (S == 8) {
    if (lane < 16) st_async4(ra, vals[0], vals[1], vals[2], vals[3], rm);
    else st_async4(ra + 16, vals[4], vals[5], vals[6], vals[7], rm);
  } else if constexpr (S == 4) {
    if (lane < 16) st_async4(ra, vals[0], vals[1], vals[2], vals[3], rm);
  } else if constexpr (S == 2) {
    if (lane < 16) st_async2(ra, vals[0], vals[1], rm);
  } else {
    if (lane < 16) st_async1(ra, vals[0], rm);
  }
}
// One float (local address dst) into all 16 CTAs; executed by lanes 0..15 of a warp.
__device__ __forceinline__ void push_scalar(float* dst, float v, uint64_t* mb, int lane) {
  if (lane < 16) {
    const uint32_t ra = mapa_u32(smem_u32(dst), (uint32_t)lane), rm = mapa_u32(smem_u32(mb), (uint32_t)lane);
    st_async1(ra, v, rm);
  }
}

enum { NB3 = 16 };
enum { B3_P1 = 0, B3_P2, B3_P3, B3_P4, B3_P5, B3_P6, B3_P7, B3_P8, B3_P9, B3_P10, B3_P11, B3_P12, B3_P13 };

struct Layout3 {
  int xin, p1, in3, rhA, pq, sc, in9, rh1, in11, rh2, y2, ew, ksl, msl, total;
};
__host__ __device__ inline Layout3 make_layout3(int S, int T_in, bool att_res) {
  Layout3 L;
  int o = 2 * NB3;
  auto take = [&](int nfloats) { int r = o; o += (nfloats + 3) & ~3; return r; };
  L.xin = take(XIN_ROWS * S);
  L.p1 = take(DH * S);
  L.in3 = take((DP + DH) * S);
  L.rhA = take(DH * S);
  L.pq = take(DH * S);                 // [S][256]
  L.sc = take(T_in * S);               // exp(score - B), [T_in][S]
  L.in9 = take(2 * DH * S);
  L.rh1 = take(DH * S);
  L.in11 = take(2 * DH * S);
  L.rh2 = take(DH * S);
  L.y2 = take(DH * S);
  L.ew = take(NW3 * 3 * F4_E * 32 * 4);              // resident recurrent gate halves: [warp][gru][F4_E][lane][4]
  const int Tj = (T_in + CS3 - 1) / CS3;
  L.ksl = take(att_res ? S * Tj * DH : 0);
  L.msl = take(att_res ? S * T_in * NW3 : 0);
  L.total = o;
  return L;
}

template <int S>
__global__ void __launch_bounds__(NT3, 1)
decoder_v3_kernel(const DecoderWeightsV3 w, const DecoderArgs a) {
  constexpr int LS = Log2<S>::v;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  const int q = (int)cluster_ctarank();
  const int n0 = (int)cluster_id_x() * S;
  const int M = w.M, Dout = w.Dout, T_in = a.T_in;
  const bool att_res = a.att_res != 0;
  const Layout3 L = make_layout3(S, T_in, att_res);
  uint64_t* mb = reinterpret_cast<uint64_t*>(smem);
  float* xin = smem + L.xin;   float* p1 = smem + L.p1;     float* in3 = smem + L.in3;
  float* rhA = smem + L.rhA;   float* pqT = smem + L.pq;    float* sc = smem + L.sc;
  float* in9 = smem + L.in9;   float* rh1 = smem + L.rh1;   float* in11 = smem + L.in11;
  float* rh2 = smem + L.rh2;   float* y2 = smem + L.y2;
  float* ksl = smem + L.ksl;   float* msl = smem + L.msl;
  const int c = q * NW3 + wp;                      // hidden unit owned by this warp
  const int sL = lane >> (5 - LS);                 // sample whose totals this lane holds after warp_reduce
  const int nL = n0 + sL;

  // ---- prologue (block-wide): zero state, barriers, resident weights ----
  for (int i = tid + 2 * NB3; i < L.total; i += NT3) smem[i] = 0.f;
  if (tid < NB3) mbar_init3(mb + tid, 1);
  __syncthreads();
  {
    const float4* src = reinterpret_cast<const float4*>(w.ew) + (size_t)q * NW3 * 3 * F4_E * 32;
    float4* dst = reinterpret_cast<float4*>(smem + L.ew);
    for (int i = tid; i < NW3 * 3 * F4_E * 32; i += NT3) dst[i] = __ldg(src + i);
  }
  const int Tj = (T_in + CS3 - 1) / CS3;
  const int j0 = min(q * Tj, T_in), j1 = min(j0 + Tj, T_in);
  if (att_res) {
    for (int i = tid; i < S * (j1 - j0) * (DH / 4); i += NT3) {
      const int c4 = i % (DH / 4), r = i / (DH / 4), s = r / (j1 - j0), jj = r - s * (j1 - j0), n = n0 + s;
      const float4 v4 = (n < a.N) ? ldg_f4(a.keys + ((size_t)n * T_in + j0 + jj) * DH + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(ksl + ((size_t)s * Tj + jj) * DH + c4 * 4) = v4;
    }
    for (int i = tid; i < S * T_in * (NW3 / 4); i += NT3) {
      const int c4 = i % (NW3 / 4), r = i / (NW3 / 4), s = r / T_in, j = r - s * T_in, n = n0 + s;
      const float4 v4 = (n < a.N) ? ldg_f4(a.memory + ((size_t)n * T_in + j) * DH + q * NW3 + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(msl + ((size_t)s * T_in + j) * NW3 + c4 * 4) = v4;
    }
  }
  // per-warp constants
  const float b_p1 = __ldg(w.p1_b + c), b_p2 = wp < 8 ? __ldg(w.p2_b + q * 8 + wp) : 0.f;
  const float b_ra = __ldg(w.ga_b + c), b_ua = __ldg(w.ga_b + DH + c), b_ca = __ldg(w.ca_b + c);
  const float b_pc = __ldg(w.pc_b + c);
  const float b_r1 = __ldg(w.g1_b + c), b_u1 = __ldg(w.g1_b + DH + c), b_c1 = __ldg(w.c1_b + c);
  const float b_r2 = __ldg(w.g2_b + c), b_u2 = __ldg(w.g2_b + DH + c), b_c2 = __ldg(w.c2_b + c);
  const int ocol0 = q * 32 + 2 * wp;                // output-projection columns of this warp
  const float b_o0 = ocol0 < Dout ? __ldg(w.o_b + ocol0) : 0.f, b_o1 = ocol0 + 1 < Dout ? __ldg(w.o_b + ocol0 + 1) : 0.f;
  float vreg[8];
  float vbound = 0.f;                               // B = ||v||_1 >= |score|
#pragma unroll
  for (int i = 0; i < 8; ++i) { vreg[i] = __ldg(w.att_v + lane + 32 * i); vbound += fabsf(vreg[i]); }
  vbound = fminf(warp_sum(vbound), 40.0f);         // keeps exp(score - B) inside the fp32 range
  const float4* ws = reinterpret_cast<const float4*>(w.stream) + ((size_t)(q * NW3 + wp) * F4_STEP) * 32 + lane;
  const float4* ewp = reinterpret_cast<const float4*>(smem + L.ew) + (size_t)wp * 3 * F4_E * 32 + lane;
  auto ew_block = [&](int gru, float4 (&e)[F4_E]) {
#pragma unroll
    for (int i = 0; i < F4_E; ++i) e[i] = ewp[(gru * F4_E + i) * 32];
  };

  float4 wA[6], wB[6];                              // streamed weights: current / next phase
  load_w<F4_P1>(ws + OFF_P1 * 32, reinterpret_cast<float4(&)[F4_P1]>(wA));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  cluster_sync_all();

  const bool free_run = a.targets == nullptr;
  const int fb0 = Dout - M;
  const uint32_t ROWB = 256u * S * 4u;              // bytes a CTA receives for one 256-unit phase

  for (int t = 0; t < a.steps; ++t) {
    const uint32_t par = (uint32_t)t & 1u;
    if (t > 0) mbar_wait3(mb + B3_P13, par ^ 1u);   // frame fed back by step t-1 (all warps: also orders the expects below)
    if (tid == 0) {
      mbar_expect3(mb + B3_P1, ROWB);
      mbar_expect3(mb + B3_P2, ROWB / 2);
      mbar_expect3(mb + B3_P3, ROWB);
      mbar_expect3(mb + B3_P4, ROWB);
      mbar_expect3(mb + B3_P5, ROWB);
      mbar_expect3(mb + B3_P6, (uint32_t)T_in * S * 4u);
      mbar_expect3(mb + B3_P7, ROWB);
      mbar_expect3(mb + B3_P8, ROWB);
      mbar_expect3(mb + B3_P9, ROWB);
      mbar_expect3(mb + B3_P10, 2 * ROWB);
      mbar_expect3(mb + B3_P11, ROWB);
      mbar_expect3(mb + B3_P12, 2 * ROWB);
      mbar_expect3(mb + B3_P13, (uint32_t)M * S * 4u);
    }
    float vals[S];
    // ===== P1: prenet dense_1 + ReLU on [frame | ctx] =====
    load_w<F4_P2>(ws + OFF_P2 * 32, reinterpret_cast<float4(&)[F4_P2]>(wB));
    {
      float acc[S];
#pragma unroll
      for (int i = 0; i < S; ++i) acc[i] = 0.f;
      dot_acc<S, 1, 1, 3>(wA, xin, lane, acc);
      warp_reduce<S, 1>(acc, lane);
      gather_all<S>(fmaxf(acc[0] + b_p1, 0.f), vals);
      push_row<S>(p1 + c * S, vals, mb + B3_P1, lane);
    }
    // ===== P2: prenet dense_2 + ReLU (8 units per CTA: warps 0..7) =====
    load_w<F4_P3>(ws + OFF_P3 * 32, reinterpret_cast<float4(&)[F4_P3]>(wA));
    float eacc[S * 4];                               // gate pre-activations: [s][r,u,cx,-]
    {
      // recurrent half of the attention-GRU gates on the OLD h_att, while P1's results are in flight
      float4 e[F4_E];
      ew_block(0, e);
#pragma unroll
      for (int i = 0; i < S * 4; ++i) eacc[i] = 0.f;
      dot_acc<S, 4, 2, 2>(e, in3 + DP * S, lane, eacc);
    }
    mbar_wait3(mb + B3_P1, par);
    if (wp < 8) {
      float acc[S];
#pragma unroll
      for (int i = 0; i < S; ++i) acc[i] = 0.f;
      dot_acc<S, 1, 1, 2>(wB, p1, lane, acc);
      warp_reduce<S, 1>(acc, lane);
      gather_all<S>(fmaxf(acc[0] + b_p2, 0.f), vals);
      push_row<S>(in3 + (q * 8 + wp) * S, vals, mb + B3_P2, lane);
    }
    // ===== P3: attention GRU gates (x half) + candidate x-part; r*h pushed =====
    load_w<F4_P4>(ws + OFF_P4 * 32, reinterpret_cast<float4(&)[F4_P4]>(wB));
    mbar_wait3(mb + B3_P2, par);
    float u_keep, cx_keep, hold_keep;
    {
      dot_acc<S, 4, 3, 1>(wA, in3, lane, eacc);
      warp_reduce<S, 4>(eacc, lane);
      const float r = sigmoid_f(eacc[0] + b_ra);
      u_keep = sigmoid_f(eacc[1] + b_ua);
      cx_keep = eacc[2];
      hold_keep = in3[(DP + c) * S + sL];
      gather_all<S>(r * hold_keep, vals);
      push_row<S>(rhA + c * S, vals, mb + B3_P3, lane);
    }
    // ===== P4: candidate h-part -> h_att' =====
    load_w<F4_P5>(ws + OFF_P5 * 32, reinterpret_cast<float4(&)[F4_P5]>(wA));
    mbar_wait3(mb + B3_P3, par);
    {
      float acc[S];
#pragma unroll
      for (int i = 0; i < S; ++i) acc[i] = 0.f;
      dot_acc<S, 1, 1, 2>(wB, rhA, lane, acc);
      warp_reduce<S, 1>(acc, lane);
      const float cnd = tanh_f(acc[0] + cx_keep + b_ca);
      gather_all<S>(u_keep * hold_keep + (1.0f - u_keep) * cnd, vals);
      push_row<S>(in3 + (DP + c) * S, vals, mb + B3_P4, lane);
    }
    // ===== P5: query layer + h_att' half of the 512->256 projection =====
    load_w<F4_P8>(ws + OFF_P8 * 32, reinterpret_cast<float4(&)[F4_P8]>(wB));
    mbar_wait3(mb + B3_P4, par);
    float y0h_keep;
    {
      float acc[S * 2];
#pragma unroll
      for (int i = 0; i < S * 2; ++i) acc[i] = 0.f;
      dot_acc<S, 2, 2, 2>(wA, in3 + DP * S, lane, acc);
      warp_reduce<S, 2>(acc, lane);
      y0h_keep = acc[1];
      gather_all<S>(acc[0], vals);
#pragma unroll
      for (int s = 0; s < S; ++s) push_scalar(pqT + s * DH + c, vals[s], mb + B3_P5, lane);   // [S][256] layout
    }
    // ===== P6: Bahdanau scores of positions [j0,j1) -> exp(score - B) =====
    load_w<F4_P9>(ws + OFF_P9 * 32, reinterpret_cast<float4(&)[F4_P9]>(wA));
    mbar_wait3(mb + B3_P5, par);
    {
      const int npairs = S * (j1 - j0);
      for (int pi = wp; pi < npairs; pi += NW3) {
        const int jj = pi / S, s = pi - jj * S, n = n0 + s;
        const float* prow = pqT + s * DH;
        float e = 0.f;
        if (att_res) {
          const float* krow = ksl + ((size_t)s * Tj + jj) * DH;
#pragma unroll
          for (int i = 0; i < 8; ++i) e = fmaf(vreg[i], tanh_f(krow[lane + 32 * i] + prow[lane + 32 * i]), e);
        } else if (n < a.N) {
          const float* krow = a.keys + ((size_t)n * T_in + (j0 + jj)) * DH;
          float kv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) kv[i] = __ldg(krow + lane + 32 * i);
#pragma unroll
          for (int i = 0; i < 8; ++i) e = fmaf(vreg[i], tanh_f(kv[i] + prow[lane + 32 * i]), e);
        }
        e = warp_sum(e);
        push_scalar(sc + (j0 + jj) * S + s, __expf(e - vbound), mb + B3_P6, lane);
      }
    }
    // recurrent half of GRU-1 gates on the OLD h1 while the scores travel
    {
      float4 e[F4_E];
      ew_block(1, e);
#pragma unroll
      for (int i = 0; i < S * 4; ++i) eacc[i] = 0.f;
      dot_acc<S, 4, 2, 2>(e, in9 + DH * S, lane, eacc);
    }
    // ===== P7: context of unit c: sum_j p_j * memory[j][c] / sum_j p_j  (+ alignments) =====
    mbar_wait3(mb + B3_P6, par);
    {
      float acc[S * 2];                              // [s][ctx, normaliser]
#pragma unroll
      for (int i = 0; i < S * 2; ++i) acc[i] = 0.f;
      for (int j = lane; j < T_in; j += 32) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const float p = sc[j * S + s];
          float m = 0.f;
          if (att_res) m = msl[((size_t)s * T_in + j) * NW3 + wp];
          else if (n0 + s < a.N) m = __ldg(a.memory + ((size_t)(n0 + s) * T_in + j) * DH + c);
          acc[s * 2] = fmaf(p, m, acc[s * 2]);
          acc[s * 2 + 1] += p;
        }
      }
      warp_reduce<S, 2>(acc, lane);
      const float inv = 1.0f / acc[1];
      gather_all<S>(acc[0] * inv, vals);
      push_row<S>(xin + (M + c) * S, vals, mb + B3_P7, lane);
      if (a.align_out != nullptr) {                  // alignments of this CTA's positions: warp w writes j0 + w, j0 + w + 16, ...
        float invs[S];
        gather_all<S>(inv, invs);
        for (int jj = wp; jj < j1 - j0; jj += NW3) {
          if (lane < S && n0 + lane < a.N) {
            float iv = invs[0];
#pragma unroll
            for (int s = 1; s < S; ++s) iv = lane == s ? invs[s] : iv;
            a.align_out[((size_t)(n0 + lane) * T_in + (j0 + jj)) * a.max_steps + t] = sc[(j0 + jj) * S + lane] * iv;
          }
        }
      }
    }
    // ===== P8: y0 = [h_att' | ctx] W_p + b (ctx half here, h half from P5) =====
    mbar_wait3(mb + B3_P7, par);
    float y0_keep;
    {
      float acc[S];
#pragma unroll
      for (int i = 0; i < S; ++i) acc[i] = 0.f;
      dot_acc<S, 1, 1, 2>(wB, xin + M * S, lane, acc);
      warp_reduce<S, 1>(acc, lane);
      y0_keep = acc[0] + y0h_keep + b_pc;
      gather_all<S>(y0_keep, vals);
      push_row<S>(in9 + c * S, vals, mb + B3_P8, lane);
    }
    // ===== P9: GRU-1 gates (x half) + candidate x-part =====
    load_w<F4_P10>(ws + OFF_P10 * 32, reinterpret_cast<float4(&)[F4_P10]>(wB));
    mbar_wait3(mb + B3_P8, par);
    {
      dot_acc<S, 4, 3, 2>(wA, in9, lane, eacc);
      warp_reduce<S, 4>(eacc, lane);
      const float r = sigmoid_f(eacc[0] + b_r1);
      u_keep = sigmoid_f(eacc[1] + b_u1);
      cx_keep = eacc[2];
      hold_keep = in9[(DH + c) * S + sL];
      gather_all<S>(r * hold_keep, vals);
      push_row<S>(rh1 + c * S, vals, mb + B3_P9, lane);
    }
    // recurrent half of GRU-2 gates on the OLD h2 while r*h1 travels
    load_w<F4_P11>(ws + OFF_P11 * 32, reinterpret_cast<float4(&)[F4_P11]>(wA));
    {
      float4 e[F4_E];
      ew_block(2, e);
#pragma unroll
      for (int i = 0; i < S * 4; ++i) eacc[i] = 0.f;
      dot_acc<S, 4, 2, 2>(e, in11 + DH * S, lane, eacc);
    }
    // ===== P10: GRU-1 candidate h-part -> h1', y1 = y0 + h1' =====
    mbar_wait3(mb + B3_P9, par);
    float y1_keep;
    {
      float acc[S];
#pragma unroll
      for (int i = 0; i < S; ++i) acc[i] = 0.f;
      dot_acc<S, 1, 1, 2>(wB, rh1, lane, acc);
      warp_reduce<S, 1>(acc, lane);
      const float cnd = tanh_f(acc[0] + cx_keep + b_c1);
      const float hn = u_keep * hold_keep + (1.0f - u_keep) * cnd;
      y1_keep = y0_keep + hn;                         // ResidualWrapper
      gather_all<S>(hn, vals);
      push_row<S>(in9 + (DH + c) * S, vals, mb + B3_P10, lane);
      gather_all<S>(y1_keep, vals);
      push_row<S>(in11 + c * S, vals, mb + B3_P10, lane);
    }
    // ===== P11: GRU-2 gates (x half) + candidate x-part =====
    load_w<F4_P12>(ws + OFF_P12 * 32, reinterpret_cast<float4(&)[F4_P12]>(wB));
    mbar_wait3(mb + B3_P10, par);
    {
      dot_acc<S, 4, 3, 2>(wA, in11, lane, eacc);
      warp_reduce<S, 4>(eacc, lane);
      const float r = sigmoid_f(eacc[0] + b_r2);
      u_keep = sigmoid_f(eacc[1] + b_u2);
      cx_keep = eacc[2];
      hold_keep = in11[(DH + c) * S + sL];
      gather_all<S>(r * hold_keep, vals);
      push_row<S>(rh2 + c * S, vals, mb + B3_P11, lane);
    }
    // ===== P12: GRU-2 candidate h-part -> h2', y2 = y1 + h2' =====
    load_w<F4_P13>(ws + OFF_P13 * 32, reinterpret_cast<float4(&)[F4_P13]>(wA));
    mbar_wait3(mb + B3_P11, par);
    {
      float acc[S];
#pragma unroll
      for (int i = 0; i < S; ++i) acc[i] = 0.f;
      dot_acc<S, 1, 1, 2>(wB, rh2, lane, acc);
      warp_reduce<S, 1>(acc, lane);
      const float cnd = tanh_f(acc[0] + cx_keep + b_c2);
      const float hn = u_keep * hold_keep + (1.0f - u_keep) * cnd;
      gather_all<S>(hn, vals);
      push_row<S>(in11 + (DH + c) * S, vals, mb + B3_P12, lane);
      gather_all<S>(y1_keep + hn, vals);
      push_row<S>(y2 + c * S, vals, mb + B3_P12, lane);
    }
    // ===== P13: output projection (2 columns per warp), frames out, next decoder input =====
    load_w<F4_P1>(ws + OFF_P1 * 32, reinterpret_cast<float4(&)[F4_P1]>(wB));
    mbar_wait3(mb + B3_P12, par);
    {
      float acc[S * 2];
#pragma unroll
      for (int i = 0; i < S * 2; ++i) acc[i] = 0.f;
      dot_acc<S, 2, 2, 2>(wA, y2, lane, acc);
      warp_reduce<S, 2>(acc, lane);
      const float o0 = acc[0] + b_o0, o1 = acc[1] + b_o1;
      const bool writer = (lane & ((32 >> LS) - 1)) == 0 && nL < a.N;    // one lane per sample
      if (writer) {
        float* orow = a.dec_out + ((size_t)nL * a.max_steps + t) * Dout;
        if (ocol0 < Dout) orow[ocol0] = o0;
        if (ocol0 + 1 < Dout) orow[ocol0 + 1] = o1;
      }
      // next input frame: last of the r frames (helpers.py:37) or the target frame (helpers.py:75)
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int col = ocol0 + cc;
        if (col >= fb0 && col < Dout) {               // warp-uniform
          float f = cc == 0 ? o0 : o1;
          if (!free_run) f = nL < a.N ? __ldg(a.targets + ((size_t)nL * a.T_tgt + (size_t)t * a.r + a.r - 1) * M + (col - fb0)) : 0.f;
          gather_all<S>(f, vals);
          push_row<S>(xin + (col - fb0) * S, vals, mb + B3_P13, lane);
        }
      }
    }
    // rotate the double buffer: wB now holds P1's block of the next step
#pragma unroll
    for (int i = 0; i < F4_P1; ++i) wA[i] = wB[i];
  }
  if (a.steps > 0) mbar_wait3(mb + B3_P13, (uint32_t)(a.steps - 1) & 1u);
  __syncthreads();
  cluster_sync_all();
}

template <int S>
cudaError_t launch_v3_t(const DecoderWeightsV3& w, const DecoderArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)make_layout3(S, a.T_in, a.att_res != 0).total * sizeof(float);
  auto kern = decoder_v3_kernel<S>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(((a.N + S - 1) / S) * CS3);
  cfg.blockDim = dim3(NT3);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS3;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, w, a);
}
}  // namespace

size_t decoder_v3_smem_bytes(int S, int T_in, bool att_res) {
  return (size_t)make_layout3(S, T_in, att_res).total * sizeof(float);
}
int decoder_v3_stream_floats_per_cta() { return NW3 * F4_STEP * 32 * 4; }
int decoder_v3_resident_floats_per_cta() { return NW3 * 3 * F4_E * 32 * 4; }

cudaError_t launch_decoder_v3(const DecoderWeightsV3& w, const DecoderArgs& a_in, int S, cudaStream_t st) {
  if (a_in.N <= 0 || a_in.steps <= 0) return cudaSuccess;
  DecoderArgs a = a_in;
  const char* env = getenv("TACO_DEC_ATT_RES");
  a.att_res = decoder_v3_smem_bytes(S, a.T_in, true) <= 227 * 1024 ? 1 : 0;
  if (env) a.att_res = a.att_res && atoi(env) != 0;
  if (decoder_v3_smem_bytes(S, a.T_in, a.att_res != 0) > 227 * 1024) return cudaErrorInvalidValue;
  switch (S) {
    case 1: return launch_v3_t<1>(w, a, st);
    case 2: return launch_v3_t<2>(w, a, st);
    case 4: return launch_v3_t<4>(w, a, st);
    case 8: return launch_v3_t<8>(w, a, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace taco
