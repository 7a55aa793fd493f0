// K7: the attention decoder loop as one persistent, cluster-resident kernel.
//
// Replaces tf.contrib.seq2seq.dynamic_decode(BasicDecoder(output_cell, helper,
// zero_state), maximum_iterations=max_iters) at reference
// models/tacotron.py:66-94 -- a tf.while_loop of ~60 small ops per step --
// including DecoderPrenetWrapper / ConcatOutputAndAttentionWrapper
// (models/rnn_wrappers.py:22-24,50-52), BahdanauAttention + AttentionWrapper
// (SURVEY.md Appendix B.2), the two ResidualWrapper(GRUCell(256)), the
// OutputProjectionWrapper(80*r) and TacoTestHelper / TacoTrainingHelper
// (models/helpers.py:26-38,68-77).  The whole loop runs on the device.
//
// Decomposition.  A thread-block CLUSTER of CS CTAs (16, or 8 where 16 cannot
// be scheduled) owns S utterances for all steps; clusters never talk to each
// other, so there is no grid-wide barrier.  Inside a cluster every weight
// matrix is cut by output columns: CTA q computes columns [q*Mc,(q+1)*Mc) of
// each of the 13 dependent mat-vec phases of a step for the S samples, then
// PUSHES its slice of the result into the shared memory of all CS CTAs
// (st.shared::cluster) and the cluster meets at a hardware cluster barrier
// (release/acquire).  Activations therefore never leave shared memory; the
// only per-step global traffic is the weight stream (L2-resident, prefetched
// into registers across the barrier), the keys/memory rows of the attention,
// and the outputs.
//
// Per-CTA mat-vec: the slice W_q[K][Mc] is streamed with one 128-bit load per
// (k, 4 columns); a thread keeps 4 x S accumulators, reduces over the threads
// that share its column group with shuffles, then across warps through shared
// memory.  Activations are stored [k][S] so the S samples of one k are one
// vector load.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"

namespace taco {

namespace {
constexpr int NT = 512;       // threads per CTA
constexpr int NW = NT / 32;   // warps per CTA
constexpr int DH = 256;       // decoder width (GRUCell(256), attention depth 256)
constexpr int DP = 128;       // prenet output

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- mbarrier / st.async primitives (PTX, sm_90+) -----------------------------
__device__ __forceinline__ void mbar_init(uint64_t* mb, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mb)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mb, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mb)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mb, uint32_t parity) {
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
// Remote store into a peer CTA's shared memory that also credits `bytes` on the
// peer's mbarrier: the consumer needs no fence and no cluster-wide barrier.
__device__ __forceinline__ void st_async_f4(uint32_t raddr, float4 v, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%2,%3,%4}, [%5];"
               ::"r"(raddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void st_async_f1(uint32_t raddr, float v, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];"
               ::"r"(raddr), "f"(v), "r"(rmbar) : "memory");
}

// ---- streamed mat-vec pieces ------------------------------------------------
// Slice W_q[K][MC] of a weight matrix; thread (kr, cg) owns rows kr, kr+KR, ... and
// the 4 columns of column group cg.  MAXI = rows per thread.
template <int MC, int KMAX>
struct GemmCfg {
  static constexpr int CG = MC / 4;            // column groups (power of two, <= 16)
  static constexpr int KR = NT / CG;           // rows covered per pass
  static constexpr int MAXI = ceil_div(KMAX, KR);
};

template <int MC, int KMAX>
__device__ __forceinline__ void gemm_load(const float* __restrict__ W, int K,
                                          float4 (&w)[GemmCfg<MC, KMAX>::MAXI]) {
  using C = GemmCfg<MC, KMAX>;
  const int tid = threadIdx.x;
  const int cg = tid & (C::CG - 1), kr = tid / C::CG;
#pragma unroll
  for (int i = 0; i < C::MAXI; ++i) {
    const int k = kr + i * C::KR;
    w[i] = (k < K) ? ldg_f4(W + (size_t)k * MC + cg * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Reduce V values per lane over the 32/CG lanes that share a column group by
// recursive halving: each round a lane keeps one half of its values and receives
// the partner's sums for that half, so V values cost ~V shuffles instead of
// V*log2(32/CG).  On return the lane holds NF = max(1, V*CG/32) finished values
// v[0..NF) for flat indices base..base+NF; `writer` is false on duplicate lanes.
template <int N, int OFF, int CG>
struct Halve {
  template <int V>
  static __device__ __forceinline__ void run(float (&v)[V], int lane, int& base, bool& writer) {
    if constexpr (OFF >= CG) {
      const bool up = (lane & OFF) != 0;
      if constexpr (N > 1) {
        constexpr int h = N / 2;
#pragma unroll
        for (int i = 0; i < h; ++i) {
          const float send = up ? v[i] : v[i + h];
          const float keep = up ? v[i + h] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        if (up) base += h;
        Halve<h, OFF / 2, CG>::run(v, lane, base, writer);
      } else {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], OFF);
        if (up) writer = false;
        Halve<1, OFF / 2, CG>::run(v, lane, base, writer);
      }
    }
  }
};

// xs: shared activations [K][S]; red: [NW][S][MC] per-warp partial sums.
template <int S, int MC, int KMAX>
__device__ __forceinline__ void gemm_fma(const float4 (&w)[GemmCfg<MC, KMAX>::MAXI], int K,
                                         const float* __restrict__ xs, float* __restrict__ red) {
  using C = GemmCfg<MC, KMAX>;
  constexpr int V = 4 * S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cg = tid & (C::CG - 1), kr = tid / C::CG;
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
  for (int i = 0; i < C::MAXI; ++i) {
    const int k = kr + i * C::KR;
    if (k < K) {
      float x[S];
      if constexpr (S % 4 == 0) {
#pragma unroll
        for (int s4 = 0; s4 < S / 4; ++s4) {
          const float4 t4 = *reinterpret_cast<const float4*>(xs + (size_t)k * S + s4 * 4);
          x[s4 * 4 + 0] = t4.x; x[s4 * 4 + 1] = t4.y; x[s4 * 4 + 2] = t4.z; x[s4 * 4 + 3] = t4.w;
        }
      } else if constexpr (S == 2) {
        const float2 t2 = *reinterpret_cast<const float2*>(xs + (size_t)k * 2);
        x[0] = t2.x; x[1] = t2.y;
      } else {
#pragma unroll
        for (int s = 0; s < S; ++s) x[s] = xs[(size_t)k * S + s];
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        ffma2(acc[s * 4 + 0], acc[s * 4 + 1], x[s], x[s], w[i].x, w[i].y);
        ffma2(acc[s * 4 + 2], acc[s * 4 + 3], x[s], x[s], w[i].z, w[i].w);
      }
    }
  }
  int base = 0;
  bool writer = true;
  Halve<V, 16, C::CG>::run(acc, lane, base, writer);
  constexpr int NF = (V * C::CG / 32) > 1 ? (V * C::CG / 32) : 1;   // finished values per lane
  if (writer) {
    float* rw = red + (size_t)warp * S * MC + cg * 4;
    if constexpr (NF >= 4) {
#pragma unroll
      for (int i = 0; i < NF; i += 4) {
        const int idx = base + i;   // = s*4 + 0
        *reinterpret_cast<float4*>(rw + (size_t)(idx >> 2) * MC) =
            make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const int idx = base + i;
        rw[(size_t)(idx >> 2) * MC + (idx & 3)] = acc[i];
      }
    }
  }
}

template <int S, int MC>
__device__ __forceinline__ float red_sum(const float* __restrict__ red, int s, int c) {
  float v = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) v += red[(size_t)(w * S + s) * MC + c];
  return v;
}

// Copy `n` floats from local `src` into every CTA of the cluster at the address that
// corresponds to local `dst`, crediting the bytes on each receiver's mbarrier `mb`.
__device__ __forceinline__ void push_block(float* dst, const float* src, int n, int CS, uint64_t* mb) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t d0 = smem_u32(dst), m0 = smem_u32(mb);
  const bool vec = ((n & 3) == 0) && ((d0 & 15u) == 0) && ((smem_u32(src) & 15u) == 0);
  for (int p = warp; p < CS; p += NW) {
    const uint32_t rbase = mapa_u32(d0, (uint32_t)p), rmb = mapa_u32(m0, (uint32_t)p);
    if (vec) {
      for (int i = lane; i < (n >> 2); i += 32)
        st_async_f4(rbase + i * 16, *reinterpret_cast<const float4*>(src + i * 4), rmb);
    } else {
      for (int i = lane; i < n; i += 32) st_async_f1(rbase + i * 4, src[i], rmb);
    }
  }
}

// Shared-memory carve-up (floats).  Everything that peers push into sits at the
// same offset in every CTA of the cluster.
enum { NBAR = 16 };
struct Layout {
  int xin, p1, in3, rhA, pq, sc, in9, rh1, in11, rh2, y2, red, stage, stage2, locu, loccx, locy0h, bias,
      ksl, msl, total;
};
__host__ __device__ inline Layout make_layout(int S, int T_in, int M, int CS, bool att_res) {
  const int Hc = DH / CS;
  Layout L;
  int o = 2 * NBAR;                 // mbarriers (8 bytes each) live at the front
  auto take = [&](int nfloats) { int r = o; o += (nfloats + 3) & ~3; return r; };
  L.xin = take((M + DH) * S);       // [frame(M) | context(256)]      rows x S
  L.p1 = take(DH * S);              // prenet layer 1
  L.in3 = take((DP + DH) * S);      // [prenet out(128) | h_att(256)]
  L.rhA = take(DH * S);             // r * h_att
  L.pq = take(DH * S);              // processed query, layout [S][256]
  L.sc = take(T_in * S);            // scores / alignments [T_in][S]
  L.in9 = take(2 * DH * S);         // [y0 | h1]
  L.rh1 = take(DH * S);
  L.in11 = take(2 * DH * S);        // [y1 | h2]
  L.rh2 = take(DH * S);
  L.y2 = take(DH * S);
  L.red = take(NW * S * 3 * Hc);    // warp partials: gates (2Hc) + candidate-x (Hc)
  L.stage = take(S * 64 > S * 2 * Hc ? S * 64 : S * 2 * Hc);
  L.stage2 = take(S * 64);
  L.locu = take(S * Hc);
  L.loccx = take(S * Hc);
  L.locy0h = take(S * Hc);
  L.bias = take(16 * Hc + 64);
  const int Tj = ceil_div(T_in, CS);
  L.ksl = take(att_res ? S * Tj * DH : 0);      // keys rows [j0,j1) of the S samples
  L.msl = take(att_res ? S * T_in * Hc : 0);    // memory columns [q*Hc,(q+1)*Hc) of the S samples
  L.total = o;
  return L;
}

// barrier slots: mb[p] completes when the outputs of phase p have landed in this CTA
enum { B_P1 = 0, B_P2, B_P3, B_P4, B_P5, B_P6, B_P7, B_P8, B_P9, B_P10, B_P11, B_P12, B_P13 };

// developer aid: clock stamps of (CTA 0, thread 0) at step TRACE_STEP, 64 slots
#define TR(i) do { if (a.trace != nullptr && t == 8 && blockIdx.x == 0 && tid == 0) a.trace[i] = clock64(); } while (0)

template <int S, int CS>
__global__ void __launch_bounds__(NT, 1)
decoder_kernel(const DecoderWeights w, const DecoderArgs a) {
  constexpr int Hc = DH / CS;     // columns of a 256-wide layer per CTA
  constexpr int Pc = DP / CS;     // columns of the 128-wide prenet layer per CTA
  constexpr int McO = (CS == 16) ? 32 : 64;   // padded slice of the 80*r output projection
  constexpr int K1MAX = 128 + DH; // num_mels <= 128
  constexpr uint32_t BLK = CS * Hc * S * 4;   // bytes a CTA receives for one Hc-wide phase

  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = (int)cluster_ctarank();
  const int n0 = (int)cluster_id_x() * S;
  const int M = w.M, K1 = M + DH, Dout = w.Dout;
  const int T_in = a.T_in;
  const bool att_res = a.att_res != 0;
  const Layout L = make_layout(S, T_in, M, CS, att_res);
  uint64_t* mb = reinterpret_cast<uint64_t*>(smem);
  float* xin = smem + L.xin;   float* p1 = smem + L.p1;     float* in3 = smem + L.in3;
  float* rhA = smem + L.rhA;   float* pqT = smem + L.pq;    float* sc = smem + L.sc;
  float* in9 = smem + L.in9;   float* rh1 = smem + L.rh1;   float* in11 = smem + L.in11;
  float* rh2 = smem + L.rh2;   float* y2 = smem + L.y2;     float* red = smem + L.red;
  float* stage = smem + L.stage; float* stage2 = smem + L.stage2;
  float* locu = smem + L.locu; float* loccx = smem + L.loccx; float* locy0h = smem + L.locy0h;
  float* bs = smem + L.bias;
  float* ksl = smem + L.ksl;   float* msl = smem + L.msl;
  float* redB = red + NW * S * 2 * Hc;   // second partial-sum region (candidate x-part)

  for (int i = tid + 2 * NBAR; i < L.total; i += NT) smem[i] = 0.f;   // zero_state + <GO> frame
  if (tid < NBAR) mbar_init(mb + tid, 1);

  // biases of this CTA's columns -> shared memory (they sit on every phase's critical path)
  float* b_p1 = bs;            float* b_ga = b_p1 + Hc;     float* b_ca = b_ga + 2 * Hc;
  float* b_pc = b_ca + Hc;     float* b_g1 = b_pc + Hc;     float* b_c1 = b_g1 + 2 * Hc;
  float* b_g2 = b_c1 + Hc;     float* b_c2 = b_g2 + 2 * Hc; float* b_p2 = b_c2 + Hc;   // [Pc]
  float* b_o = b_p2 + Hc;      // [McO] (<= 64)
  __syncthreads();
  if (tid < Hc) {
    b_p1[tid] = w.p1_b[q * Hc + tid]; b_ca[tid] = w.ca_b[q * Hc + tid]; b_pc[tid] = w.pc_b[q * Hc + tid];
    b_c1[tid] = w.c1_b[q * Hc + tid]; b_c2[tid] = w.c2_b[q * Hc + tid];
  }
  if (tid < 2 * Hc) {
    b_ga[tid] = w.ga_b[q * 2 * Hc + tid]; b_g1[tid] = w.g1_b[q * 2 * Hc + tid]; b_g2[tid] = w.g2_b[q * 2 * Hc + tid];
  }
  if (tid < Pc) b_p2[tid] = w.p2_b[q * Pc + tid];
  if (tid < McO) b_o[tid] = w.o_b[q * McO + tid];

  // per-CTA weight slices
  const float* W1 = w.p1_s + (size_t)q * K1 * Hc;
  const float* W2 = w.p2_s + (size_t)q * DH * Pc;
  const float* WgA = w.ga_s + (size_t)q * (DP + DH) * 2 * Hc;
  const float* WcxA = w.cxa_s + (size_t)q * DP * Hc;
  const float* WchA = w.cha_s + (size_t)q * DH * Hc;
  const float* Wqp = w.qp_s + (size_t)q * DH * 2 * Hc;
  const float* Wpc = w.pc_s + (size_t)q * DH * Hc;
  const float* Wg1 = w.g1_s + (size_t)q * 2 * DH * 2 * Hc;
  const float* Wcx1 = w.cx1_s + (size_t)q * DH * Hc;
  const float* Wch1 = w.ch1_s + (size_t)q * DH * Hc;
  const float* Wg2 = w.g2_s + (size_t)q * 2 * DH * 2 * Hc;
  const float* Wcx2 = w.cx2_s + (size_t)q * DH * Hc;
  const float* Wch2 = w.ch2_s + (size_t)q * DH * Hc;
  const float* Wo = w.o_s + (size_t)q * DH * McO;

  // attention slice of this CTA: encoder positions [j0, j1)
  const int Tj = ceil_div(T_in, CS);
  const int j0 = min(q * Tj, T_in), j1 = min(j0 + Tj, T_in);
  float vreg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) vreg[i] = __ldg(w.att_v + lane + 32 * i);
  if (att_res) {   // keys rows / memory columns this CTA touches every step -> shared memory, once
    for (int i = tid; i < S * (j1 - j0) * (DH / 4); i += NT) {
      const int c4 = i % (DH / 4), r = i / (DH / 4), s = r / (j1 - j0), jj = r - s * (j1 - j0), n = n0 + s;
      const float4 v4 = (n < a.N) ? ldg_f4(a.keys + ((size_t)n * T_in + j0 + jj) * DH + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(ksl + ((size_t)s * Tj + jj) * DH + c4 * 4) = v4;
    }
    for (int i = tid; i < S * T_in * (Hc / 4); i += NT) {
      const int c4 = i % (Hc / 4), r = i / (Hc / 4), s = r / T_in, j = r - s * T_in, n = n0 + s;
      const float4 v4 = (n < a.N) ? ldg_f4(a.memory + ((size_t)n * T_in + j) * DH + q * Hc + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(msl + ((size_t)s * T_in + j) * Hc + c4 * 4) = v4;
    }
  }

  // epilogue role: thread o -> (sample es, column ec) with the column fastest
  const int es = tid / Hc, ec = tid % Hc;        // valid when tid < S*Hc
  const bool e_on = tid < S * Hc;

  // streamed weights of the NEXT phase (loaded right after the current phase's FMA loop)
  float4 wa[GemmCfg<2 * Hc, 2 * DH>::MAXI];      // widest: GRU gates, K=512
  float4 wb[GemmCfg<Hc, K1MAX>::MAXI];           // Hc-wide matrices (K <= 384)
  float4 wo[GemmCfg<McO, DH>::MAXI];
  float4 w2[GemmCfg<Pc, DH>::MAXI];

  gemm_load<Hc, K1MAX>(W1, K1, wb);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  cluster_sync_all();   // buffers zeroed and mbarriers initialised everywhere before anyone pushes

  const bool free_run = a.targets == nullptr;
  for (int t = 0; t < a.steps; ++t) {
    const uint32_t par = (uint32_t)t & 1u;
    if (free_run && t > 0) mbar_wait(mb + B_P13, par ^ 1u);   // fed-back frame of step t-1 has landed
    if (tid == 0) {   // post this step's expected byte counts (remote credits may already have arrived)
      mbar_expect_tx(mb + B_P1, BLK);
      mbar_expect_tx(mb + B_P2, CS * Pc * S * 4);
      mbar_expect_tx(mb + B_P3, BLK);
      mbar_expect_tx(mb + B_P4, BLK);
      mbar_expect_tx(mb + B_P5, BLK);
      mbar_expect_tx(mb + B_P6, (uint32_t)T_in * S * 4);
      mbar_expect_tx(mb + B_P7, BLK);
      mbar_expect_tx(mb + B_P8, BLK);
      mbar_expect_tx(mb + B_P9, BLK);
      mbar_expect_tx(mb + B_P10, 2 * BLK);
      mbar_expect_tx(mb + B_P11, BLK);
      mbar_expect_tx(mb + B_P12, 2 * BLK);
      if (free_run) mbar_expect_tx(mb + B_P13, (uint32_t)M * S * 4);
    }
    // ---- teacher forcing: next input = mel_targets[:, (t-1)*r + r-1, :] (helpers.py:48,75) ----
    if (!free_run) {
      if (t > 0) {
        for (int i = tid; i < M * S; i += NT) {
          const int f = i / S, s = i - f * S, n = n0 + s;
          xin[i] = (n < a.N) ? __ldg(a.targets + ((size_t)n * a.T_tgt + (size_t)(t - 1) * a.r + a.r - 1) * M + f) : 0.f;
        }
      }
      __syncthreads();
    }
    TR(0);
    // ================= P1: decoder prenet dense_1 + ReLU  [frame|ctx] -> 256 =================
    gemm_fma<S, Hc, K1MAX>(wb, K1, xin, red);
    TR(1);
    gemm_load<Pc, DH>(W2, DH, w2);
    __syncthreads();
    TR(2);
    if (e_on) stage[ec * S + es] = fmaxf(red_sum<S, Hc>(red, es, ec) + b_p1[ec], 0.f);
    __syncthreads();
    TR(3);
    push_block(p1 + q * Hc * S, stage, Hc * S, CS, mb + B_P1);
    mbar_wait(mb + B_P1, par);
    TR(4);
    // ================= P2: prenet dense_2 + ReLU  256 -> 128 =================
    gemm_fma<S, Pc, DH>(w2, DH, p1, red);
    TR(5);
    gemm_load<2 * Hc, 2 * DH>(WgA, DP + DH, wa);
    gemm_load<Hc, K1MAX>(WcxA, DP, wb);
    __syncthreads();
    TR(6);
    if (tid < S * Pc) {
      const int s = tid / Pc, c = tid % Pc;
      stage[c * S + s] = fmaxf(red_sum<S, Pc>(red, s, c) + b_p2[c], 0.f);
    }
    __syncthreads();
    TR(7);
    push_block(in3 + q * Pc * S, stage, Pc * S, CS, mb + B_P2);
    mbar_wait(mb + B_P2, par);
    TR(8);
    // ================= P3: attention GRU gates + candidate x-part =================
    gemm_fma<S, 2 * Hc, 2 * DH>(wa, DP + DH, in3, red);
    gemm_fma<S, Hc, K1MAX>(wb, DP, in3, redB);
    TR(9);
    gemm_load<Hc, K1MAX>(WchA, DH, wb);
    __syncthreads();
    TR(10);
    if (e_on) {
      const float r = sigmoid_f(red_sum<S, 2 * Hc>(red, es, ec) + b_ga[ec]);
      const float u = sigmoid_f(red_sum<S, 2 * Hc>(red, es, Hc + ec) + b_ga[Hc + ec]);
      const float hold = in3[(DP + q * Hc + ec) * S + es];
      stage[ec * S + es] = r * hold;
      locu[tid] = u;
      loccx[tid] = red_sum<S, Hc>(redB, es, ec);
    }
    __syncthreads();
    TR(11);
    push_block(rhA + q * Hc * S, stage, Hc * S, CS, mb + B_P3);
    mbar_wait(mb + B_P3, par);
    TR(12);
    // ================= P4: attention GRU candidate h-part -> h_att' =================
    gemm_fma<S, Hc, K1MAX>(wb, DH, rhA, red);
    TR(13);
    gemm_load<2 * Hc, 2 * DH>(Wqp, DH, wa);
    __syncthreads();
    TR(14);
    if (e_on) {
      const float c = tanh_f(red_sum<S, Hc>(red, es, ec) + loccx[tid] + b_ca[ec]);
      const float u = locu[tid];
      const float hold = in3[(DP + q * Hc + ec) * S + es];
      stage[ec * S + es] = u * hold + (1.0f - u) * c;
    }
    __syncthreads();
    TR(15);
    push_block(in3 + (DP + q * Hc) * S, stage, Hc * S, CS, mb + B_P4);
    mbar_wait(mb + B_P4, par);
    TR(16);
    // ================= P5: query layer + h_att' part of the 512->256 projection =================
    gemm_fma<S, 2 * Hc, 2 * DH>(wa, DH, in3 + DP * S, red);
    TR(17);
    gemm_load<Hc, K1MAX>(Wpc, DH, wb);
    __syncthreads();
    TR(18);
    if (e_on) {
      stage[es * Hc + ec] = red_sum<S, 2 * Hc>(red, es, ec);          // layout [S][Hc] for pqT
      locy0h[tid] = red_sum<S, 2 * Hc>(red, es, Hc + ec);
    }
    __syncthreads();
    TR(19);
#pragma unroll
    for (int s = 0; s < S; ++s) push_block(pqT + s * DH + q * Hc, stage + s * Hc, Hc, CS, mb + B_P5);
    mbar_wait(mb + B_P5, par);
    TR(20);
    // ================= P6: Bahdanau scores for positions [j0,j1) =================
    {
      const int npairs = S * (j1 - j0);
      for (int pi = warp; pi < npairs; pi += NW) {
        const int jj = pi / S, s = pi - jj * S, n = n0 + s;
        float e = 0.f;
        const float* prow = pqT + s * DH;
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
        if (lane == 0) stage[jj * S + s] = e;
      }
    }
    __syncthreads();
    TR(21);
    if (j1 > j0) push_block(sc + j0 * S, stage, (j1 - j0) * S, CS, mb + B_P6);
    mbar_wait(mb + B_P6, par);
    TR(22);
    // ================= P7: softmax over all T_in (no mask) + context slice =================
    if (warp < S) {
      const int s = warp;
      float m = -INFINITY;
      for (int j = lane; j < T_in; j += 32) m = fmaxf(m, sc[j * S + s]);
      m = warp_max(m);
      float l = 0.f;
      for (int j = lane; j < T_in; j += 32) {
        const float e = __expf(sc[j * S + s] - m);
        sc[j * S + s] = e;
        l += e;
      }
      l = warp_sum(l);
      const float inv = 1.0f / l;
      for (int j = lane; j < T_in; j += 32) sc[j * S + s] *= inv;
    }
    __syncthreads();
    TR(23);
    if (a.align_out != nullptr) {
      for (int i = tid; i < S * (j1 - j0); i += NT) {
        const int jj = i / S, s = i - jj * S, n = n0 + s;
        if (n < a.N) a.align_out[((size_t)n * T_in + (j0 + jj)) * a.max_steps + t] = sc[(j0 + jj) * S + s];
      }
    }
    {
      constexpr int JC = NT / (S * Hc);              // j-chunks per (sample, dim)
      const int d = tid % Hc, jc = (tid / Hc) % JC, s = tid / (Hc * JC), n = n0 + s;
      float acc0 = 0.f, acc1 = 0.f;
      if (att_res) {
        const float* mp = msl + (size_t)s * T_in * Hc + d;
        int j = jc;
        for (; j + JC < T_in; j += 2 * JC) {
          acc0 = fmaf(sc[j * S + s], mp[(size_t)j * Hc], acc0);
          acc1 = fmaf(sc[(j + JC) * S + s], mp[(size_t)(j + JC) * Hc], acc1);
        }
        if (j < T_in) acc0 = fmaf(sc[j * S + s], mp[(size_t)j * Hc], acc0);
      } else if (n < a.N) {
        const float* mp = a.memory + (size_t)n * T_in * DH + q * Hc + d;
        int j = jc;
        for (; j + 3 * JC < T_in; j += 4 * JC) {
          const float m0 = __ldg(mp + (size_t)j * DH), m1 = __ldg(mp + (size_t)(j + JC) * DH);
          const float m2 = __ldg(mp + (size_t)(j + 2 * JC) * DH), m3 = __ldg(mp + (size_t)(j + 3 * JC) * DH);
          acc0 = fmaf(sc[j * S + s], m0, acc0);
          acc1 = fmaf(sc[(j + JC) * S + s], m1, acc1);
          acc0 = fmaf(sc[(j + 2 * JC) * S + s], m2, acc0);
          acc1 = fmaf(sc[(j + 3 * JC) * S + s], m3, acc1);
        }
        for (; j < T_in; j += JC) acc0 = fmaf(sc[j * S + s], __ldg(mp + (size_t)j * DH), acc0);
      }
      red[(s * JC + jc) * Hc + d] = acc0 + acc1;
      __syncthreads();
      if (e_on) {
        float v = 0.f;
#pragma unroll
        for (int c = 0; c < JC; ++c) v += red[(es * JC + c) * Hc + ec];
        stage[ec * S + es] = v;
      }
    }
    __syncthreads();
    TR(24);
    push_block(xin + (M + q * Hc) * S, stage, Hc * S, CS, mb + B_P7);
    mbar_wait(mb + B_P7, par);
    TR(25);
    // ================= P8: y0 = [h_att'|ctx] W_p + b  (ctx part; h part from P5) =================
    gemm_fma<S, Hc, K1MAX>(wb, DH, xin + M * S, red);
    TR(26);
    gemm_load<2 * Hc, 2 * DH>(Wg1, 2 * DH, wa);
    gemm_load<Hc, K1MAX>(Wcx1, DH, wb);
    __syncthreads();
    TR(27);
    if (e_on) stage[ec * S + es] = red_sum<S, Hc>(red, es, ec) + locy0h[tid] + b_pc[ec];
    __syncthreads();
    TR(28);
    push_block(in9 + q * Hc * S, stage, Hc * S, CS, mb + B_P8);
    mbar_wait(mb + B_P8, par);
    TR(29);
    // ================= P9/P10: residual GRU 1 =================
    gemm_fma<S, 2 * Hc, 2 * DH>(wa, 2 * DH, in9, red);
    gemm_fma<S, Hc, K1MAX>(wb, DH, in9, redB);
    TR(30);
    gemm_load<Hc, K1MAX>(Wch1, DH, wb);
    __syncthreads();
    TR(31);
    if (e_on) {
      const float r = sigmoid_f(red_sum<S, 2 * Hc>(red, es, ec) + b_g1[ec]);
      const float u = sigmoid_f(red_sum<S, 2 * Hc>(red, es, Hc + ec) + b_g1[Hc + ec]);
      stage[ec * S + es] = r * in9[(DH + q * Hc + ec) * S + es];
      locu[tid] = u;
      loccx[tid] = red_sum<S, Hc>(redB, es, ec);
    }
    __syncthreads();
    TR(32);
    push_block(rh1 + q * Hc * S, stage, Hc * S, CS, mb + B_P9);
    mbar_wait(mb + B_P9, par);
    TR(33);
    gemm_fma<S, Hc, K1MAX>(wb, DH, rh1, red);
    TR(34);
    gemm_load<2 * Hc, 2 * DH>(Wg2, 2 * DH, wa);
    gemm_load<Hc, K1MAX>(Wcx2, DH, wb);
    __syncthreads();
    TR(35);
    if (e_on) {
      const float c = tanh_f(red_sum<S, Hc>(red, es, ec) + loccx[tid] + b_c1[ec]);
      const float u = locu[tid];
      const float hold = in9[(DH + q * Hc + ec) * S + es];
      const float hn = u * hold + (1.0f - u) * c;
      stage[ec * S + es] = hn;
      stage2[ec * S + es] = in9[(q * Hc + ec) * S + es] + hn;     // y1 = y0 + GRU1(y0)  (ResidualWrapper)
    }
    __syncthreads();
    TR(36);
    push_block(in9 + (DH + q * Hc) * S, stage, Hc * S, CS, mb + B_P10);
    push_block(in11 + q * Hc * S, stage2, Hc * S, CS, mb + B_P10);
    mbar_wait(mb + B_P10, par);
    TR(37);
    // ================= P11/P12: residual GRU 2 =================
    gemm_fma<S, 2 * Hc, 2 * DH>(wa, 2 * DH, in11, red);
    gemm_fma<S, Hc, K1MAX>(wb, DH, in11, redB);
    TR(38);
    gemm_load<Hc, K1MAX>(Wch2, DH, wb);
    __syncthreads();
    TR(39);
    if (e_on) {
      const float r = sigmoid_f(red_sum<S, 2 * Hc>(red, es, ec) + b_g2[ec]);
      const float u = sigmoid_f(red_sum<S, 2 * Hc>(red, es, Hc + ec) + b_g2[Hc + ec]);
      stage[ec * S + es] = r * in11[(DH + q * Hc + ec) * S + es];
      locu[tid] = u;
      loccx[tid] = red_sum<S, Hc>(redB, es, ec);
    }
    __syncthreads();
    TR(40);
    push_block(rh2 + q * Hc * S, stage, Hc * S, CS, mb + B_P11);
    mbar_wait(mb + B_P11, par);
    TR(41);
    gemm_fma<S, Hc, K1MAX>(wb, DH, rh2, red);
    TR(42);
    gemm_load<McO, DH>(Wo, DH, wo);
    gemm_load<Hc, K1MAX>(W1, K1, wb);
    __syncthreads();
    TR(43);
    if (e_on) {
      const float c = tanh_f(red_sum<S, Hc>(red, es, ec) + loccx[tid] + b_c2[ec]);
      const float u = locu[tid];
      const float hold = in11[(DH + q * Hc + ec) * S + es];
      const float hn = u * hold + (1.0f - u) * c;
      stage[ec * S + es] = hn;
      stage2[ec * S + es] = in11[(q * Hc + ec) * S + es] + hn;    // y2 = y1 + GRU2(y1)
    }
    __syncthreads();
    TR(44);
    push_block(in11 + (DH + q * Hc) * S, stage, Hc * S, CS, mb + B_P12);
    push_block(y2 + q * Hc * S, stage2, Hc * S, CS, mb + B_P12);
    mbar_wait(mb + B_P12, par);
    TR(45);
    // ================= P13: output projection 256 -> 80*r, write frames, feed back =================
    gemm_fma<S, McO, DH>(wo, DH, y2, red);
    __syncthreads();
    TR(46);
    const int fb0 = Dout - M;                                  // first fed-back column (helpers.py:37)
    const int c_lo = max(q * McO, fb0), c_hi = min((q + 1) * McO, Dout);
    if (tid < S * McO) {
      const int s = tid / McO, c = tid % McO, col = q * McO + c, n = n0 + s;
      if (col < Dout) {
        const float v = red_sum<S, McO>(red, s, c) + b_o[c];
        if (n < a.N) a.dec_out[((size_t)n * a.max_steps + t) * Dout + col] = v;
        if (col >= fb0) stage[(col - c_lo) * S + s] = v;
      }
    }
    __syncthreads();
    TR(47);
    if (free_run && c_hi > c_lo) push_block(xin + (c_lo - fb0) * S, stage, (c_hi - c_lo) * S, CS, mb + B_P13);
  }
  // nobody may exit while a peer can still write into its shared memory
  if (free_run && a.steps > 0) mbar_wait(mb + B_P13, (uint32_t)(a.steps - 1) & 1u);
  cluster_sync_all();
}

template <int S, int CS>
cudaError_t launch_decoder_t(const DecoderWeights& w, const DecoderArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)make_layout(S, a.T_in, w.M, CS, a.att_res != 0).total * sizeof(float);
  auto kern = decoder_kernel<S, CS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (CS > 8) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  const int nclusters = ceil_div(a.N, S);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nclusters * CS);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, w, a);
}

template <int CS>
cudaError_t launch_decoder_cs(const DecoderWeights& w, const DecoderArgs& a, int S, cudaStream_t st) {
  switch (S) {
    case 1: return launch_decoder_t<1, CS>(w, a, st);
    case 2: return launch_decoder_t<2, CS>(w, a, st);
    case 4: return launch_decoder_t<4, CS>(w, a, st);
    case 8: return launch_decoder_t<8, CS>(w, a, st);
    default: return cudaErrorInvalidValue;
  }
}

template <int CS>
int max_active_clusters() {
  auto kern = decoder_kernel<1, CS>;
  const size_t smem = (size_t)make_layout(1, 128, 80, CS, false).total * sizeof(float);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (CS > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CS);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

}  // namespace

size_t decoder_smem_bytes(int S, int T_in, int M, int CS, bool att_res) {
  return (size_t)make_layout(S, T_in, M, CS, att_res).total * sizeof(float);
}

int decoder_pick_cluster_size() {
  if (max_active_clusters<16>() >= 1) return 16;
  return 8;
}

int decoder_max_clusters(int CS) { return CS == 16 ? max_active_clusters<16>() : max_active_clusters<8>(); }

cudaError_t launch_decoder(const DecoderWeights& w, const DecoderArgs& a_in, int S, cudaStream_t st) {
  if (a_in.N <= 0 || a_in.steps <= 0) return cudaSuccess;
  DecoderArgs a = a_in;
  // keep this CTA's keys rows / memory columns in shared memory when they fit next to the activations
  const char* env = getenv("TACO_DEC_ATT_RES");
  a.att_res = decoder_smem_bytes(S, a.T_in, w.M, w.CS, true) <= 227 * 1024 ? 1 : 0;
  if (env) a.att_res = a.att_res && atoi(env) != 0;
  if (decoder_smem_bytes(S, a.T_in, w.M, w.CS, a.att_res != 0) > 227 * 1024) return cudaErrorInvalidValue;
  if (w.CS == 16) return launch_decoder_cs<16>(w, a, S, st);
  if (w.CS == 8) return launch_decoder_cs<8>(w, a, S, st);
  return cudaErrorInvalidValue;
}

}  // namespace taco
