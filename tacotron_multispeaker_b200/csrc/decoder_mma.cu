// K7 (v4): the attention decoder loop as one persistent cluster kernel whose mat-vecs run on
// the warp-level tensor-core path (mma.sync m16n8k16, bf16 hi/lo split = fp32-class accuracy).
//
// Operator: tf.contrib.seq2seq.dynamic_decode(BasicDecoder(output_cell, helper, zero_state),
// maximum_iterations=max_iters) of reference models/tacotron.py:66-94 with DecoderPrenetWrapper /
// ConcatOutputAndAttentionWrapper (models/rnn_wrappers.py:22-24,50-52), BahdanauAttention +
// AttentionWrapper, two ResidualWrapper(GRUCell(256)), the 80*r output
// projection and TacoTestHelper / TacoTrainingHelper (models/helpers.py:26-38,68-77).
//
// What the measurements of round 1 said about the previous kernels (decoder.cu, decoder_v3.cu):
// they are bound by instruction issue (~5800 SASS instructions per thread and step, 16 warps)
// and by the cluster exchange (~500-700 clk per all-gather, ~2.5 clk per DSMEM transaction),
// not by bytes.  This kernel is built around those two numbers:
//
//  * a cluster of 16 CTAs owns S <= 8 utterances (S is a RUNTIME value per cluster, so a batch
//    of 32 is cut 5,5,5,5,4,4,4 over the 7 clusters of 16 that fit on a B200 at once);
//  * every weight matrix is cut by output columns into 16-column tiles, one (or two/three) per
//    CTA and phase.  A tile times the S activations is a chain of m16n8k16 MMAs over 16-row
//    chunks of K: A = W^T tile (bf16 hi and lo, pre-packed on the host in register-fragment
//    order, streamed from L2 straight into registers one phase ahead), B = 16 inputs x 8
//    samples (bf16 hi and lo, one LDS.128 per chunk), D = hi*hi + lo*hi + hi*lo in fp32.
//    K is reduced inside the tensor core: no shuffle trees, ~10 instructions per chunk;
//  * warps split the chunks of a phase; their partial tiles meet in shared memory after one
//    block barrier and ONE warp (warp 0) finishes the phase: sum, bias, gate math with the
//    fp32 recurrent state it keeps in registers, hi/lo split, and one 16-byte st.async per lane
//    and peer -- the S*64-byte block of a CTA is contiguous in the receiver's buffer, so a push
//    is 16 DSMEM transactions instead of 256;
//  * activations live in shared memory in MMA-fragment order X[chunk][sample][16 words]
//    (words t*4+{0,1} = hi pairs k=2t,2t+1 / 2t+8,2t+9, words t*4+{2,3} = lo pairs), which is
//    exactly what the producing lane holds, so nothing is ever transposed;
//  * attention: exp(score - B) with B = min(||v||_1, 40) >= score is pushed instead of the raw
//    score, so the softmax needs no max pass and its normaliser is summed with the context.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"

namespace taco {

namespace {
constexpr int CS = 16;         // CTAs per cluster
constexpr int NT = 512;        // threads per CTA
constexpr int NW = 16;         // warps per CTA
constexpr int DH = 256, DP = 128;
constexpr int RS = 20;         // floats per (sample) row of a partial tile (16 + pad, keeps float4 alignment)
constexpr int SLOT = 8 * RS;   // floats per partial tile slot

// chunk-tiles per warp and phase (upper bounds; the real counts come from the table)
constexpr int N1 = DM_NCH1, N2 = 2, N3 = 4, N4 = 2, N5 = 2, N8 = 2, N9 = 6, N10 = 2, N11 = 6, N12 = 2, N13 = 2;
constexpr int O1 = 0, O2 = O1 + 2 * N1, O3 = O2 + 2 * N2, O4 = O3 + 2 * N3, O5 = O4 + 2 * N4, O8 = O5 + 2 * N5,
              O9 = O8 + 2 * N8, O10 = O9 + 2 * N9, O11 = O10 + 2 * N10, O12 = O11 + 2 * N11, O13 = O12 + 2 * N12,
              F4_STEP = O13 + 2 * N13;
static_assert(F4_STEP == DM_F4_STEP, "decoder_mma: stream size out of sync with kernels.cuh");
constexpr int NWB = 2 * N9;    // weight register buffer (uint4)

enum { B_P1 = 0, B_P2, B_P3, B_P4, B_P5, B_P6, B_P7, B_P8, B_P9, B_P10, B_P11, B_P12, B_P13, NBAR = 16 };
// phase index inside DecoderMmaWeights::tab
enum { T_P1 = 0, T_P2, T_P3, T_P4, T_P5, T_P8, T_P9, T_P10, T_P11, T_P12, T_P13 };

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t mb, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mb, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"   // with a suspend-time hint: sleep in hardware, not in a spin
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(mb), "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
               ::"r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t a, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(raddr), "r"(a), "r"(rmbar) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// D += A(16x16, row) * B(16x8, col), bf16 inputs, fp32 accumulate
__device__ __forceinline__ void mma16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// fp32 -> bf16 hi (round to nearest) + bf16 lo (remainder); packs two values per 32-bit word
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
  const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0));
  const __nv_bfloat16 l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
  hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
}

// position of column c (0..15) of a tile inside a 16-float row: the four columns
// {2t, 2t+1, 2t+8, 2t+9} that lane (n, t) finishes sit at floats t*4 .. t*4+3
__host__ __device__ __forceinline__ int pos16(int c) { return ((c & 7) >> 1) * 4 + (c >> 3) * 2 + (c & 1); }

template <int NCH>
__device__ __forceinline__ void load_w(uint4 (&wb)[NWB], const uint4* __restrict__ g, int cnt) {
#pragma unroll
  for (int i = 0; i < NCH; ++i)
    if (i < cnt) {
      wb[2 * i] = __ldg(g + (2 * i) * 32);
      wb[2 * i + 1] = __ldg(g + (2 * i + 1) * 32);
    }
}

// one warp's share of a phase: cnt chunks of one tile; partial tile -> red slot of this warp
template <int NCH>
__device__ __forceinline__ void mma_chunks(const uint4 (&wb)[NWB], uint32_t xaddr, uint32_t csb, int cnt, float* slot,
                                           int g, int t, long long* tr = nullptr) {
  float hh[4] = {0.f, 0.f, 0.f, 0.f}, hl[4] = {0.f, 0.f, 0.f, 0.f}, lh[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < NCH; ++i)
    if (i < cnt) {
      const uint4 xf = lds128(xaddr + i * csb);
      if (tr) { tr[48 + 2 * i] = clock64() + (xf.x & 1u); }
      if (tr) { tr[49 + 2 * i] = clock64() + (wb[2 * i].x & 1u) + (wb[2 * i + 1].x & 1u); }
      mma16816(hh, wb[2 * i], xf.x, xf.y);       // W_hi * x_hi
      mma16816(lh, wb[2 * i + 1], xf.x, xf.y);   // W_lo * x_hi
      mma16816(hl, wb[2 * i], xf.z, xf.w);       // W_hi * x_lo
    }
  // D[row g / g+8 = tile column][col 2t, 2t+1 = sample]  ->  slot[sample][pos16(column)]
  // (written even when cnt == 0 so that the reducer never sums a stale slot)
  const int p = (g >> 1) * 4 + (g & 1);
  if (tr) tr[56] = clock64() + (__float_as_uint(hh[0] + hl[0] + lh[0]) & 1u);
  slot[(2 * t) * RS + p] = hh[0] + (hl[0] + lh[0]);
  slot[(2 * t + 1) * RS + p] = hh[1] + (hl[1] + lh[1]);
  slot[(2 * t) * RS + p + 2] = hh[2] + (hl[2] + lh[2]);
  slot[(2 * t + 1) * RS + p + 2] = hh[3] + (hl[3] + lh[3]);
}

// sum of the partial tiles in slots [s0, s0+ns) for lane (n, t): columns {2t, 2t+1, 2t+8, 2t+9}
template <int NS>
__device__ __forceinline__ float4 red_tile(const float* red, int s0, int n, int t) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(red + (s0 + s) * SLOT + n * RS + t * 4);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  return a;
}

struct Smem {   // byte offsets from the start of dynamic shared memory
  uint32_t buf[DM_NBUF];
  uint32_t pq, sc, red, reds, stage, bias, vatt, ksl, msl, total;
};
__host__ __device__ inline Smem make_smem(int S, int T_in, int FC, bool att_res) {
  Smem L;
  uint32_t o = NBAR * 8;
  auto take = [&](uint32_t bytes) { uint32_t r = o; o += (bytes + 15u) & ~15u; return r; };
  const uint32_t csb = (uint32_t)S * 64u;
  const int nch[DM_NBUF] = {FC, 16, 16, 8, 16, 16, 16, 16, 16, 16, 16, 16, 16};
  for (int b = 0; b < DM_NBUF; ++b) L.buf[b] = take(nch[b] * csb);
  L.pq = take(16 * csb);                     // processed query, fp32, X-like layout [chunk][n][16 floats]
  L.sc = take((uint32_t)T_in * S * 4);       // exp(score - B)  [j][n]
  L.red = take(NW * SLOT * 4);               // partial tiles
  L.reds = take(NW * 8 * 4);                 // partial softmax normalisers
  const int Tj = (T_in + CS - 1) / CS;
  L.stage = take((uint32_t)Tj * S * 4);
  L.bias = take(DM_NBIAS * 4);
  L.vatt = take(DH * 4);
  L.ksl = take(att_res ? (uint32_t)S * Tj * DH * 4 : 0);        // keys rows [j0,j1) of the S samples
  L.msl = take(att_res ? (uint32_t)S * T_in * 16 * 4 : 0);      // memory columns of this CTA, tile order
  L.total = o;
  return L;
}

// bias table (floats, tile order so that lane t reads one float4 at [t*4])
enum { BI_P1 = 0, BI_P2 = 16, BI_RA = 32, BI_UA = 48, BI_CA = 64, BI_PC = 80, BI_R1 = 96, BI_U1 = 112, BI_C1 = 128,
       BI_R2 = 144, BI_U2 = 160, BI_C2 = 176, BI_OA = 192, BI_OB = 208 };
static_assert(BI_OB + 16 == DM_NBIAS, "bias table size");

#define TRM(i) do { if (a.trace != nullptr && step == 8 && blockIdx.x == 0 && tid == 0) a.trace[i] = clock64(); } while (0)

__global__ void __launch_bounds__(NT, 1)
decoder_mma_kernel(const DecoderMmaWeights w, const DecoderArgs a, const int nclusters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;       // MMA fragment coordinates; as reducer: sample g, column group t
  const int q = (int)cluster_ctarank();
  const int cid = (int)cluster_id_x();
  // balanced cut of the batch over the clusters
  const int base = a.N / nclusters, rem = a.N % nclusters;
  const int S = base + (cid < rem ? 1 : 0);
  const int n0 = cid * base + min(cid, rem);
  const int M = w.M, Dout = w.Dout, FC = M >> 4, T_in = a.T_in;
  const bool att_res = a.att_res != 0;
  const Smem L = make_smem(a.s_max, T_in, FC, att_res);   // same carve-up in every cluster of the launch
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t csb = (uint32_t)S * 64u;
  float* red = reinterpret_cast<float*>(smem_raw + L.red);
  float* reds = reinterpret_cast<float*>(smem_raw + L.reds);
  float* stage = reinterpret_cast<float*>(smem_raw + L.stage);
  float* bias = reinterpret_cast<float*>(smem_raw + L.bias);
  float* vatt = reinterpret_cast<float*>(smem_raw + L.vatt);
  float* sc = reinterpret_cast<float*>(smem_raw + L.sc);
  float* ksl = reinterpret_cast<float*>(smem_raw + L.ksl);
  float* msl = reinterpret_cast<float*>(smem_raw + L.msl);
  const uint32_t mb0 = sbase;
  if (S == 0) {   // cannot happen (nclusters <= N) but keeps every CTA of a cluster on the same path
    cluster_sync_all();
    cluster_sync_all();
    return;
  }

  // ---- prologue -------------------------------------------------------------------------
  for (uint32_t i = NBAR * 8 + tid * 4; i < L.total; i += NT * 4) *reinterpret_cast<uint32_t*>(smem_raw + i) = 0u;
  if (tid < NBAR) mbar_init(mb0 + tid * 8, 1);
  __syncthreads();
  if (tid < DM_NBIAS) bias[tid] = __ldg(w.bias + q * DM_NBIAS + tid);
  if (tid < DH) vatt[tid] = __ldg(w.att_v + tid);
  const int Tj = (T_in + CS - 1) / CS;
  const int j0 = min(q * Tj, T_in), j1 = min(j0 + Tj, T_in), nj = j1 - j0;
  if (att_res) {
    for (int i = tid; i < S * nj * (DH / 4); i += NT) {
      const int c4 = i % (DH / 4), r = i / (DH / 4), s = r / nj, jj = r - s * nj;
      *reinterpret_cast<float4*>(ksl + ((size_t)s * Tj + jj) * DH + c4 * 4) =
          ldg_f4(a.keys + ((size_t)(n0 + s) * T_in + j0 + jj) * DH + c4 * 4);
    }
    for (int i = tid; i < S * T_in * 16; i += NT) {
      const int c = i & 15, r = i >> 4, s = r / T_in, j = r - s * T_in;
      msl[(size_t)r * 16 + pos16(c)] = __ldg(a.memory + ((size_t)(n0 + s) * T_in + j) * DH + q * 16 + c);
    }
  }
  float vbound = 0.f;   // B = min(||v||_1, 40) >= any score (|tanh| <= 1)
#pragma unroll
  for (int i = 0; i < 8; ++i) vbound += fabsf(__ldg(w.att_v + lane + 32 * i));
  vbound = fminf(warp_sum(vbound), 40.0f);

  // per-warp work table: tile | buf | chunk0 | count
  auto tab_cnt = [&](int ph) { return (int)(w.tab[ph][warp] & 7u); };
  auto tab_x = [&](int ph) {   // shared address of this lane's first B fragment of the phase
    const uint32_t e = w.tab[ph][warp];
    const uint32_t bufi = (e >> 8) & 15u, c0 = (e >> 3) & 31u;
    return sbase + L.buf[bufi] + c0 * csb + (uint32_t)min(g, S - 1) * 64u + (uint32_t)t * 16u;
  };
  const uint4* ws = reinterpret_cast<const uint4*>(w.stream) + ((size_t)(q * NW + warp) * F4_STEP) * 32 + lane;
  float* myslot = red + warp * SLOT;

  // recurrent state and carried values of the reducer lane (n = g, columns {2t,2t+1,2t+8,2t+9} of this CTA's tile)
  float4 hA = make_float4(0.f, 0.f, 0.f, 0.f), h1 = hA, h2 = hA, ukeep = hA, cxkeep = hA, y0h = hA, y0 = hA, y1 = hA;
  const bool red_on = warp == 0 && g < S;
  // push one finished 4-column group of every active lane into buffer `dst` (byte offset) chunk `chunk` of all peers
  auto push_x = [&](uint32_t dst, int chunk, float4 v, int bar, int pstep, int pfirst) {
    uint32_t h01, l01, h23, l23;
    split2(v.x, v.y, h01, l01);
    split2(v.z, v.w, h23, l23);
    const uint32_t la = sbase + dst + (uint32_t)chunk * csb + (uint32_t)lane * 16u, lm = mb0 + bar * 8;
#pragma unroll
    for (int p = 0; p < CS; ++p)
      if ((p - pfirst) % pstep == 0) st_async_v4(mapa_u32(la, p), h01, h23, l01, l23, mapa_u32(lm, p));
  };
  auto push_f4 = [&](uint32_t dst, int chunk, float4 v, int bar) {   // fp32 block (processed query)
    const uint32_t la = sbase + dst + (uint32_t)chunk * csb + (uint32_t)lane * 16u, lm = mb0 + bar * 8;
#pragma unroll
    for (int p = 0; p < CS; ++p)
      st_async_v4(mapa_u32(la, p), __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w),
                  mapa_u32(lm, p));
  };
  auto bias4 = [&](int off) { return *reinterpret_cast<const float4*>(bias + off + t * 4); };

  uint4 wb[NWB];
  load_w<N1>(wb, ws + O1 * 32, tab_cnt(T_P1));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  cluster_sync_all();   // buffers zeroed and mbarriers initialised everywhere before anyone pushes

  const bool free_run = a.targets == nullptr;
  const uint32_t BLK = CS * csb;
  const int fb_tile0 = (Dout - M) >> 4, ntiles = Dout >> 4;

  for (int step = 0; step < a.steps; ++step) {
    const uint32_t par = (uint32_t)step & 1u;
    if (free_run && step > 0) mbar_wait(mb0 + B_P13 * 8, par ^ 1u);   // fed-back frame of step-1 has landed
    if (tid == 0) {   // this step's expected byte counts (remote credits may already have arrived)
      mbar_expect_tx(mb0 + B_P1 * 8, BLK);
      mbar_expect_tx(mb0 + B_P2 * 8, BLK / 2);
      mbar_expect_tx(mb0 + B_P3 * 8, BLK);
      mbar_expect_tx(mb0 + B_P4 * 8, BLK);
      mbar_expect_tx(mb0 + B_P5 * 8, BLK);
      mbar_expect_tx(mb0 + B_P6 * 8, (uint32_t)T_in * S * 4u);
      mbar_expect_tx(mb0 + B_P7 * 8, BLK);
      mbar_expect_tx(mb0 + B_P8 * 8, BLK);
      mbar_expect_tx(mb0 + B_P9 * 8, BLK);
      mbar_expect_tx(mb0 + B_P10 * 8, 2 * BLK);
      mbar_expect_tx(mb0 + B_P11 * 8, BLK);
      mbar_expect_tx(mb0 + B_P12 * 8, 2 * BLK);
      if (free_run) mbar_expect_tx(mb0 + B_P13 * 8, (uint32_t)FC * csb);
    }
    if (!free_run) {   // teacher forcing: input = mel_targets[:, (step-1)*r + r-1, :]  (helpers.py:48,75)
      if (step > 0 && tid < S * FC * 4) {
        const int tt = tid & 3, r2 = tid >> 2, n = r2 % S, ch = r2 / S;
        const float* src = a.targets + ((size_t)(n0 + n) * a.T_tgt + (size_t)(step - 1) * a.r + a.r - 1) * M + ch * 16;
        const float2 lo2 = __ldg(reinterpret_cast<const float2*>(src + 2 * tt));
        const float2 hi2 = __ldg(reinterpret_cast<const float2*>(src + 2 * tt + 8));
        uint32_t h01, l01, h23, l23;
        split2(lo2.x, lo2.y, h01, l01);
        split2(hi2.x, hi2.y, h23, l23);
        sts128(sbase + L.buf[DM_BF] + ch * csb + n * 64 + tt * 16, h01, h23, l01, l23);
      }
      __syncthreads();
    }
    TRM(0);
    // ================= P1: decoder prenet dense_1 + ReLU on [frame | context] =================
    mma_chunks<N1>(wb, tab_x(T_P1), csb, tab_cnt(T_P1), myslot, g, t);
    load_w<N2>(wb, ws + O2 * 32, tab_cnt(T_P2));
    __syncthreads();
    TRM(1);
    if (red_on) {
      float4 v = red_tile<DM_P1_SLOTS>(red, 0, g, t);
      const float4 b = bias4(BI_P1);
      v.x = fmaxf(v.x + b.x, 0.f); v.y = fmaxf(v.y + b.y, 0.f); v.z = fmaxf(v.z + b.z, 0.f); v.w = fmaxf(v.w + b.w, 0.f);
      push_x(L.buf[DM_BP1], q, v, B_P1, 1, 0);
    }
    TRM(2);
    mbar_wait(mb0 + B_P1 * 8, par);
    TRM(3);
    // ================= P2: prenet dense_2 + ReLU (CTA pair 2c, 2c+1 computes chunk c; each feeds half the peers) ====
    mma_chunks<N2>(wb, tab_x(T_P2), csb, tab_cnt(T_P2), myslot, g, t);
    load_w<N3>(wb, ws + O3 * 32, tab_cnt(T_P3));
    __syncthreads();
    TRM(4);
    if (red_on) {
      float4 v = red_tile<8>(red, 0, g, t);
      const float4 b = bias4(BI_P2);
      v.x = fmaxf(v.x + b.x, 0.f); v.y = fmaxf(v.y + b.y, 0.f); v.z = fmaxf(v.z + b.z, 0.f); v.w = fmaxf(v.w + b.w, 0.f);
      push_x(L.buf[DM_BP2], q >> 1, v, B_P2, 2, q & 1);
    }
    TRM(5);
    mbar_wait(mb0 + B_P2 * 8, par);
    TRM(6);
    // ================= P3: attention GRU gates r,u on [prenet | h_att] and candidate x-part =================
    mma_chunks<N3>(wb, tab_x(T_P3), csb, tab_cnt(T_P3), myslot, g, t);
    load_w<N4>(wb, ws + O4 * 32, tab_cnt(T_P4));
    __syncthreads();
    TRM(7);
    if (red_on) {
      const float4 r = red_tile<6>(red, 0, g, t), u = red_tile<6>(red, 6, g, t);
      cxkeep = red_tile<4>(red, 12, g, t);
      const float4 br = bias4(BI_RA), bu = bias4(BI_UA);
      ukeep = make_float4(sigmoid_f(u.x + bu.x), sigmoid_f(u.y + bu.y), sigmoid_f(u.z + bu.z), sigmoid_f(u.w + bu.w));
      const float4 rh = make_float4(sigmoid_f(r.x + br.x) * hA.x, sigmoid_f(r.y + br.y) * hA.y,
                                    sigmoid_f(r.z + br.z) * hA.z, sigmoid_f(r.w + br.w) * hA.w);
      push_x(L.buf[DM_BRA], q, rh, B_P3, 1, 0);
    }
    TRM(8);
    mbar_wait(mb0 + B_P3 * 8, par);
    TRM(9);
    // ================= P4: candidate h-part -> h_att' =================
    {
      const uint32_t xa4 = tab_x(T_P4);
      const int cn4 = tab_cnt(T_P4);
      TRM(40);
      mma_chunks<N4>(wb, xa4, csb, cn4, myslot, g, t, (a.trace != nullptr && step == 8 && blockIdx.x == 0 && tid == 0) ? a.trace : nullptr);
      TRM(41);
      load_w<N5>(wb, ws + O5 * 32, tab_cnt(T_P5));
      TRM(42);
    }
    __syncthreads();
    TRM(10);
    if (red_on) {
      const float4 c = red_tile<8>(red, 0, g, t), b = bias4(BI_CA);
      if (a.trace != nullptr && step == 8 && blockIdx.x == 0 && tid == 0) a.trace[43] = clock64() + (__float_as_uint(c.x) & 1u);
      hA.x = ukeep.x * hA.x + (1.0f - ukeep.x) * tanh_f(c.x + cxkeep.x + b.x);
      hA.y = ukeep.y * hA.y + (1.0f - ukeep.y) * tanh_f(c.y + cxkeep.y + b.y);
      hA.z = ukeep.z * hA.z + (1.0f - ukeep.z) * tanh_f(c.z + cxkeep.z + b.z);
      hA.w = ukeep.w * hA.w + (1.0f - ukeep.w) * tanh_f(c.w + cxkeep.w + b.w);
      if (a.trace != nullptr && step == 8 && blockIdx.x == 0 && tid == 0) a.trace[44] = clock64() + (__float_as_uint(hA.x) & 1u);
      push_x(L.buf[DM_BHA], q, hA, B_P4, 1, 0);
    }
    TRM(11);
    mbar_wait(mb0 + B_P4 * 8, par);
    TRM(12);
    // ================= P5: query layer (tile A) and h_att' half of the 512->256 projection (tile B) =================
    mma_chunks<N5>(wb, tab_x(T_P5), csb, tab_cnt(T_P5), myslot, g, t);
    load_w<N8>(wb, ws + O8 * 32, tab_cnt(T_P8));
    __syncthreads();
    TRM(13);
    if (red_on) {
      const float4 pq = red_tile<8>(red, 0, g, t);
      y0h = red_tile<8>(red, 8, g, t);
      push_f4(L.pq, q, pq, B_P5);
    }
    TRM(14);
    mbar_wait(mb0 + B_P5 * 8, par);
    TRM(15);
    // ================= P6: Bahdanau scores of positions [j0,j1): exp(v . tanh(keys + pq) - B) =================
    {
      const int npairs = S * nj;
      // lane's 8 inputs k = lane + 32 i sit in chunk (lane>>4) + 2i at tile position pos16(lane & 15)
      const int pqoff = (lane >> 4) * (int)(csb >> 2) + pos16(lane & 15);
      const float* pqb = reinterpret_cast<const float*>(smem_raw + L.pq);
      for (int pi = warp; pi < npairs; pi += NW) {
        const int jj = pi / S, s = pi - jj * S;
        const float* prow = pqb + pqoff + s * 16;
        float e = 0.f;
        if (att_res) {
          const float* krow = ksl + ((size_t)s * Tj + jj) * DH + lane;
#pragma unroll
          for (int i = 0; i < 8; ++i) e = fmaf(vatt[lane + 32 * i], tanh_f(krow[32 * i] + prow[i * 2 * (int)(csb >> 2)]), e);
        } else {
          const float* krow = a.keys + ((size_t)(n0 + s) * T_in + (j0 + jj)) * DH + lane;
          float kv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) kv[i] = __ldg(krow + 32 * i);
#pragma unroll
          for (int i = 0; i < 8; ++i) e = fmaf(vatt[lane + 32 * i], tanh_f(kv[i] + prow[i * 2 * (int)(csb >> 2)]), e);
        }
        e = warp_sum(e);
        if (lane == 0) stage[jj * S + s] = __expf(fmaxf(e - vbound, -80.0f));
      }
    }
    __syncthreads();
    TRM(16);
    if (warp == 0 && nj > 0) {
      const uint32_t la = sbase + L.sc + (uint32_t)(j0 * S) * 4u, lm = mb0 + B_P6 * 8;
      for (int i = lane; i < nj * S; i += 32) {
        const uint32_t v = __float_as_uint(stage[i]);
#pragma unroll
        for (int p = 0; p < CS; ++p) st_async_b32(mapa_u32(la + i * 4, p), v, mapa_u32(lm, p));
      }
    }
    TRM(17);
    mbar_wait(mb0 + B_P6 * 8, par);
    TRM(18);
    // ================= P7: context slice sum_j p_j memory[j][16q..16q+15] / sum_j p_j =================
    {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      float ssum = 0.f;
      if (g < S) {
        for (int j = warp; j < T_in; j += NW) {
          const float p = sc[j * S + g];
          float4 m;
          if (att_res) {
            m = *reinterpret_cast<const float4*>(msl + ((size_t)g * T_in + j) * 16 + t * 4);
          } else {
            const float* mp = a.memory + ((size_t)(n0 + g) * T_in + j) * DH + q * 16 + 2 * t;
            const float2 m01 = __ldg(reinterpret_cast<const float2*>(mp));
            const float2 m89 = __ldg(reinterpret_cast<const float2*>(mp + 8));
            m = make_float4(m01.x, m01.y, m89.x, m89.y);
          }
          acc.x = fmaf(p, m.x, acc.x); acc.y = fmaf(p, m.y, acc.y); acc.z = fmaf(p, m.z, acc.z); acc.w = fmaf(p, m.w, acc.w);
          ssum += p;
        }
      }
      *reinterpret_cast<float4*>(myslot + g * RS + t * 4) = acc;
      if (t == 0) reds[warp * 8 + g] = ssum;
    }
    __syncthreads();
    TRM(19);
    if (warp == 0) {
      float inv = 0.f;
      if (g < S) {
        const float4 c = red_tile<16>(red, 0, g, t);
        float ssum = 0.f;
#pragma unroll
        for (int s = 0; s < NW; ++s) ssum += reds[s * 8 + g];
        inv = 1.0f / ssum;
        push_x(L.buf[DM_BC], q, make_float4(c.x * inv, c.y * inv, c.z * inv, c.w * inv), B_P7, 1, 0);
      }
      if (a.align_out != nullptr) {   // alignments of this CTA's positions (tacotron.py:104 layout [N,T_in,steps])
        for (int i0 = 0; i0 < nj * 8; i0 += 32) {
          const int i = i0 + lane, jj = i >> 3, s = i & 7;
          const float iv = __shfl_sync(0xffffffffu, inv, (s < S ? s : 0) * 4);
          if (jj < nj && s < S) a.align_out[((size_t)(n0 + s) * T_in + (j0 + jj)) * a.max_steps + step] = stage[jj * S + s] * iv;
        }
      }
    }
    TRM(20);
    mbar_wait(mb0 + B_P7 * 8, par);
    TRM(21);
    // ================= P8: y0 = [h_att' | ctx] W_p + b (ctx half here, h half from P5) =================
    mma_chunks<N8>(wb, tab_x(T_P8), csb, tab_cnt(T_P8), myslot, g, t);
    load_w<N9>(wb, ws + O9 * 32, tab_cnt(T_P9));
    __syncthreads();
    TRM(22);
    if (red_on) {
      const float4 c = red_tile<8>(red, 0, g, t), b = bias4(BI_PC);
      y0 = make_float4(c.x + y0h.x + b.x, c.y + y0h.y + b.y, c.z + y0h.z + b.z, c.w + y0h.w + b.w);
      push_x(L.buf[DM_BY0], q, y0, B_P8, 1, 0);
    }
    TRM(23);
    mbar_wait(mb0 + B_P8 * 8, par);
    TRM(24);
    // ================= P9: decoder GRU 1 gates on [y0 | h1] + candidate x-part =================
    mma_chunks<N9>(wb, tab_x(T_P9), csb, tab_cnt(T_P9), myslot, g, t);
    load_w<N10>(wb, ws + O10 * 32, tab_cnt(T_P10));
    __syncthreads();
    TRM(25);
    if (red_on) {
      const float4 r = red_tile<6>(red, 0, g, t), u = red_tile<6>(red, 6, g, t);
      cxkeep = red_tile<4>(red, 12, g, t);
      const float4 br = bias4(BI_R1), bu = bias4(BI_U1);
      ukeep = make_float4(sigmoid_f(u.x + bu.x), sigmoid_f(u.y + bu.y), sigmoid_f(u.z + bu.z), sigmoid_f(u.w + bu.w));
      const float4 rh = make_float4(sigmoid_f(r.x + br.x) * h1.x, sigmoid_f(r.y + br.y) * h1.y,
                                    sigmoid_f(r.z + br.z) * h1.z, sigmoid_f(r.w + br.w) * h1.w);
      push_x(L.buf[DM_BR1], q, rh, B_P9, 1, 0);
    }
    TRM(26);
    mbar_wait(mb0 + B_P9 * 8, par);
    TRM(27);
    // ================= P10: GRU 1 candidate h-part -> h1', y1 = y0 + h1' (ResidualWrapper) =================
    mma_chunks<N10>(wb, tab_x(T_P10), csb, tab_cnt(T_P10), myslot, g, t);
    load_w<N11>(wb, ws + O11 * 32, tab_cnt(T_P11));
    __syncthreads();
    TRM(28);
    if (red_on) {
      const float4 c = red_tile<8>(red, 0, g, t), b = bias4(BI_C1);
      h1.x = ukeep.x * h1.x + (1.0f - ukeep.x) * tanh_f(c.x + cxkeep.x + b.x);
      h1.y = ukeep.y * h1.y + (1.0f - ukeep.y) * tanh_f(c.y + cxkeep.y + b.y);
      h1.z = ukeep.z * h1.z + (1.0f - ukeep.z) * tanh_f(c.z + cxkeep.z + b.z);
      h1.w = ukeep.w * h1.w + (1.0f - ukeep.w) * tanh_f(c.w + cxkeep.w + b.w);
      y1 = make_float4(y0.x + h1.x, y0.y + h1.y, y0.z + h1.z, y0.w + h1.w);
      push_x(L.buf[DM_BH1], q, h1, B_P10, 1, 0);
      push_x(L.buf[DM_BY1], q, y1, B_P10, 1, 0);
    }
    TRM(29);
    mbar_wait(mb0 + B_P10 * 8, par);
    TRM(30);
    // ================= P11: decoder GRU 2 gates on [y1 | h2] + candidate x-part =================
    mma_chunks<N11>(wb, tab_x(T_P11), csb, tab_cnt(T_P11), myslot, g, t);
    load_w<N12>(wb, ws + O12 * 32, tab_cnt(T_P12));
    __syncthreads();
    TRM(31);
    if (red_on) {
      const float4 r = red_tile<6>(red, 0, g, t), u = red_tile<6>(red, 6, g, t);
      cxkeep = red_tile<4>(red, 12, g, t);
      const float4 br = bias4(BI_R2), bu = bias4(BI_U2);
      ukeep = make_float4(sigmoid_f(u.x + bu.x), sigmoid_f(u.y + bu.y), sigmoid_f(u.z + bu.z), sigmoid_f(u.w + bu.w));
      const float4 rh = make_float4(sigmoid_f(r.x + br.x) * h2.x, sigmoid_f(r.y + br.y) * h2.y,
                                    sigmoid_f(r.z + br.z) * h2.z, sigmoid_f(r.w + br.w) * h2.w);
      push_x(L.buf[DM_BR2], q, rh, B_P11, 1, 0);
    }
    TRM(32);
    mbar_wait(mb0 + B_P11 * 8, par);
    TRM(33);
    // ================= P12: GRU 2 candidate h-part -> h2', y2 = y1 + h2' =================
    mma_chunks<N12>(wb, tab_x(T_P12), csb, tab_cnt(T_P12), myslot, g, t);
    load_w<N13>(wb, ws + O13 * 32, tab_cnt(T_P13));
    __syncthreads();
    TRM(34);
    if (red_on) {
      const float4 c = red_tile<8>(red, 0, g, t), b = bias4(BI_C2);
      h2.x = ukeep.x * h2.x + (1.0f - ukeep.x) * tanh_f(c.x + cxkeep.x + b.x);
      h2.y = ukeep.y * h2.y + (1.0f - ukeep.y) * tanh_f(c.y + cxkeep.y + b.y);
      h2.z = ukeep.z * h2.z + (1.0f - ukeep.z) * tanh_f(c.z + cxkeep.z + b.z);
      h2.w = ukeep.w * h2.w + (1.0f - ukeep.w) * tanh_f(c.w + cxkeep.w + b.w);
      push_x(L.buf[DM_BH2], q, h2, B_P12, 1, 0);
      push_x(L.buf[DM_BY2], q, make_float4(y1.x + h2.x, y1.y + h2.y, y1.z + h2.z, y1.w + h2.w), B_P12, 1, 0);
    }
    TRM(35);
    mbar_wait(mb0 + B_P12 * 8, par);
    TRM(36);
    // ================= P13: output projection tiles 2q, 2q+1 -> frames, feed the last frame back =================
    mma_chunks<N13>(wb, tab_x(T_P13), csb, tab_cnt(T_P13), myslot, g, t);
    load_w<N1>(wb, ws + O1 * 32, tab_cnt(T_P1));
    __syncthreads();
    TRM(37);
    if (red_on) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int tile = 2 * q + half;
        if (tile < ntiles) {
          float4 o = red_tile<8>(red, half * 8, g, t);
          const float4 b = bias4(half ? BI_OB : BI_OA);
          o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          float* orow = a.dec_out + ((size_t)(n0 + g) * a.max_steps + step) * Dout + tile * 16 + 2 * t;
          *reinterpret_cast<float2*>(orow) = make_float2(o.x, o.y);
          *reinterpret_cast<float2*>(orow + 8) = make_float2(o.z, o.w);
          if (free_run && tile >= fb_tile0) push_x(L.buf[DM_BF], tile - fb_tile0, o, B_P13, 1, 0);   // helpers.py:37
        }
      }
    }
    TRM(38);
  }
  // nobody may exit while a peer can still write into its shared memory
  if (free_run && a.steps > 0) mbar_wait(mb0 + B_P13 * 8, (uint32_t)(a.steps - 1) & 1u);
  cluster_sync_all();
}

}  // namespace

size_t decoder_mma_smem_bytes(int s_max, int T_in, int M, bool att_res) {
  return make_smem(s_max, T_in, M >> 4, att_res).total;
}

int decoder_mma_max_clusters() {
  auto kern = decoder_mma_kernel;
  const int smem = 200 * 1024;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
      cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CS);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

cudaError_t launch_decoder_mma(const DecoderMmaWeights& w, const DecoderArgs& a_in, int nclusters, cudaStream_t st) {
  if (a_in.N <= 0 || a_in.steps <= 0) return cudaSuccess;
  if (nclusters < 1 || nclusters > a_in.N) return cudaErrorInvalidValue;
  DecoderArgs a = a_in;
  a.s_max = (a.N + nclusters - 1) / nclusters;
  if (a.s_max > 8 || (w.M & 15) || (w.Dout & 15) || w.M > 128) return cudaErrorInvalidValue;
  const char* env = getenv("TACO_DEC_ATT_RES");
  a.att_res = decoder_mma_smem_bytes(a.s_max, a.T_in, w.M, true) <= 227 * 1024 ? 1 : 0;
  if (env) a.att_res = a.att_res && atoi(env) != 0;
  const size_t smem = decoder_mma_smem_bytes(a.s_max, a.T_in, w.M, a.att_res != 0);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  auto kern = decoder_mma_kernel;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nclusters * CS);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, w, a, nclusters);
}

}  // namespace taco
