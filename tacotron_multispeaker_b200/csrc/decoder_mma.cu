// K7 (v5): the attention decoder loop as one persistent cluster kernel whose mat-vecs run on
// the warp-level tensor-core path (mma.sync m16n8k16, bf16 hi/lo split = fp32-class accuracy).
//
// Operator: tf.contrib.seq2seq.dynamic_decode(BasicDecoder(output_cell, helper, zero_state),
// maximum_iterations=max_iters) of reference models/tacotron.py:66-94 with DecoderPrenetWrapper /
// ConcatOutputAndAttentionWrapper (models/rnn_wrappers.py:22-24,50-52), BahdanauAttention +
// AttentionWrapper, two ResidualWrapper(GRUCell(256)), the 80*r output
// projection and TacoTestHelper / TacoTrainingHelper (models/helpers.py:26-38,68-77).
//
// The earlier kernels (decoder.cu, decoder_v3.cu) are bound by instruction issue and by the cluster
// exchange (~500-700 clk per all-gather), not by bytes.  This one is built around the measurements
// listed in DESIGN.md section 4 (K7):
//
//  * a cluster of 16 CTAs owns S <= 8 utterances (S is a RUNTIME value per cluster, so a batch
//    of 32 is cut 5,5,5,5,4,4,4 over the 7 clusters of 16 that fit on a B200 at once);
//  * every weight matrix is cut by output columns into 16-column tiles, one (or two/three) per
//    CTA and phase.  A tile times the S activations is a chain of m16n8k16 MMAs over 16-row
//    chunks of K: A = W^T tile (bf16 hi and lo, pre-packed on the host in register-fragment
//    order: a 1 KB "chunk-tile"), B = 16 inputs x 8 samples (bf16 hi and lo, one LDS.128 per
//    chunk, requested one chunk ahead), D = hi*hi + lo*hi + hi*lo in fp32;
//  * the chunk-tiles of both decoder GRUs live in TENSOR MEMORY (written once with tcgen05.st,
//    fetched with tcgen05.ld.32x32b.x8 into the A registers 40-50 clk before use); the rest is
//    re-read from L2 by cp.async into a per-warp ring in shared memory whose groups are counted
//    explicitly (loads straight into registers are tied to the register scoreboards and cannot
//    run more than one exchange window ahead);
//  * operands that are complete before a phase's exchange (recurrent states, the context of the
//    previous step, y0 / h1' for the sums y1, y2) are multiplied in an EARLIER exchange window
//    (mma_split<0>); the late part after the wait adds only what has just arrived.  Operand
//    passes are skipped with real branches: a predicated-off HMMA still occupies the tensor pipe;
//  * warps split the chunks of a phase (host-side work table); their partial tiles meet in shared
//    memory after ONE block barrier, reducer groups on the least loaded warps (12-15 / 8-11 / 4-7)
//    finish the phase -- sum, bias, gate math on the fp32 recurrent state, hi/lo split -- and each
//    reducer warp pushes the rows it staged itself to all 16 peers (st.async + mbarrier tx bytes);
//  * the warp index comes from a shuffle (ptxas then proves it uniform: no WARPSYNC / BSSY around
//    per-warp decisions), every shared-memory region starts on a 128-byte line, the clock-stamp
//    tracing is a separate template instantiation;
//  * activations live in shared memory in MMA-fragment order X[chunk][sample][16 words]
//    (words t*4+{0,1} = hi pairs k=2t,2t+1 / 2t+8,2t+9, words t*4+{2,3} = lo pairs), which is
//    exactly what the producing lane holds, so nothing is ever transposed;
//  * attention: v.tanh(k+p) through e^{2k} e^{2p} with one MUFU.RCP per four elements; exp(score - B)
//    with B = min(||v||_1, 40) >= score is pushed instead of the raw score (16-byte words), so the
//    softmax needs no max pass and its normaliser is summed with the context.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"

namespace taco {

namespace {
constexpr int CS = 16;         // CTAs per cluster
constexpr int NT = 512;        // threads per CTA
constexpr int NW = 16;         // warps per CTA
constexpr int DH = 256;
constexpr int RS = 20;         // floats per (sample) row of a partial tile (16 + pad, keeps float4 alignment)
constexpr int SLOT = 8 * RS;   // floats per partial tile slot

// chunk-tiles per warp and phase (upper bounds; the real counts come from the table)
constexpr int N1 = DM_NCH1, N2 = 2, N3 = 4, N4 = 2, N5 = 2, N8 = 2, N9 = 6, N10 = 2, N11 = 6, N12 = 2, N13 = 2;
constexpr int O1 = 0, O2 = O1 + 2 * N1, O3 = O2 + 2 * N2, O4 = O3 + 2 * N3, O5 = O4 + 2 * N4, O8 = O5 + 2 * N5,
              O9 = O8 + 2 * N8, O10 = O9 + 2 * N9, O11 = O10 + 2 * N10, O12 = O11 + 2 * N11, O13 = O12 + 2 * N12,
              F4_STEP = O13 + 2 * N13;
static_assert(F4_STEP == DM_F4_STEP, "decoder_mma: stream size out of sync with kernels.cuh");
constexpr int NWB = 12;        // weight registers: six slots of (hi, lo) uint4

// register slots of the chunks of each phase
template <int... SL> struct Slots {};
using SL1 = Slots<4, 5, 0>;
using SL2 = Slots<1, 2>;
using SL3 = Slots<3, 4, 5, 0>;
using SL4 = Slots<1, 2>;
using SL5 = Slots<3, 4>;
using SL8 = Slots<0, 1>;
using SL9 = Slots<2, 3, 4, 5, 0, 1>;
using SL10 = Slots<0, 1>;
using SL11 = Slots<2, 3, 4, 5, 0, 1>;
using SL12 = Slots<0, 1>;
using SL13 = Slots<2, 3>;
static_assert(N1 == 3 && N3 == 4 && N9 == 6 && N11 == 6, "slot lists assume these chunk counts");
// chunk-tile index of the TMEM-resident phases inside a warp's 128-column slice (8 columns per chunk-tile)
constexpr int TC9 = 0, TC10 = TC9 + N9, TC11 = TC10 + N10, TC12 = TC11 + N11;
static_assert(TC12 + N12 == 16, "the four GRU phases fill the 16 chunk-tiles per warp that TMEM holds");

enum { B_P1 = 0, B_P2, B_P3, B_P4, B_P5, B_P6, B_P7, B_P8, B_P9, B_P10, B_P11, B_P12, B_P13, NBAR = 16 };
// phase index inside DecoderMmaWeights::tab
enum { T_P1 = 0, T_P2, T_P3, T_P4, T_P5, T_P8, T_P9, T_P10, T_P11, T_P12, T_P13 };
// reducer state slots (float4 per lane of warp 0)
enum { ST_HA = 0, ST_H1, ST_H2, ST_U, ST_CX, ST_Y0H, ST_Y0, ST_Y1, NSTATE };

// ---- shared memory: fixed part at compile-time offsets, then the S-dependent buffers ------------
constexpr uint32_t OFF_MBAR = 0;
constexpr uint32_t OFF_WTAB = OFF_MBAR + NBAR * 8;                 // [11][16] (x offset << 3 | count)
constexpr uint32_t OFF_RED = OFF_WTAB + DM_NPHASE * 16 * 4;        // partial tiles [16][8][RS]
constexpr uint32_t OFF_REDS = OFF_RED + NW * SLOT * 4;             // partial softmax normalisers [16][8]
constexpr uint32_t OFF_BIAS = OFF_REDS + NW * 8 * 4;
constexpr uint32_t OFF_VATT = OFF_BIAS + DM_NBIAS * 4;
constexpr uint32_t OFF_STATE = OFF_VATT + DH * 4;                  // [NSTATE][8 samples][16 columns] fp32
constexpr uint32_t OFF_STG = OFF_STATE + NSTATE * 128 * 4;         // two staged blocks of 8 x 64 bytes
constexpr uint32_t OFF_INV = OFF_STG + 2 * 512;                    // softmax normalisers 1/sum per sample
constexpr uint32_t OFF_P6T = OFF_INV + 32;                         // per-warp pair geometry of the score phase
constexpr uint32_t OFF_TMEM = OFF_P6T + 64;                       // TMEM base address written by tcgen05.alloc
constexpr uint32_t OFF_VB = OFF_TMEM + 4;                         // sum_k v_k - min(||v||_1, 40) (read once per step)
constexpr uint32_t OFF_SEQ = OFF_TMEM + 16;                       // per-warp order of the streamed chunk-tiles [16][16]
constexpr uint32_t OFF_X = (OFF_SEQ + NW * 16 * 4 + 127u) & ~127u;
static_assert(OFF_X % 16 == 0 && OFF_RED % 16 == 0 && OFF_STATE % 16 == 0 && OFF_BIAS % 16 == 0, "alignment");

// chunks in front of buffer b (order DM_BF, DM_BC, DM_BP1, DM_BP2(8 chunks), DM_BHA, ...)
// (DM_BY1 and DM_BY2 take no space: y1 = y0 + h1' and y2 = y1 + h2' are never formed, their terms are multiplied)
__host__ __device__ __forceinline__ int cum_chunks(int b, int FC) {
  return b == 0 ? 0 : FC + 16 * (b - 1) - (b > DM_BP2 ? 8 : 0) - (b > DM_BY1 ? 16 : 0) - (b > DM_BY2 ? 16 : 0);
}
struct Dyn { uint32_t pq, sc, stage, pn, ksl, msl, ring, total; };
// att_res: bit 0 = e^{2 keys} of this CTA's pairs resident, bit 1 = this CTA's memory columns resident
__host__ __device__ inline Dyn make_dyn(int S, int T_in, int FC, int att_res, int d0, int d1) {
  Dyn d;
  const uint32_t csb = (uint32_t)S * 64u;
  const uint32_t npq = (uint32_t)(T_in * S) / CS + 4;              // (position, sample) pairs per CTA, upper bound
  // every region starts on a 128-byte line: a warp-wide LDS.128 / cp.async of 512 contiguous bytes then touches four
  // lines, not five (measured: +-7 % of the step time depending on where the ring happened to land)
  auto up = [](uint32_t v) { return (v + 127u) & ~127u; };
  d.pq = up(OFF_X + (uint32_t)cum_chunks(DM_NBUF, FC) * csb);      // exp(2 * processed query) fp32 [chunk][n][16]
  d.sc = up(d.pq + 16 * csb);                                      // exp(score - B)  [j][n]
  d.stage = up(d.sc + (uint32_t)T_in * S * 4 + 16u);               // this CTA's pairs
  d.pn = d.stage + ((npq * 4 + 15u) & ~15u);                       // sample index of this CTA's pairs (bytes)
  d.ksl = up(d.pn + npq);                                          // exp(2 * keys) rows of this CTA's pairs
  d.msl = up(d.ksl + ((att_res & 1) ? npq * DH * 4 : 0u));         // memory columns of this CTA, tile order
  d.ring = up(d.msl + ((att_res & 2) ? (uint32_t)S * T_in * 64u : 0u));   // weight ring: d0 KB for each of warps 0-7, d1 KB for 8-15
  d.total = d.ring + (uint32_t)(8 * d0 + 8 * d1) * 1024u;
  return d;
}

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
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"   // suspend-time hint: sleep in hardware
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(mb), "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint4 v, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
               ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rmbar) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// weight stream: read once per step and SM, keep it out of L1
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float rcp_approx(float x) {   // MUFU.RCP: 1/inf = 0, no range fix-ups
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// ---- the streamed weights: a per-warp ring in shared memory, filled by cp.async ------------------------------------------
// Chunk-tiles that fit neither in tensor memory nor in registers are re-read from L2 every step.  Loading them straight
// into registers (the first version) ties their arrival to the register scoreboards: a phase that waits for the chunks
// it requested two phases ago also waits for everything requested since, so the effective prefetch distance was one
// exchange window and the slowest CTA of a cluster ate the L2 latency in every streamed phase.  cp.async groups are
// counted explicitly: the ring is always D chunk-tiles (1 KB = hi + lo fragments of 32 lanes) ahead of the consumer, each
// lane copies and later reads only its own 32 bytes (no cross-lane visibility needed), and `wait_group D - cnt` waits for
// exactly the oldest cnt entries.
struct Ring { uint32_t base, seq; int D, n, rp, kf; };   // base: this lane's 16 bytes of slot 0; seq: smem address of the order table
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(int pend) {   // pend is warp-uniform, 0..7
  switch (pend) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
  }
}
// request the next cnt chunk-tiles of this warp's order into the slots that were consumed last
__device__ __forceinline__ void ring_refill(Ring& r, const uint4* __restrict__ ws, int cnt) {
  int rf = r.rp - cnt;   // the slots consumed last (nothing was taken since)
  if (rf < 0) rf += r.D;
  for (int k = 0; k < cnt; ++k) {
    uint32_t off;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(off) : "r"(r.seq + (uint32_t)r.kf * 4u));
    const uint4* src = ws + (size_t)off * 32;
    const uint32_t dst = r.base + (uint32_t)rf * 1024u;
    cp_async16(dst, src);
    cp_async16(dst + 512u, src + 32);
    cp_async_commit();
    rf = rf + 1 == r.D ? 0 : rf + 1;
    r.kf = r.kf + 1 == r.n ? 0 : r.kf + 1;
  }
}

// ---- tensor memory as a weight store --------------------------------------------------------------------------------
// The 256 KB of TMEM are not needed for accumulators here (the mat-vecs run on mma.sync), so they hold the A fragments of
// the four heaviest phases (both decoder GRUs: 16 chunk-tiles of 1 KB per warp).  A warp reads and writes its own lane
// quarter (warp % 4) with the 32x32b shape: x8 = the eight 32-bit words (hi uint4, lo uint4) of one chunk-tile per lane.
// Measured (tools/ubench/tmem_ld.cu): 40-50 clk load+wait, ~800 B/clk/SM with 16 warps, and no LSU / L2 traffic.
__device__ __forceinline__ void tmem_ld8_if(bool on, uint32_t taddr, uint4& hi, uint4& lo) {   // `on` is warp-uniform
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.u32 p, %9, 0;\n"
      "@!p bra.uni TLD_SKIP;\n"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
      "TLD_SKIP:\n"
      "}\n"
      : "+r"(hi.x), "+r"(hi.y), "+r"(hi.z), "+r"(hi.w), "+r"(lo.x), "+r"(lo.y), "+r"(lo.z), "+r"(lo.w)
      : "r"(taddr), "r"((uint32_t)on));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4& hi, const uint4& lo) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(taddr), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the registers of a slot are only defined after tcgen05.wait::ld: make the compiler see them as produced there
__device__ __forceinline__ void reg_fence(uint4& a, uint4& b) {
  asm volatile("" : "+r"(a.x), "+r"(a.y), "+r"(a.z), "+r"(a.w), "+r"(b.x), "+r"(b.y), "+r"(b.z), "+r"(b.w));
}

// D += A(16x16, row) * B(16x8, col), bf16 inputs, fp32 accumulate
__device__ __forceinline__ void mma16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// One operand pass of a work item: chunks 0..cnt-1 of the activation buffer at xaddr against the weight slots SL.
// The B fragment of chunk i+1 is requested before the MMAs of chunk i (two register sets), so that the three
// accumulator chains run at the tensor pipe's dependent-issue latency instead of LDS + MMA per chunk.
// (I0, I1: the chunk range of this call, for work items whose early part is spread over two windows; the caller
// guarantees I0 < cnt)
template <int I0, int I1, int... SL>
__device__ __forceinline__ void mma_pass(float (&hh)[4], float (&hl)[4], float (&lh)[4], const uint4 (&wb)[NWB], uint32_t xaddr,
                                         uint32_t csb, int cnt) {
  constexpr int sl[] = {SL...};
  constexpr int MAXC = (int)sizeof...(SL) < I1 ? (int)sizeof...(SL) : I1;
  uint4 xa = make_uint4(0u, 0u, 0u, 0u), xb = make_uint4(0u, 0u, 0u, 0u);
  if (I0 & 1) xb = lds128(xaddr + I0 * csb); else xa = lds128(xaddr + I0 * csb);
#pragma unroll
  for (int i = I0; i < MAXC; ++i) {
    if (i + 1 < MAXC && i + 1 < cnt) {
      if (i & 1) xa = lds128(xaddr + (i + 1) * csb); else xb = lds128(xaddr + (i + 1) * csb);
    }
    if (i < cnt) {
      const uint4& xf = (i & 1) ? xb : xa;
      mma16816(hh, wb[2 * sl[i]], xf.x, xf.y);       // W_hi * x_hi
      mma16816(lh, wb[2 * sl[i] + 1], xf.x, xf.y);   // W_lo * x_hi
      mma16816(hl, wb[2 * sl[i]], xf.z, xf.w);       // W_hi * x_lo
    }
  }
}
// fp32 -> bf16 hi (round to nearest) + bf16 lo (remainder); packs two values per 32-bit word
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
  const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0));
  const __nv_bfloat16 l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
  hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
}
// the 16-byte word group lane (n, t) contributes to an activation chunk: columns {2t,2t+1,2t+8,2t+9}
__device__ __forceinline__ uint4 pack_x(float4 v) {
  uint4 r;
  split2(v.x, v.y, r.x, r.z);
  split2(v.z, v.w, r.y, r.w);
  return r;
}

// The weight registers are six slots of one chunk-tile (hi, lo uint4) each.  Every phase names the slots its chunks
// use, chosen so that consecutive phases use disjoint slots wherever they fit: the stream of phase p+1 (or p+2) is
// requested while phase p waits for its exchange, i.e. off the critical path and at least one phase ahead.
// chunks I0, I0+1, ... of a phase (stream offset OFF, in uint4 per lane) into slots SL...
template <int OFF, int I0, int... SL>
__device__ __forceinline__ void load_w(uint4 (&wb)[NWB], const uint4* __restrict__ ws, int cnt, Slots<SL...>) {
  constexpr int sl[] = {SL...};
#pragma unroll
  for (int k = 0; k < (int)sizeof...(SL); ++k)
    if (I0 + k < cnt) {
      wb[2 * sl[k]] = ldg_stream(ws + (OFF + 2 * (I0 + k)) * 32);
      wb[2 * sl[k] + 1] = ldg_stream(ws + (OFF + 2 * (I0 + k) + 1) * 32);
    }
}

// the same from tensor memory: chunk-tile I0+k of a phase lives in column group TC0+I0+k of the warp's TMEM slice
template <int TC0, int I0, int... SL>
__device__ __forceinline__ void load_t(uint4 (&wb)[NWB], uint32_t tw, int cnt, Slots<SL...>) {
  constexpr int sl[] = {SL...};
#pragma unroll
  for (int k = 0; k < (int)sizeof...(SL); ++k)
    tmem_ld8_if(I0 + k < cnt, tw + (uint32_t)(TC0 + I0 + k) * 8u, wb[2 * sl[k]], wb[2 * sl[k] + 1]);
}
// chunk-tiles 0..cnt-1 at consecutive column groups from tx (runtime base: the spare columns of another warp's slice)
template <int... SL>
__device__ __forceinline__ void load_tx(uint4 (&wb)[NWB], uint32_t tx, int cnt, Slots<SL...>) {
  constexpr int sl[] = {SL...};
#pragma unroll
  for (int k = 0; k < (int)sizeof...(SL); ++k) tmem_ld8_if(k < cnt, tx + (uint32_t)k * 8u, wb[2 * sl[k]], wb[2 * sl[k] + 1]);
}
// the oldest cnt chunk-tiles of the ring into the register slots SL
template <int... SL>
__device__ __forceinline__ void ring_take(uint4 (&wb)[NWB], Ring& r, int cnt, Slots<SL...>) {
  constexpr int sl[] = {SL...};
  if (cnt == 0) return;
  cp_async_wait_pending(r.D - cnt);
#pragma unroll
  for (int k = 0; k < (int)sizeof...(SL); ++k)
    if (k < cnt) {
      const uint32_t src = r.base + (uint32_t)r.rp * 1024u;
      wb[2 * sl[k]] = lds128(src);
      wb[2 * sl[k] + 1] = lds128(src + 512u);
      r.rp = r.rp + 1 == r.D ? 0 : r.rp + 1;
    }
}
template <int... SL>
__device__ __forceinline__ void wait_t(uint4 (&wb)[NWB], Slots<SL...>) {
  constexpr int sl[] = {SL...};
  tmem_wait_ld();
#pragma unroll
  for (int k = 0; k < (int)sizeof...(SL); ++k) reg_fence(wb[2 * sl[k]], wb[2 * sl[k] + 1]);
}

// one warp's share of a phase: cnt chunks of one tile (chunk i in slot SL[i]); partial tile -> red slot of this warp
// NX > 0: the same weight chunks also multiply the chunks at byte offsets d1 (and d2) from xaddr when nx >= 1 (2) --
// a sum of activation vectors (y1 = y0 + h1', y2 = y0 + h1' + h2') is never formed or exchanged, only its terms are.
template <int NX, int... SL>
__device__ __forceinline__ void mma_chunks(const uint4 (&wb)[NWB], uint32_t xaddr, uint32_t csb, int cnt, uint32_t slot,
                                           int g, int t, Slots<SL...>, int nx = 0, uint32_t d1 = 0, uint32_t d2 = 0) {
  constexpr int I0 = 0, I1 = 8;
  float hh[4] = {0.f, 0.f, 0.f, 0.f}, hl[4] = {0.f, 0.f, 0.f, 0.f}, lh[4] = {0.f, 0.f, 0.f, 0.f};
  // one branch per operand pass (a block of this size is not if-converted), per-chunk predicates inside
  if (cnt > 0) {
    mma_pass<I0, I1, SL...>(hh, hl, lh, wb, xaddr, csb, cnt);
  }
  if (NX >= 1 && cnt > 0 && nx >= 1) {
    mma_pass<I0, I1, SL...>(hh, hl, lh, wb, xaddr + d1, csb, cnt);
  }
  if (NX >= 2 && cnt > 0 && nx >= 2) {
    mma_pass<I0, I1, SL...>(hh, hl, lh, wb, xaddr + d2, csb, cnt);
  }
  // D[row g / g+8 = tile column][col 2t, 2t+1 = sample]  ->  slot[sample][column]   (bank-conflict free with RS = 20)
  // (written even when cnt == 0 so that the reducer never sums a stale slot)
  const uint32_t p = slot + (uint32_t)(((2 * t) * RS + g) * 4);
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(p), "f"(hh[0] + (hl[0] + lh[0])) : "memory");
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(p + RS * 4), "f"(hh[1] + (hl[1] + lh[1])) : "memory");
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(p + 32), "f"(hh[2] + (hl[2] + lh[2])) : "memory");
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(p + RS * 4 + 32), "f"(hh[3] + (hl[3] + lh[3])) : "memory");
}

// The same work item in two parts.  PART 0 (early) runs in the exchange window BEFORE the phase's mbarrier wait and
// multiplies every operand that is already complete (the previous recurrent state, the context of the previous step,
// y0 / h1' for the sums y1, y2); its partial tile stays in four registers.  PART 1 (late) runs after the wait, adds the
// operand that has just arrived and writes the partial tile.  Work item flags: nx (extra buffers), early (bit 30).
//   nx == 0: the one buffer is early iff flagged;  nx >= 1: every buffer but the last extra one is early.
// I0, I1: chunk range of this call; ACC: the early part continues one started in an earlier window.
template <int PART, int NX, int I0, int I1, bool ACC, int... SL>
__device__ __forceinline__ void mma_split(float (&pre)[4], const uint4 (&wb)[NWB], uint32_t xl, uint32_t csb, uint32_t e,
                                          uint32_t slot, int g, int t, Slots<SL...>, uint32_t d1 = 0, uint32_t d2 = 0) {
  const int cnt = (int)(e & 7u), nx = (int)((e >> 28) & 3u);
  const bool early = ((e >> 30) & 1u) != 0;
  const uint32_t xaddr = xl + ((e & 0x0fffffffu) >> 3);
  const bool do0 = PART == 0 ? (nx != 0 || early) : (nx == 0 && !early);
  const bool do1 = NX >= 1 && (PART == 0 ? nx >= 2 : nx == 1);
  const bool do2 = NX >= 2 && PART == 1 && nx == 2;
  float hh[4], hl[4] = {0.f, 0.f, 0.f, 0.f}, lh[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 4; ++k) hh[k] = (PART == 0 && !ACC) ? 0.f : pre[k];
  // one branch per operand pass (a block of this size is not if-converted), per-chunk predicates inside
  if (do0 && cnt > I0) {
    mma_pass<I0, I1, SL...>(hh, hl, lh, wb, xaddr, csb, cnt);
  }
  if (NX >= 1 && do1 && cnt > I0) {
    mma_pass<I0, I1, SL...>(hh, hl, lh, wb, xaddr + d1, csb, cnt);
  }
  if (NX >= 2 && do2 && cnt > I0) {
    mma_pass<I0, I1, SL...>(hh, hl, lh, wb, xaddr + d2, csb, cnt);
  }
  if (PART == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) pre[k] = hh[k] + (hl[k] + lh[k]);
  } else {
    const uint32_t p = slot + (uint32_t)(((2 * t) * RS + g) * 4);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(p), "f"(hh[0] + (hl[0] + lh[0])) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(p + RS * 4), "f"(hh[1] + (hl[1] + lh[1])) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(p + 32), "f"(hh[2] + (hl[2] + lh[2])) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(p + RS * 4 + 32), "f"(hh[3] + (hl[3] + lh[3])) : "memory");
  }
}

// reducer thread (n, c): sum over the partial tiles in slots [s0, s0+NS) of element [n][c]
template <int NS>
__device__ __forceinline__ float red_sum(uint32_t red_nc, int s0) {
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int s = 0; s < NS; s += 2) {
    a += lds_f(red_nc + (s0 + s) * SLOT * 4);
    if (s + 1 < NS) b += lds_f(red_nc + (s0 + s + 1) * SLOT * 4);
  }
  return a + b;
}
// reducer thread (n, c) holds v = activation [n][16q + c]: split into bf16 hi/lo, pair with the neighbouring column
// (lane ^ 1) and write the two words of the pair into the staged block row `stg_n` (MMA-fragment order).
// Must be executed by whole warps.
__device__ __forceinline__ void stage_x(uint32_t stg_n, int c, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  const uint32_t hv = __bfloat16_as_ushort(h), lv = __bfloat16_as_ushort(l);
  const uint32_t hn = __shfl_xor_sync(0xffffffffu, hv, 1), ln = __shfl_xor_sync(0xffffffffu, lv, 1);
  if ((c & 1) == 0) {
    const uint32_t wa = stg_n + (uint32_t)((((c & 7) >> 1) * 4 + (c >> 3)) * 4);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(wa), "r"(hv | (hn << 16)) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(wa + 8), "r"(lv | (ln << 16)) : "memory");
  }
}

// bias table (floats, tile order so that lane t reads one float4 at [t*4])
enum { BI_P1 = 0, BI_P2 = 16, BI_RA = 32, BI_UA = 48, BI_CA = 64, BI_PC = 80, BI_R1 = 96, BI_U1 = 112, BI_C1 = 128,
       BI_R2 = 144, BI_U2 = 160, BI_C2 = 176, BI_OA = 192, BI_OB = 208 };
static_assert(BI_OB + 16 == DM_NBIAS, "bias table size");

#define TRM(i) do { if (TRACE && a.trace != nullptr && step == 8 && blockIdx.x == a.trace_cta && tid == 0) a.trace[i] = clock64(); } while (0)
// per-warp stamp: 16 consecutive entries starting at `base`
#define TRW(base) do { if (TRACE && a.trace != nullptr && step == 8 && blockIdx.x == a.trace_cta && lane == 0) a.trace[(base) + warp] = clock64(); } while (0)

// TRACE: clock stamps of one CTA (developer aid).  A separate instantiation: even predicated off, the stamp
// instructions (LDC of the pointer, CS2R, STG) took ~10 % of the stall samples of the production kernel.
template <bool TRACE>
__global__ void __launch_bounds__(NT, 1)
decoder_mma_kernel(const DecoderMmaWeights w, const DecoderArgs a, const int nclusters) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // the warp index through a shuffle from lane 0: ptxas then knows it is warp-uniform (uniform registers and branches
  // instead of predication, BSSY/BSYNC and WARPSYNC around every per-warp decision)
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int g = lane >> 2, t = lane & 3;       // MMA fragment coordinates; as reducer: sample g, column group t
  const int q = (int)cluster_ctarank();
  const int cid = (int)cluster_id_x();
  // balanced cut of the batch over the clusters
  const int base = a.N / nclusters, rem = a.N % nclusters;
  const int S = base + (cid < rem ? 1 : 0);
  const int n0 = cid * base + min(cid, rem);
  const int M = w.M, Dout = w.Dout, FC = M >> 4, T_in = a.T_in;
  const bool res_k = (a.att_res & 1) != 0, res_m = (a.att_res & 2) != 0;
  const Dyn L = make_dyn(S, T_in, FC, a.att_res, a.ring_d0, a.ring_d1);
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t csb = (uint32_t)S * 64u;
  const uint32_t mb0 = sbase + OFF_MBAR;
  // attention scores are cut by flattened (position j, sample n) pair index p = j*S + n: 16 balanced ranges
  // (cut at multiples of four pairs: a CTA's exp(score) block goes out as 16-byte words)
  const int NP = T_in * S, NQ = (NP + 3) >> 2;
  const int p0 = 4 * ((q * NQ) / CS), npq = min(NP, 4 * (((q + 1) * NQ) / CS)) - p0;
  if (S == 0) {   // cannot happen (nclusters <= N); keeps every CTA of a cluster on the same path
    cluster_sync_all();
    cluster_sync_all();
    return;
  }

  // ---- prologue -------------------------------------------------------------------------
  for (uint32_t i = OFF_WTAB + tid * 4; i < L.total; i += NT * 4) *reinterpret_cast<uint32_t*>(smem_raw + i) = 0u;
  if (tid < NBAR) mbar_init(mb0 + tid * 8, 1);
  __syncthreads();
  if (tid < DM_NPHASE * 16) {   // per-warp work table: byte offset of the warp's first chunk and its chunk count
    const uint32_t e = w.tab[tid >> 4][tid & 15];
    const uint32_t bufi = (e >> 8) & 15u, c0 = (e >> 3) & 31u;
    reinterpret_cast<uint32_t*>(smem_raw + OFF_WTAB)[tid] =
        (((e >> 16) & 1u) << 30) | (((e >> 14) & 3u) << 28) | ((OFF_X + ((uint32_t)cum_chunks((int)bufi, FC) + c0) * csb) << 3) | (e & 7u);
  }
  if (tid < DM_NBIAS) reinterpret_cast<float*>(smem_raw + OFF_BIAS)[tid] = __ldg(w.bias + q * DM_NBIAS + tid);
  if (tid < DH) reinterpret_cast<float*>(smem_raw + OFF_VATT)[tid] = __ldg(w.att_v + tid);
  if (res_k) {
    float* ksl = reinterpret_cast<float*>(smem_raw + L.ksl);
    for (int i = tid; i < npq * (DH / 4); i += NT) {   // exp(2 key): tanh(k + p) = 1 - 2 / (1 + e^{2k} e^{2p})
      const int c4 = i % (DH / 4), pp = i / (DH / 4), p = p0 + pp, j = p / S, n = p - j * S;   // (prologue only)
      float4 k4 = ldg_f4(a.keys + ((size_t)(n0 + n) * T_in + j) * DH + c4 * 4);
      k4.x = __expf(2.0f * fminf(fmaxf(k4.x, -30.f), 30.f)); k4.y = __expf(2.0f * fminf(fmaxf(k4.y, -30.f), 30.f));
      k4.z = __expf(2.0f * fminf(fmaxf(k4.z, -30.f), 30.f)); k4.w = __expf(2.0f * fminf(fmaxf(k4.w, -30.f), 30.f));
      *reinterpret_cast<float4*>(ksl + (size_t)pp * DH + c4 * 4) = k4;
    }
  }
  if (res_m) {
    float* msl = reinterpret_cast<float*>(smem_raw + L.msl);
    for (int i = tid; i < S * T_in * 4; i += NT) {   // 16 columns = four float4 per (sample, position)
      const int c4 = i & 3, r = i >> 2, s = r / T_in, j = r - s * T_in;
      *reinterpret_cast<float4*>(msl + ((size_t)j * S + s) * 16 + c4 * 4) =
          ldg_f4(a.memory + ((size_t)(n0 + s) * T_in + j) * DH + q * 16 + c4 * 4);   // [j][n][16]
    }
  }
  float vbound = 0.f, vsum = 0.f;   // B = min(||v||_1, 40) >= any score (|tanh| <= 1); sum_k v_k
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float vv = __ldg(w.att_v + lane + 32 * i); vbound += fabsf(vv); vsum += vv; }
  vbound = fminf(warp_sum(vbound), 40.0f);
  vsum = warp_sum(vsum);
  if (tid == 0) sts_f(sbase + OFF_VB, vsum - vbound);   // written before the barriers of the prologue

  const uint4* ws = reinterpret_cast<const uint4*>(w.stream) + ((size_t)(q * NW + warp) * F4_STEP) * 32 + lane;
  const uint32_t wtab = sbase + OFF_WTAB + warp * 4;                       // + phase * 64
  const uint32_t xl = sbase + (uint32_t)min(g, S - 1) * 64u + (uint32_t)t * 16u;   // lane part of a B-fragment address
  const uint32_t myslot = sbase + OFF_RED + warp * SLOT * 4;
  // reducer group: warps 0..3, thread (rn, rc) = (sample, column of this CTA's tile)
  const int rn = (tid >> 4) & 7, rc = tid & 15;
  // Reducer groups sit on the warps with the least MMA / weight-ring work (the work table loads warps 0-7 most):
  // group 0 (sums, gate math, staging, the exchange) = warps 12-15, group 1 = warps 8-11, group 2 = warps 4-7.
  const bool red_grp = warp >= 12, red_grp1 = (warp >> 2) == 2, red_grp2 = (warp >> 2) == 1;
  const uint32_t red_nc = sbase + OFF_RED + (rn * RS + rc) * 4;
  const uint32_t st_nc = sbase + OFF_STATE + (rn * 16 + rc) * 4;           // + slot * 512
  const uint32_t bias_c = sbase + OFF_BIAS + rc * 4;                       // + table * 4
  const uint32_t stg_n = sbase + OFF_STG + rn * 64;                        // + block * 512
  const uint32_t rmb0 = mapa_u32(mb0, (uint32_t)warp);                      // peer `warp`: its mbarriers ...
  const uint32_t rx = mapa_u32(sbase + lane * 16, (uint32_t)warp);          // ... and this lane's slot in a chunk at offset 0
  // A reducer warp sends what it staged itself: warp r of a reducer group owns the rows of samples 2r, 2r+1 of staged
  // block `blk` (8 word groups of 16 bytes) and pushes them to all 16 peers (or, `pair`, to the 8 peers of this CTA's
  // parity) -- no block barrier between the reduction and the exchange, and the other 12 warps are already in the window.
  auto send_rows = [&](int blk, uint32_t dst, int bar, bool pair) {
    __syncwarp();
    const int wg = lane & 7, sn = 2 * (warp & 3) + (wg >> 2);
    if (sn < S) {
      const uint32_t off = (uint32_t)sn * 64u + (uint32_t)(wg & 3) * 16u;
      const uint4 v = lds128(sbase + OFF_STG + (uint32_t)blk * 512u + off);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        if (pair && it >= 2) break;
        const uint32_t k = (uint32_t)(it * 4 + (lane >> 3));
        const uint32_t peer = pair ? 2u * k + (uint32_t)(q & 1) : k;
        st_async_v4(mapa_u32(sbase + dst + off, peer), v, mapa_u32(mb0 + (uint32_t)bar * 8u, peer));
      }
    }
  };
#define WCNT(ph) ((uint32_t)__shfl_sync(0xffffffffu, lds32(wtab + (ph) * 64), 0))
#define XBUF(b) (OFF_X + (uint32_t)cum_chunks((b), FC) * csb)

  // P6: sample index of every local pair (one integer division each, once)
  for (int pp = tid; pp < npq; pp += NT) smem_raw[L.pn + pp] = (unsigned char)((p0 + pp) % S);
  uint4 wb[NWB];
  if (warp == 0) {   // all 512 TMEM columns: the kernel owns the SM anyway (512 threads x 128 registers)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sbase + OFF_TMEM) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();   // work table and TMEM base visible
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = lds32(sbase + OFF_TMEM);
  // this warp's slice: lane quarter warp % 4 (hardware rule), 128 columns = 16 chunk-tiles: P9 0-5, P10 6-7, P11 8-13, P12 14-15
  const uint32_t tw = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 128u;
  {
    auto fill = [&](int tp, int off, int tc0) {
      const int cnt = (int)(WCNT(tp) & 7u);
      for (int i0 = 0; i0 < cnt; i0 += 3) {   // three chunk-tiles in flight: the L2 latency is paid twice, not six times
        uint4 hi[3], lo[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (i0 + k < cnt) { hi[k] = ldg_stream(ws + (off + 2 * (i0 + k)) * 32); lo[k] = ldg_stream(ws + (off + 2 * (i0 + k) + 1) * 32); }
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (i0 + k < cnt) tmem_st8(tw + (uint32_t)(tc0 + i0 + k) * 8u, hi[k], lo[k]);
      }
    };
    fill(T_P9, O9, TC9); fill(T_P10, O10, TC10); fill(T_P11, O11, TC11); fill(T_P12, O12, TC12);
  }
  // Warps 8..15 leave 4-8 of their 16 chunk-tiles unused, and the four warps of a lane quarter can read each other's
  // columns.  The spare ones hold the operands of the EARLY parts that would otherwise come from L2 one phase sooner
  // than the slot schedule can request them (measured: up to 1 k clk of L2-latency spread on the slowest CTA):
  //   the h_att rows of the attention GRU's gates (P3, warps 2-5 and 8-11, 4 chunk-tiles each)
  //       -> slice of warp 12 + (warp % 4), chunk-tiles 4-7 (first h-warp of the lane quarter) or 12-15 (second);
  //   the context rows of the prenet (P1, warps 0-7, 2 chunk-tiles each)
  //       -> slice of warp 8 + (warp % 4), chunk-tiles 6-7 (warps 0-3) or 14-15 (warps 4-7).
  const bool early1 = ((WCNT(T_P1) >> 30) & 1u) != 0, early3 = ((WCNT(T_P3) >> 30) & 1u) != 0;
  #define TX1 (tw - (uint32_t)(warp >> 2) * 128u + 2u * 128u + (warp < 4 ? 6u : 14u) * 8u)
#define TX3 (tw - (uint32_t)(warp >> 2) * 128u + 3u * 128u + (warp < 8 ? 4u : 12u) * 8u)
  {
    auto fillx = [&](int tp, int off, uint32_t tx) {
      const int cnt = (int)(WCNT(tp) & 7u);
      for (int i = 0; i < cnt; ++i) {
        const uint4 hi = ldg_stream(ws + (off + 2 * i) * 32), lo = ldg_stream(ws + (off + 2 * i + 1) * 32);
        tmem_st8(tx + (uint32_t)i * 8u, hi, lo);
      }
    };
    if (early1) fillx(T_P1, O1, TX1);
    if (early3) fillx(T_P3, O3, TX3);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  // the streamed chunk-tiles of this warp in consumption order (stream offsets in 512-byte rows)
  Ring rg;
  rg.seq = sbase + OFF_SEQ + (uint32_t)warp * 64u;
  rg.n = 0;
  {
    auto add = [&](int tp, int off) {
      const int cnt = (int)(WCNT(tp) & 7u);
      for (int i = 0; i < cnt; ++i) {
        if (lane == 0) asm volatile("st.shared.b32 [%0], %1;" ::"r"(rg.seq + (uint32_t)rg.n * 4u), "r"((uint32_t)(off + 2 * i)) : "memory");
        ++rg.n;
      }
    };
    if (!early1) add(T_P1, O1);
    add(T_P2, O2);
    if (!early3) add(T_P3, O3);
    add(T_P4, O4); add(T_P5, O5); add(T_P8, O8); add(T_P13, O13);
  }
  __syncwarp();
  {
    const int dmax = warp < 8 ? a.ring_d0 : a.ring_d1;
    rg.D = rg.n < dmax ? rg.n : dmax;
    rg.base = sbase + L.ring + (uint32_t)(warp < 8 ? warp * a.ring_d0 : 8 * a.ring_d0 + (warp - 8) * a.ring_d1) * 1024u + (uint32_t)lane * 16u;
    rg.rp = 0; rg.kf = 0;
    ring_refill(rg, ws, rg.D);   // rp - D = rp (mod D): fills every slot once, the ring is full
  }
  if (!early1) ring_take(wb, rg, (int)(WCNT(T_P1) & 7u), SL1());
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  cluster_sync_all();   // buffers zeroed and mbarriers initialised everywhere before anyone pushes

  const bool free_run = a.targets == nullptr;
  const uint32_t BLK = CS * csb;
  const uint32_t dH1 = (uint32_t)(cum_chunks(DM_BH1, FC) - cum_chunks(DM_BY0, FC)) * csb;   // y0 chunk i -> h1' chunk i
  const uint32_t dH2 = (uint32_t)(cum_chunks(DM_BH2, FC) - cum_chunks(DM_BY0, FC)) * csb;   // y0 chunk i -> h2' chunk i
  const int fb_tile0 = (Dout - M) >> 4, ntiles = Dout >> 4;

#define RTAKE(TP, SL) ring_take(wb, rg, (int)(WCNT(TP) & 7u), SL());
#define RFILL(TP) ring_refill(rg, ws, (int)(WCNT(TP) & 7u));
#define TLOADP(TP, TC0, I0, ...) load_t<TC0, I0>(wb, tw, (int)(WCNT(TP) & 7u), Slots<__VA_ARGS__>());
#define TWAIT(SL) wait_t(wb, SL());
#define MMA(SL, TP) { const uint32_t e = WCNT(TP); mma_chunks<0>(wb, xl + ((e & 0x0fffffffu) >> 3), csb, e & 7, myslot, g, t, SL()); }
#define MMA_PRE(NX, SL, TP) mma_split<0, NX, 0, 8, false>(pre, wb, xl, csb, WCNT(TP), myslot, g, t, SL(), dH1, dH2);
#define MMA_PRE_R(NX, SL, TP, I0, I1, ACC) mma_split<0, NX, I0, I1, ACC>(pre, wb, xl, csb, WCNT(TP), myslot, g, t, SL(), dH1, dH2);
#define MMA_POST(NX, SL, TP) mma_split<1, NX, 0, 8, false>(pre, wb, xl, csb, WCNT(TP), myslot, g, t, SL(), dH1, dH2);
// Window work of the decoder GRUs: the early part of P11 is spread over the windows of P9 (chunks 0-3) and P10 (4-5),
// the early part of P13 runs in the window of P12 (same-box A/B, tools/ab: each placement was worth 0.5-1 %).
#define W9_WORK TLOADP(T_P10, TC10, 0, 0, 1) TLOADP(T_P11, TC11, 0, 2, 3, 4, 5) TWAIT(SL11) MMA_PRE_R(1, SL11, T_P11, 0, 4, false)
#define W10_WORK TLOADP(T_P11, TC11, 4, 0, 1) TWAIT(SL11) MMA_PRE_R(1, SL11, T_P11, 4, 6, true)
#define W11_WORK TLOADP(T_P12, TC12, 0, 0, 1)
#define W12_WORK RTAKE(T_P13, SL13) MMA_PRE(2, SL13, T_P13)
#define ST(slot) (st_nc + (slot) * 512)
#define BIAS(tab) lds_f(bias_c + (tab) * 4)
  float pre[4] = {0.f, 0.f, 0.f, 0.f};   // early part of the next phase's partial tile (step 0, P1: the context is zero)
  int trb = -1;   // developer aid: base of the per-warp stamps inside the GRU macros
  for (int step = 0; step < a.steps; ++step) {
    const uint32_t par = (uint32_t)step & 1u;
    if (free_run && step > 0) mbar_wait(mb0 + B_P13 * 8, par ^ 1u);   // fed-back frame of step-1 has landed
    if (tid == 0) {   // this step's expected byte counts (remote credits may already have arrived)
      mbar_expect_tx(mb0 + B_P1 * 8, BLK);
      mbar_expect_tx(mb0 + B_P2 * 8, BLK / 2);
      mbar_expect_tx(mb0 + B_P3 * 8, BLK);
      mbar_expect_tx(mb0 + B_P4 * 8, BLK);
      mbar_expect_tx(mb0 + B_P5 * 8, BLK);
      mbar_expect_tx(mb0 + B_P6 * 8, (uint32_t)NQ * 16u);
      mbar_expect_tx(mb0 + B_P7 * 8, BLK);
      mbar_expect_tx(mb0 + B_P8 * 8, BLK);
      mbar_expect_tx(mb0 + B_P9 * 8, BLK);
      mbar_expect_tx(mb0 + B_P10 * 8, BLK);
      mbar_expect_tx(mb0 + B_P11 * 8, BLK);
      mbar_expect_tx(mb0 + B_P12 * 8, BLK);
      if (free_run) mbar_expect_tx(mb0 + B_P13 * 8, (uint32_t)FC * csb);
    }
    if (!free_run) {   // teacher forcing: input = mel_targets[:, (step-1)*r + r-1, :]  (helpers.py:48,75)
      if (step > 0 && tid < S * FC * 4) {
        const int tt = tid & 3, r2 = tid >> 2, n = r2 % S, ch = r2 / S;
        const float* src = a.targets + ((size_t)(n0 + n) * a.T_tgt + (size_t)(step - 1) * a.r + a.r - 1) * M + ch * 16;
        const float2 lo2 = __ldg(reinterpret_cast<const float2*>(src + 2 * tt));
        const float2 hi2 = __ldg(reinterpret_cast<const float2*>(src + 2 * tt + 8));
        const uint4 pk = pack_x(make_float4(lo2.x, lo2.y, hi2.x, hi2.y));
        sts128(sbase + XBUF(DM_BF) + ch * csb + n * 64 + tt * 16, pk.x, pk.y, pk.z, pk.w);
      }
      __syncthreads();
    }
    TRM(0);
    // ================= P1: decoder prenet dense_1 + ReLU on [frame | context] =================
    MMA_POST(0, SL1, T_P1)                      // frame chunks; the context chunks were multiplied in the window of P13
    __syncthreads();
    TRM(1);
    if (red_grp) {
      stage_x(stg_n, rc, fmaxf(red_sum<DM_P1_SLOTS>(red_nc, 0) + BIAS(BI_P1), 0.f));
      send_rows(0, XBUF(DM_BP1) + q * csb, B_P1, false);
    }
    if (!early1) { RFILL(T_P1) }                // window of P1: refill the ring, P2's chunks into registers
    RTAKE(T_P2, SL2)
    TRM(2);
    mbar_wait(mb0 + B_P1 * 8, par);
    TRM(3);
    // ================= P2: prenet dense_2 + ReLU (CTA pair 2c, 2c+1 computes chunk c; each feeds half the peers) ====
    MMA(SL2, T_P2)
    __syncthreads();
    TRM(4);
    if (red_grp) {
      stage_x(stg_n, rc, fmaxf(red_sum<8>(red_nc, 0) + BIAS(BI_P2), 0.f));
      send_rows(0, XBUF(DM_BP2) + (q >> 1) * csb, B_P2, true);
    }
    RFILL(T_P2)                                 // window of P2: the x rows of P3 from the ring ...
    if (!early3) { RTAKE(T_P3, SL3) }
    if (early3) {                               // ... and the h_att rows of the attention GRU's gates, from TMEM
      load_tx(wb, TX3, (int)(WCNT(T_P3) & 7u), SL3());
      TWAIT(SL3)
      MMA_PRE(0, SL3, T_P3)
    } else { pre[0] = pre[1] = pre[2] = pre[3] = 0.f; }
    TRM(5);
    mbar_wait(mb0 + B_P2 * 8, par);
    TRM(6);
    // ================= GRU phases: gates r,u on [x | h] + candidate x-part; then candidate h-part =================
#define GRU_GATES(NXG, SLG, TP, BI_R, BI_U, ST_H, BUF_R, BAR, URGENT, LATER)                             \
    MMA_POST(NXG, SLG, TP)                                                                              \
    URGENT   /* weights of the NEXT phase go into the slots this phase used last */                  \
    if (TRACE && trb >= 0) TRW(trb);                                                                          \
    __syncthreads();                                                                                 \
    /* three reducer groups work side by side: warps 0-3 r (-> r*h staged), 4-7 u, 8-11 the candidate's x part */ \
    if (red_grp) {                                                                                   \
      const float r = sigmoid_f(red_sum<6>(red_nc, 0) + BIAS(BI_R));                                 \
      stage_x(stg_n, rc, r * lds_f(ST(ST_H)));                                                       \
      /* the exchange must not complete before groups 1, 2 have read their partial tiles (the next phase overwrites them) */ \
      asm volatile("bar.sync 1, 384;" ::: "memory");                                                 \
      send_rows(0, XBUF(BUF_R) + q * csb, BAR, false);                                               \
    } else if (red_grp1) {                                                                           \
      sts_f(ST(ST_U), sigmoid_f(red_sum<6>(red_nc, 6) + BIAS(BI_U)));                                \
      asm volatile("bar.arrive 1, 384;" ::: "memory");                                               \
    } else if (red_grp2) {                                                                           \
      sts_f(ST(ST_CX), red_sum<4>(red_nc, 12));                                                      \
      asm volatile("bar.arrive 1, 384;" ::: "memory");                                               \
    }                                                                                                \
    if (TRACE && trb >= 0) TRW(trb + 16);                                                                     \
    LATER
    // candidate: h' = u h + (1-u) tanh(c_h + c_x + b); y_out = y_in + h' (ResidualWrapper) when BUF_Y >= 0
#define GRU_CAND(SLC, TP, BI_C, ST_H, ST_YIN, ST_YOUT, BUF_H, BUF_Y, BAR, URGENT, LATER)                     \
    MMA(SLC, TP)                                                                                      \
    URGENT                                                                                           \
    if (TRACE && trb >= 0) TRW(trb);                                                                          \
    __syncthreads();                                                                                 \
    if (red_grp) {                                                                                   \
      const float c = tanh_f(red_sum<8>(red_nc, 0) + lds_f(ST(ST_CX)) + BIAS(BI_C));                 \
      const float u = lds_f(ST(ST_U));                                                               \
      const float h = u * lds_f(ST(ST_H)) + (1.0f - u) * c;                                          \
      sts_f(ST(ST_H), h);                                                                            \
      stage_x(stg_n, rc, h);                                                                         \
      if (BUF_Y >= 0) {                                                                              \
        const float y = lds_f(ST(ST_YIN)) + h;                                                       \
        if (ST_YOUT >= 0) sts_f(ST(ST_YOUT < 0 ? 0 : ST_YOUT), y);                                   \
        stage_x(stg_n + 512, rc, y);                                                                 \
      }                                                                                              \
      send_rows(0, XBUF(BUF_H) + q * csb, BAR, false);                                               \
      if (BUF_Y >= 0) send_rows(1, XBUF(BUF_Y < 0 ? 0 : BUF_Y) + q * csb, BAR, false);               \
    }                                                                                                \
    if (TRACE && trb >= 0) TRW(trb + 16);                                                                     \
    LATER

    // ----- P3 / P4: attention GRU on [prenet | h_att] -----
    TRW(192);
    trb = 208;
    GRU_GATES(0, SL3, T_P3, BI_RA, BI_UA, ST_HA, DM_BRA, B_P3, , if (!early3) { RFILL(T_P3) } RTAKE(T_P4, SL4))
    trb = -1;
    TRM(8);
    mbar_wait(mb0 + B_P3 * 8, par);
    TRM(9);
    GRU_CAND(SL4, T_P4, BI_CA, ST_HA, ST_HA, -1, DM_BHA, -1, B_P4, , RFILL(T_P4) RTAKE(T_P5, SL5))
    TRM(11);
    mbar_wait(mb0 + B_P4 * 8, par);
    TRM(12);
    // ================= P5: query layer (tile A) and h_att' half of the 512->256 projection (tile B) =================
    MMA(SL5, T_P5)
    __syncthreads();
    TRM(13);
    if (red_grp) {   // the query is pushed as e^{2 pq} (fp32) for the score phase
      sts_f(stg_n + rc * 4, __expf(2.0f * fminf(fmaxf(red_sum<8>(red_nc, 0), -30.f), 30.f)));
      send_rows(0, L.pq + q * csb, B_P5, false);
    } else if (red_grp1) {
      sts_f(ST(ST_Y0H), red_sum<8>(red_nc, 8));
    }
    RFILL(T_P5)                           // window of P5: refill
    TRM(14);
    mbar_wait(mb0 + B_P5 * 8, par);
    TRM(15);
    TRW(256);
    // ================= P6: Bahdanau scores of this CTA's (position, sample) pairs: exp(v . tanh(keys + pq) - B) ======
    // v.tanh(k + p) = sum v - 2 sum_k v_k / (1 + e^{2k} e^{2p}); e^{2k} is resident, e^{2p} was pushed by P5.
    // Warp w takes the local pairs 2w, 2w+1 (+32, ...): two per pass, one butterfly reduces both.
    // Four elements share one MUFU.RCP: v1/a + v2/b + v3/c + v4/d = (N_ab cd + N_cd ab) / (ab cd) with N_ab = v1 b + v2 a and
    // the denominators clamped to 2^30 (tanh = 1 - 2^-29 there: below fp32 resolution of the score).
    {
      // lane's 8 inputs: k = 4 lane + {0..3} (chunk lane>>2, columns 4 (lane&3)..) and 128 + the same (chunk + 8)
      const uint32_t pq_l = sbase + L.pq + (uint32_t)(lane >> 2) * csb + (uint32_t)(lane & 3) * 16u;
      const float4 v0 = lds_f4(sbase + OFF_VATT + lane * 16), v1 = lds_f4(sbase + OFF_VATT + 512 + lane * 16);
      constexpr float BIG = 1073741824.0f;   // 2^30
      auto quad = [&](const float4& k, const float4& e, const float4& v) -> float {
        const float da = fminf(fmaf(k.x, e.x, 1.0f), BIG), db = fminf(fmaf(k.y, e.y, 1.0f), BIG);
        const float dc = fminf(fmaf(k.z, e.z, 1.0f), BIG), dd = fminf(fmaf(k.w, e.w, 1.0f), BIG);
        const float ab = da * db, cd = dc * dd;
        const float nab = fmaf(v.x, db, v.y * da), ncd = fmaf(v.z, dd, v.w * dc);
        return fmaf(nab, cd, ncd * ab) * rcp_approx(ab * cd);
      };
      auto pair_sum = [&](int pl) -> float {   // sum_k v_k / (1 + e^{2k} e^{2p}) over this lane's 8 inputs of local pair pl
        const int n = smem_raw[L.pn + pl];
        const float4 e0 = lds_f4(pq_l + (uint32_t)n * 64u), e1 = lds_f4(pq_l + (uint32_t)n * 64u + 8 * csb);   // e^{2 pq}
        float4 k0, k1;                                                                                     // e^{2 key}
        if (res_k) {
          k0 = lds_f4(sbase + L.ksl + (uint32_t)pl * (DH * 4) + lane * 16);
          k1 = lds_f4(sbase + L.ksl + (uint32_t)pl * (DH * 4) + 512 + lane * 16);
        } else {
          const int j = (p0 + pl) / S;
          const float* krow = a.keys + ((size_t)(n0 + n) * T_in + j) * DH + 4 * lane;
          k0 = ldg_f4(krow); k1 = ldg_f4(krow + 128);
          k0.x = __expf(2.0f * fminf(fmaxf(k0.x, -30.f), 30.f)); k0.y = __expf(2.0f * fminf(fmaxf(k0.y, -30.f), 30.f));
          k0.z = __expf(2.0f * fminf(fmaxf(k0.z, -30.f), 30.f)); k0.w = __expf(2.0f * fminf(fmaxf(k0.w, -30.f), 30.f));
          k1.x = __expf(2.0f * fminf(fmaxf(k1.x, -30.f), 30.f)); k1.y = __expf(2.0f * fminf(fmaxf(k1.y, -30.f), 30.f));
          k1.z = __expf(2.0f * fminf(fmaxf(k1.z, -30.f), 30.f)); k1.w = __expf(2.0f * fminf(fmaxf(k1.w, -30.f), 30.f));
        }
        return quad(k0, e0, v0) + quad(k1, e1, v1);
      };
      for (int pp = 2 * warp; pp < npq; pp += 2 * NW) {
        const bool two = pp + 1 < npq;
        const float sa = pair_sum(pp), sb = pair_sum(two ? pp + 1 : pp);
        // lanes 0..15 end up with pair A's sum, lanes 16..31 with pair B's
        const bool up = lane >= 16;
        float sv = (up ? sb : sa) + __shfl_xor_sync(0xffffffffu, up ? sa : sb, 16);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
        const float ex = __expf(fmaxf(fmaf(-2.0f, sv, lds_f(sbase + OFF_VB)), -80.0f));
        if (lane == 0 || (lane == 16 && two)) sts_f(sbase + L.stage + (uint32_t)(up ? pp + 1 : pp) * 4u, ex);
      }
    }
    TRW(272);
    __syncthreads();
    TRM(16);
    // warp p -> peer p: this CTA's pairs into sc[p0 ..], four per DSMEM transaction (the tail of the last quad is padding)
    if (lane < ((npq + 3) >> 2)) st_async_v4(rx + L.sc + (uint32_t)p0 * 4u, lds128(sbase + L.stage + lane * 16), rmb0 + B_P6 * 8);
    RTAKE(T_P8, SL8)                      // window of P6: P8's chunks into registers
    TRW(288);
    TRM(17);
    mbar_wait(mb0 + B_P6 * 8, par);
    TRM(18);
    TRW(304);
    // ================= P7: context slice sum_j p_j memory[j][16q..16q+15] / sum_j p_j =================
    // lane (g, t): sample g, columns 4t..4t+3; warp w takes the positions j = w (mod 16)
    {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), acc2 = make_float4(0.f, 0.f, 0.f, 0.f);
      float ssum = 0.f, ssum2 = 0.f;
      if (g < S) {
        const int niter = (T_in - warp + NW - 1) / NW;
        if (res_m) {
          uint32_t pa = sbase + L.sc + (uint32_t)(warp * S + g) * 4u;
          uint32_t ma = sbase + L.msl + (uint32_t)((warp * S + g) * 16 + t * 4) * 4u;
          const uint32_t dp = (uint32_t)NW * S * 4u, dm = (uint32_t)NW * S * 64u;
          int i = 0;
          for (; i + 4 <= niter; i += 4) {
            const float p0v = lds_f(pa), p1v = lds_f(pa + dp), p2v = lds_f(pa + 2 * dp), p3v = lds_f(pa + 3 * dp);
            const float4 m0 = lds_f4(ma), m1 = lds_f4(ma + dm), m2 = lds_f4(ma + 2 * dm), m3 = lds_f4(ma + 3 * dm);
            acc.x = fmaf(p0v, m0.x, acc.x); acc.y = fmaf(p0v, m0.y, acc.y); acc.z = fmaf(p0v, m0.z, acc.z); acc.w = fmaf(p0v, m0.w, acc.w);
            acc2.x = fmaf(p1v, m1.x, acc2.x); acc2.y = fmaf(p1v, m1.y, acc2.y); acc2.z = fmaf(p1v, m1.z, acc2.z); acc2.w = fmaf(p1v, m1.w, acc2.w);
            acc.x = fmaf(p2v, m2.x, acc.x); acc.y = fmaf(p2v, m2.y, acc.y); acc.z = fmaf(p2v, m2.z, acc.z); acc.w = fmaf(p2v, m2.w, acc.w);
            acc2.x = fmaf(p3v, m3.x, acc2.x); acc2.y = fmaf(p3v, m3.y, acc2.y); acc2.z = fmaf(p3v, m3.z, acc2.z); acc2.w = fmaf(p3v, m3.w, acc2.w);
            ssum += p0v + p2v; ssum2 += p1v + p3v;
            pa += 4 * dp; ma += 4 * dm;
          }
          if (i + 2 <= niter) {
            const float p0v = lds_f(pa), p1v = lds_f(pa + dp);
            const float4 m0 = lds_f4(ma), m1 = lds_f4(ma + dm);
            acc.x = fmaf(p0v, m0.x, acc.x); acc.y = fmaf(p0v, m0.y, acc.y); acc.z = fmaf(p0v, m0.z, acc.z); acc.w = fmaf(p0v, m0.w, acc.w);
            acc2.x = fmaf(p1v, m1.x, acc2.x); acc2.y = fmaf(p1v, m1.y, acc2.y); acc2.z = fmaf(p1v, m1.z, acc2.z); acc2.w = fmaf(p1v, m1.w, acc2.w);
            ssum += p0v; ssum2 += p1v;
            pa += 2 * dp; ma += 2 * dm; i += 2;
          }
          if (i < niter) {
            const float p0v = lds_f(pa);
            const float4 m0 = lds_f4(ma);
            acc.x = fmaf(p0v, m0.x, acc.x); acc.y = fmaf(p0v, m0.y, acc.y); acc.z = fmaf(p0v, m0.z, acc.z); acc.w = fmaf(p0v, m0.w, acc.w);
            ssum += p0v;
          }
        } else {
          const float* sc = reinterpret_cast<const float*>(smem_raw + L.sc) + g;
          for (int j = warp; j < T_in; j += NW) {
            const float pv = sc[j * S];
            const float4 m0 = ldg_f4(a.memory + ((size_t)(n0 + g) * T_in + j) * DH + q * 16 + 4 * t);
            acc.x = fmaf(pv, m0.x, acc.x); acc.y = fmaf(pv, m0.y, acc.y); acc.z = fmaf(pv, m0.z, acc.z); acc.w = fmaf(pv, m0.w, acc.w);
            ssum += pv;
          }
        }
        acc.x += acc2.x; acc.y += acc2.y; acc.z += acc2.z; acc.w += acc2.w;
        ssum += ssum2;
      }
      sts_f4(myslot + (g * RS + t * 4) * 4, acc);   // partial context of sample g, columns 4t..4t+3
      if (t == 0) sts_f(sbase + OFF_REDS + (uint32_t)(warp * 8 + g) * 4u, ssum);
    }
    TRW(320);
    __syncthreads();
    TRM(19);
    if (red_grp) {
      const float* reds = reinterpret_cast<const float*>(smem_raw + OFF_REDS) + rn;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int s = 0; s < NW; s += 2) { s0 += reds[s * 8]; s1 += reds[s * 8 + 8]; }
      const float inv = rcp_approx(s0 + s1);   // >= T_in e^{-80} > 0 (1 ulp; the oracle divides)
      stage_x(stg_n, rc, red_sum<16>(red_nc, 0) * inv);
      if (rc == 0) reinterpret_cast<float*>(smem_raw + OFF_INV)[rn] = inv;
      send_rows(0, XBUF(DM_BC) + q * csb, B_P7, false);
    }
    TRM(20);
    mbar_wait(mb0 + B_P7 * 8, par);
    TRM(21);
    // ================= P8: y0 = [h_att' | ctx] W_p + b (ctx half here, h half from P5) =================
    MMA(SL8, T_P8)
    __syncthreads();
    TRM(22);
    if (red_grp) {
      stage_x(stg_n, rc, red_sum<8>(red_nc, 0) + lds_f(ST(ST_Y0H)) + BIAS(BI_PC));
      send_rows(0, XBUF(DM_BY0) + q * csb, B_P8, false);
    } else if (warp == 0 && a.align_out != nullptr) {
      // alignments of this CTA's pairs (tacotron.py:104: [N,T_in,steps]); the normalisers were written in P7 before the barrier above
      const float* stage = reinterpret_cast<const float*>(smem_raw + L.stage);
      const float* invs = reinterpret_cast<const float*>(smem_raw + OFF_INV);
      for (int pp = lane; pp < npq; pp += 32) {
        const int n = smem_raw[L.pn + pp], j = (p0 + pp - n) / S;
        a.align_out[((size_t)(n0 + n) * T_in + j) * a.max_steps + step] = stage[pp] * invs[n];
      }
    }
    RFILL(T_P8)
    TRW(160);
    TLOADP(T_P9, TC9, 0, 2, 3, 4, 5, 0, 1)   // window of P8: all of P9 from tensor memory, then its h1 rows
    TRW(144);
    TWAIT(SL9)
    TRW(64);
    MMA_PRE(0, SL9, T_P9)
    TRW(80);
    TRM(23);
    mbar_wait(mb0 + B_P8 * 8, par);
    TRM(24);
    // ----- P9 / P10: decoder GRU 1 on [y0 | h1], y1 = y0 + h1' -----
    GRU_GATES(0, SL9, T_P9, BI_R1, BI_U1, ST_H1, DM_BR1, B_P9, , W9_WORK)
    TRM(26);
    mbar_wait(mb0 + B_P9 * 8, par);
    TRM(27);
    TRW(96);
    trb = 112;
    TWAIT(SL10)
    GRU_CAND(SL10, T_P10, BI_C1, ST_H1, ST_H1, -1, DM_BH1, -1, B_P10, , W10_WORK)
    trb = -1;
    TRM(29);
    mbar_wait(mb0 + B_P10 * 8, par);
    TRM(30);
    // ----- P11 / P12: decoder GRU 2 on [y1 | h2], y2 = y1 + h2' -----
    GRU_GATES(1, SL11, T_P11, BI_R2, BI_U2, ST_H2, DM_BR2, B_P11, , W11_WORK)
    TRM(32);
    mbar_wait(mb0 + B_P11 * 8, par);
    TRM(33);
    TWAIT(SL12)
    GRU_CAND(SL12, T_P12, BI_C2, ST_H2, ST_H2, -1, DM_BH2, -1, B_P12, , W12_WORK)
    TRM(35);
    mbar_wait(mb0 + B_P12 * 8, par);
    TRM(36);
    // ================= P13: output projection tiles 2q, 2q+1 -> frames, feed the last frame back =================
    MMA_POST(2, SL13, T_P13)
    __syncthreads();
    TRM(37);
    if (warp >= 8) {   // warps 12-15: tile 2q, warps 8-11: tile 2q+1
      const int half = red_grp ? 0 : 1, tile = 2 * q + half;
      const float o = red_sum<8>(red_nc, half * 8) + BIAS(half ? BI_OB : BI_OA);
      if (tile < ntiles && rn < S) a.dec_out[((size_t)(n0 + rn) * a.max_steps + step) * Dout + tile * 16 + rc] = o;
      if (free_run) {   // next decoder input = last frame of the group (helpers.py:37)
        stage_x(stg_n + half * 512, rc, o);
        if (tile < ntiles && tile >= fb_tile0) send_rows(half, XBUF(DM_BF) + (tile - fb_tile0) * csb, B_P13, false);
      }
    }
    // the only second block barrier of a step: the next step's P1 overwrites the partial tiles, and in most CTAs (no
    // feedback tile) nothing else orders it after this reduction
    __syncthreads();
    RFILL(T_P13)                       // window of P13: refill, the frame rows of the next step's prenet into registers
    if (!early1) { RTAKE(T_P1, SL1) }
    if (early1) {                      // ... and the context rows of the next step's prenet, from TMEM
      load_tx(wb, TX1, (int)(WCNT(T_P1) & 7u), SL1());
      TWAIT(SL1)
      MMA_PRE(0, SL1, T_P1)
    } else { pre[0] = pre[1] = pre[2] = pre[3] = 0.f; }
    TRM(38);
  }
  // nobody may exit while a peer can still write into its shared memory
  if (free_run && a.steps > 0) mbar_wait(mb0 + B_P13 * 8, (uint32_t)(a.steps - 1) & 1u);
  asm volatile("cp.async.wait_group 0;" ::: "memory");   // nothing may still be in flight into this CTA's shared memory
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(lds32(sbase + OFF_TMEM)) : "memory");
}

}  // namespace

size_t decoder_mma_smem_bytes(int s_max, int T_in, int M, int att_res, int ring_d0, int ring_d1) {
  return make_dyn(s_max, T_in, M >> 4, att_res, ring_d0, ring_d1).total;
}

int decoder_mma_max_clusters() {
  auto kern = decoder_mma_kernel<false>;
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
  // shared memory: the attention operands resident if they fit, then the deepest weight ring that fits
  // (warps 0-7 consume up to 14 streamed chunk-tiles per step, 4 of them in one phase; warps 8-15 at most 6, 2 at a time)
  // residency mask: both, e^{2 keys} only (also saves the per-step exponentials), memory columns only, none
  const char* env = getenv("TACO_DEC_ATT_RES");
  const int allow = env ? (atoi(env) & 3) : 3;
  static const int masks[4] = {3, 1, 2, 0};
  static const int rings[3][2] = {{6, 3}, {5, 3}, {4, 2}};
  size_t smem = 0;
  bool ok = false;
  for (int m = 0; m < 4 && !ok; ++m) {
    const int res = masks[m];
    if ((res & allow) != res) continue;
    for (int k = 0; k < 3 && !ok; ++k) {
      smem = decoder_mma_smem_bytes(a.s_max, a.T_in, w.M, res, rings[k][0], rings[k][1]);
      if (smem <= 227 * 1024) { a.att_res = res; a.ring_d0 = rings[k][0]; a.ring_d1 = rings[k][1]; ok = true; }
    }
  }
  if (!ok) return cudaErrorInvalidValue;
  auto kern = a.trace != nullptr ? decoder_mma_kernel<true> : decoder_mma_kernel<false>;
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
