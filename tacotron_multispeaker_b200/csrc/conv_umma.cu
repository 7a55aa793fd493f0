// conv1d('same') / dense as a TMA-fed tcgen05 implicit GEMM (sm_100a).
//
// Same operator as conv_gemm.cu (reference models/modules.py:93-101, :10,
// :59-60, :79-89, models/tacotron.py:68,101) on the 5th-generation tensor
// cores.  D[128 rows (n, t0..t0+127)] x [128 output channels] accumulates in
// TMEM (fp32) over k-blocks of 64 input channels per filter tap:
//
//   A tile  = activations x[n, t0 + j - pad_left + (0..127), c0..c0+63]  (bf16)
//             one 3-D TMA box per (tap j, channel block); rows outside [0,T)
//             are zero-filled by the TMA unit -- that IS the 'same' padding,
//             so no im2col buffer and no boundary code.
//   B tile  = W^T[out channel o0..o0+127][j*Cp + c0 .. +63]  (bf16, K-major)
//
// fp32 parity: the reference computes in fp32, so each operand is split into
// bf16 hi + bf16 lo (x = hi + lo to ~2^-17) and every k-step issues the three
// products hi*hi + hi*lo + lo*hi into the same fp32 TMEM accumulator (the
// dropped lo*lo term is ~2^-18 relative).  NSPLIT=1 is the plain bf16 mode.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one
// elected thread), warps 2..5 = epilogue (TMEM -> registers -> smem transpose
// -> coalesced global stores, with bias / activation / folded BN / residual /
// highway gate).  3-stage smem ring (64 KB per stage) with full/empty
// mbarriers; tcgen05.commit releases stages and signals the epilogue.
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"

namespace taco {

namespace {
constexpr int BM = 128, BN = 128, BK = 64, MAX_STAGES = 3, NTHREADS = 192;
constexpr uint32_t TILE_BYTES = BM * BK * 2;            // 16 KB: one bf16 operand tile (A or B)
constexpr uint32_t STAGE_BYTES = 4 * TILE_BYTES;        // A_hi | A_lo | B_hi | B_lo
constexpr int EPI_LD = 36;                              // floats per staged row: 32 + 4 keeps 128-bit rows conflict free
constexpr uint32_t EPI_BYTES = 4 * 32 * EPI_LD * 4;     // per-warp 32 x 36 fp32 transpose staging (both epilogue paths)
// dynamic smem: 1024 (alignment slack) + stages * 64 KB + [transpose staging] + barriers
__host__ __device__ constexpr uint32_t smem_bytes(int stages, bool transpose_epi) {
  return 1024u + (uint32_t)stages * STAGE_BYTES + (transpose_epi ? EPI_BYTES : 0u) + 256u;
}

struct UmmaArgs {
  int N, T, Cp;            // activation rows / padded channels (Cp % 64 == 0)
  int taps, bank;          // bank > 1: conv index ci = bank-1-blockIdx.z has ci+1 taps
  int Cout;                // output channels per conv
  const float* bias; const float* scale; const float* shift;
  const float* res; long long res_bs; int ldres;
  float* out; long long out_bs; int ldo; int col_off;
  int act, epi;
  int stages;              // smem ring depth (1..3)
  int vec_epi;             // 1: rows are 16 B aligned -> direct 128-bit stores from the TMEM registers
  int nx, ny, ntiles;      // tile list: nx = N * ceil(T / 128) row tiles, ny column tiles, ntiles = nx * ny * bank
  int acc_cols;            // TMEM columns to allocate: 256 (two accumulators, persistent CTAs) or 128 (one tile per CTA)
};

// ---- PTX wrappers -------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t mb, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mb, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mb) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(mb), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t mb, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(mb), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t mb, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(mb), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mb) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb) : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows are 128 B apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int NSPLIT>
__global__ void __launch_bounds__(NTHREADS, 3)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const UmmaArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const int STAGES = p.stages;
  const uint32_t epi_bytes = EPI_BYTES;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  float* epi = reinterpret_cast<float*>(smem_al + STAGES * STAGE_BYTES);
  const uint32_t bar0 = smem_base + STAGES * STAGE_BYTES + epi_bytes;    // full[S], empty[S], tmem_full[2], tmem_empty[2], tmem ptr
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + 2 + b); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_al + STAGES * STAGE_BYTES + epi_bytes + 8 * (2 * MAX_STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tps = (p.T + BM - 1) / BM;
  const int kcb = p.Cp / BK;
  // Persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ... in the order (conv of the bank, heavy first; column
  // tile; row tile fastest, so that the CTAs working at the same time share their weight tiles in L2).  The TMA, MMA and
  // epilogue roles each walk the same list; the accumulator alternates between two 128-column TMEM buffers, so the
  // epilogue of tile i overlaps the loads and MMAs of tile i+1.
  struct Tile { int n, t0, o0, ci, taps, pl, nkb, brow0; };
  auto decode = [&](int L) {
    Tile q;
    const int x = L % p.nx, r = L / p.nx, y = r % p.ny, z = r / p.ny;   // row tile fastest: neighbours share the weight tile
    q.n = x / tps; q.t0 = (x - q.n * tps) * BM; q.o0 = y * BN;
    q.ci = p.bank > 1 ? p.bank - 1 - z : 0;                               // heavy convs first
    q.taps = p.bank > 1 ? q.ci + 1 : p.taps;
    q.pl = (q.taps - 1) >> 1;
    q.nkb = q.taps * kcb;
    q.brow0 = q.ci * p.Cout + q.o0;
    return q;
  };

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB_hi)) : "memory");
    if (NSPLIT > 1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA_lo)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB_lo)) : "memory");
    }
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 4); }   // 4 epilogue warps release a buffer
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {   // TMEM: two 128-column fp32 accumulators, allocated (and later freed) by this warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"((uint32_t)p.acc_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;                                                         // k-block counter over all tiles of this CTA
      for (int L = blockIdx.x; L < p.ntiles; L += gridDim.x) {
        const Tile q = decode(L);
        for (int kb = 0; kb < q.nkb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(empty_bar(s), ((it / STAGES) & 1) ^ 1);
          const uint32_t st = smem_base + s * STAGE_BYTES;
          mbar_expect_tx(full_bar(s), NSPLIT > 1 ? STAGE_BYTES : 2 * TILE_BYTES);
          const int j = kb / kcb, c0 = (kb - j * kcb) * BK;
          const int tt = q.t0 + j - q.pl;                                 // may be < 0 or run past T: zero fill
          tma_load_3d(st, &tmA_hi, full_bar(s), c0, tt, q.n);
          tma_load_2d(st + 2 * TILE_BYTES, &tmB_hi, full_bar(s), j * p.Cp + c0, q.brow0);
          if (NSPLIT > 1) {
            tma_load_3d(st + TILE_BYTES, &tmA_lo, full_bar(s), c0, tt, q.n);
            tma_load_2d(st + 3 * TILE_BYTES, &tmB_lo, full_bar(s), j * p.Cp + c0, q.brow0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // kind::f16: C=F32 (bit4), A=BF16 (bit7), B=BF16 (bit10), both K-major, N>>3 at bit17, M>>4 at bit24
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int it = 0, i = 0;
      for (int L = blockIdx.x; L < p.ntiles; L += gridDim.x, ++i) {
        const Tile q = decode(L);
        const int b = i & 1;
        mbar_wait(tempty_bar(b), ((i >> 1) & 1) ^ 1);                     // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(b * BN);
        for (int kb = 0; kb < q.nkb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(full_bar(s), (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t st = smem_base + s * STAGE_BYTES;
          const uint64_t a_hi = umma_desc(st), a_lo = umma_desc(st + TILE_BYTES);
          const uint64_t b_hi = umma_desc(st + 2 * TILE_BYTES), b_lo = umma_desc(st + 3 * TILE_BYTES);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 32 >> 4);                // 16 bf16 = 32 B along K inside the swizzle row
            umma_bf16(acc, a_hi + adv, b_hi + adv, idesc, (kb | kk) != 0);
            if (NSPLIT > 1) {
              umma_bf16(acc, a_hi + adv, b_lo + adv, idesc, 1u);
              umma_bf16(acc, a_lo + adv, b_hi + adv, idesc, 1u);
            }
          }
          umma_commit(empty_bar(s));                                      // stage free once these MMAs have read it
        }
        umma_commit(tfull_bar(b));                                        // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;                                         // TMEM lanes this warp may read
    int i = 0;
    for (int L = blockIdx.x; L < p.ntiles; L += gridDim.x, ++i) {
    const Tile q = decode(L);
    const int n = q.n, t0 = q.t0, o0 = q.o0, ci = q.ci;
    const int ab = i & 1;
    const uint32_t acc = tmem_base + (uint32_t)(ab * BN);
    mbar_wait(tfull_bar(ab), (i >> 1) & 1);
    tc_fence_after();
    const float* bias = p.bias ? p.bias + ci * p.Cout : nullptr;
    const float* scale = p.scale ? p.scale + ci * p.Cout : nullptr;
    const float* shift = p.shift ? p.shift + ci * p.Cout : nullptr;
    const int col_off = p.col_off + ci * p.Cout;
    if (p.vec_epi) {
      // ---- vector path: a thread reads ONE ROW of the accumulator from TMEM (32 columns per load).  Storing from there
      // would touch 32 different rows per instruction (32 half-used sectors); the chunk is transposed through a
      // 32 x 36 shared-memory tile instead, so that a store instruction covers 4 rows x 128 contiguous bytes and the
      // residual is read the same way.  Bias / activation / BN affine / residual / highway gate run after the transpose.
      float* stg = epi + (warp - 2) * 32 * EPI_LD;
      const int tq = t0 + quarter * 32;
      const int rsub = lane >> 3, c4 = lane & 7;                          // after the transpose: row 4 i + rsub, columns 4 c4 .. 4 c4 + 3
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        const int cbase = o0 + ch * 32;
        if (cbase >= p.Cout) break;                                       // warp-uniform
        uint32_t v[32];
        tmem_ld32(acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ch * 32), v);
        if (p.epi == EPI_PLAIN) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<float4*>(stg + lane * EPI_LD + 4 * g) =
                make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
          __syncwarp();
          const int col = cbase + 4 * c4;
          const bool cok = col < p.Cout;                                  // Cout % 4 == 0 on this path
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), sc4 = make_float4(1.f, 1.f, 1.f, 1.f), sh4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cok) {
            if (bias) b4 = ldg_f4(bias + col);
            if (scale) { sc4 = ldg_f4(scale + col); sh4 = ldg_f4(shift + col); }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + rsub, t = tq + r;
            float4 x = *reinterpret_cast<const float4*>(stg + r * EPI_LD + 4 * c4);
            x.x = apply_act(x.x + b4.x, p.act); x.y = apply_act(x.y + b4.y, p.act);
            x.z = apply_act(x.z + b4.z, p.act); x.w = apply_act(x.w + b4.w, p.act);
            if (scale) { x.x = fmaf(x.x, sc4.x, sh4.x); x.y = fmaf(x.y, sc4.y, sh4.y); x.z = fmaf(x.z, sc4.z, sh4.z); x.w = fmaf(x.w, sc4.w, sh4.w); }
            if (cok && t < p.T) {
              if (p.res) {
                const float4 r4 = ldg_f4(p.res + (long long)n * p.res_bs + (long long)t * p.ldres + col);
                x.x += r4.x; x.y += r4.y; x.z += r4.z; x.w += r4.w;
              }
              *reinterpret_cast<float4*>(p.out + (long long)n * p.out_bs + (long long)t * p.ldo + col_off + col) = x;
            }
          }
        } else {   // EPI_HIGHWAY: columns (2c, 2c+1) = (H_c, T_c); 16 channels per 32-column load
          // staged per row: a_c = relu(H_c) * sigmoid(T_c) at floats 0..15, b_c = 1 - sigmoid(T_c) at floats 16..31
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int colb = cbase + 8 * g;
            const bool bok = colb < p.Cout;                               // Cout % 8 == 0 on this path
            const float4 b0 = bok ? ldg_f4(bias + colb) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 b1 = bok ? ldg_f4(bias + colb + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float a4[4], g4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int c2 = 8 * g + 2 * k;
              const float H = fmaxf(__uint_as_float(v[c2]) + bb[2 * k], 0.f), Tg = sigmoid_f(__uint_as_float(v[c2 + 1]) + bb[2 * k + 1]);
              a4[k] = H * Tg; g4[k] = 1.0f - Tg;
            }
            *reinterpret_cast<float4*>(stg + lane * EPI_LD + 4 * g) = make_float4(a4[0], a4[1], a4[2], a4[3]);
            *reinterpret_cast<float4*>(stg + lane * EPI_LD + 16 + 4 * g) = make_float4(g4[0], g4[1], g4[2], g4[3]);
          }
          __syncwarp();
          const int chn0 = cbase >> 1;                                    // first of the 16 output channels of this chunk
          const int rs2 = lane >> 2, q4 = lane & 3;                       // row 8 i + rs2, channels chn0 + 4 q4 ..
          const bool cok = cbase + 8 * q4 < p.Cout;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = 8 * i + rs2, t = tq + r;
            const float4 av = *reinterpret_cast<const float4*>(stg + r * EPI_LD + 4 * q4);
            const float4 gv = *reinterpret_cast<const float4*>(stg + r * EPI_LD + 16 + 4 * q4);
            if (cok && t < p.T) {
              const float4 xin = ldg_f4(p.res + (long long)n * p.res_bs + (long long)t * p.ldres + chn0 + 4 * q4);
              *reinterpret_cast<float4*>(p.out + (long long)n * p.out_bs + (long long)t * p.ldo + col_off + chn0 + 4 * q4) =
                  make_float4(fmaf(xin.x, gv.x, av.x), fmaf(xin.y, gv.y, av.y), fmaf(xin.z, gv.z, av.z), fmaf(xin.w, gv.w, av.w));
            }
          }
        }
        __syncwarp();
      }
    } else {
      // ---- transpose path (rows not 16 B aligned, e.g. ldo = 1025): smem transpose -> coalesced scalar stores ----
      float* stg = epi + (warp - 2) * 32 * EPI_LD;
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        const int cbase = o0 + ch * 32;
        if (cbase >= p.Cout) break;                                       // warp-uniform
        uint32_t v[32];
        tmem_ld32(acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ch * 32), v);
#pragma unroll
        for (int c = 0; c < 32; ++c) stg[lane * EPI_LD + c] = __uint_as_float(v[c]);
        __syncwarp();
        const int col = cbase + lane;
        const bool cok = col < p.Cout;
        float b = 0.f, sc = 1.f, sh = 0.f;
        if (cok) {
          if (bias) b = __ldg(bias + col);
          if (scale) { sc = __ldg(scale + col); sh = __ldg(shift + col); }
        }
        const int rmax = min(32, p.T - (t0 + quarter * 32));
        // row pointers advance by the leading dimension: no 64-bit index arithmetic per element
        const int tq = t0 + quarter * 32;
        if (p.epi == EPI_PLAIN) {
          float* optr = p.out + (long long)n * p.out_bs + (long long)tq * p.ldo + col_off + col;
          const float* rptr = p.res ? p.res + (long long)n * p.res_bs + (long long)tq * p.ldres + col : nullptr;
          const bool has_scale = scale != nullptr;
#pragma unroll 8
          for (int r = 0; r < rmax; ++r) {
            float x = apply_act(stg[r * EPI_LD + lane] + b, p.act);
            if (has_scale) x = fmaf(x, sc, sh);
            if (cok) {
              if (rptr) x += __ldg(rptr);
              *optr = x;
            }
            optr += p.ldo;
            if (rptr) rptr += p.ldres;
          }
        } else {   // EPI_HIGHWAY: even lane = H_c, odd lane = T_c of channel c = col/2
          const int chn = col >> 1;
          float* optr = p.out + (long long)n * p.out_bs + (long long)tq * p.ldo + col_off + chn;
          const float* rptr = p.res + (long long)n * p.res_bs + (long long)tq * p.ldres + chn;
#pragma unroll 4
          for (int r = 0; r < rmax; ++r) {
            const float x = stg[r * EPI_LD + lane] + b;
            const float tg = __shfl_down_sync(0xffffffffu, x, 1);
            if (cok && !(lane & 1)) {
              const float H = fmaxf(x, 0.f), Tg = sigmoid_f(tg);
              const float xin = __ldg(rptr);
              *optr = H * Tg + xin * (1.0f - Tg);
            }
            optr += p.ldo;
            rptr += p.ldres;
          }
        }
        __syncwarp();
      }
    }
    // this warp has read its lanes of the accumulator: hand the buffer back to the MMA issuer
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty_bar(ab));
    }   // tile loop
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.acc_cols) : "memory");
  }
}

// fp32 -> bf16 hi + bf16 lo (x ~= hi + lo), channels zero-padded to Cp.
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ x, long long x_bs, int ldx, int N, int T, int C, int Cp,
                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int Cq = Cp >> 2;
  const long long total = (long long)N * T * Cq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % Cq);
    const long long row = i / Cq;
    const int t = (int)(row % T), n = (int)(row / T);
    const int c = q * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c + 3 < C) v = ldg_f4(x + n * x_bs + (long long)t * ldx + c);
    else if (c < C) {
      const float* px = x + n * x_bs + (long long)t * ldx;
      v.x = px[c];
      if (c + 1 < C) v.y = px[c + 1];
      if (c + 2 < C) v.z = px[c + 2];
    }
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(v.x - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v.y - __bfloat162float(h1));
    const __nv_bfloat16 l2 = __float2bfloat16_rn(v.z - __bfloat162float(h2)), l3 = __float2bfloat16_rn(v.w - __bfloat162float(h3));
    __nv_bfloat162 hA, hB, lA, lB;
    hA.x = h0; hA.y = h1; hB.x = h2; hB.y = h3; lA.x = l0; lA.y = l1; lB.x = l2; lB.y = l3;
    uint2 ph, plo;
    ph.x = *reinterpret_cast<uint32_t*>(&hA); ph.y = *reinterpret_cast<uint32_t*>(&hB);
    plo.x = *reinterpret_cast<uint32_t*>(&lA); plo.y = *reinterpret_cast<uint32_t*>(&lB);
    reinterpret_cast<uint2*>(hi)[i] = ph;
    reinterpret_cast<uint2*>(lo)[i] = plo;
  }
}

__global__ void __launch_bounds__(256)
pack_wt_kernel(const float* __restrict__ w, int ldw, int taps, int Cin, int Cout, int Cp, int Kld, int row0,
               __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long total = (long long)Cout * taps * Cp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const int j = (int)((i / Cp) % taps);
    const int o = (int)(i / ((long long)Cp * taps));
    const float v = c < Cin ? __ldg(w + (long long)(j * Cin + c) * ldw + o) : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const long long dst = (long long)(row0 + o) * Kld + (long long)j * Cp + c;
    hi[dst] = h;
    lo[dst] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// ---- host: tensor maps ------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_tmapEncodeTiled get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}
bool make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
              const cuuint32_t* box) {
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint32_t es[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

void launch_split_bf16(const float* x, int64_t x_bs, int ldx, int N, int T, int C, int Cp, void* hi, void* lo,
                       cudaStream_t st) {
  const long long total = (long long)N * T * (Cp / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  split_bf16_kernel<<<blocks, 256, 0, st>>>(x, x_bs, ldx, N, T, C, Cp, reinterpret_cast<__nv_bfloat16*>(hi),
                                            reinterpret_cast<__nv_bfloat16*>(lo));
}

void launch_pack_wt(const float* w, int ldw, int taps, int Cin, int Cout, int Cp, int Kld, int row0, void* hi,
                    void* lo, cudaStream_t st) {
  const long long total = (long long)Cout * taps * Cp;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  pack_wt_kernel<<<blocks, 256, 0, st>>>(w, ldw, taps, Cin, Cout, Cp, Kld, row0, reinterpret_cast<__nv_bfloat16*>(hi),
                                         reinterpret_cast<__nv_bfloat16*>(lo));
}

cudaError_t launch_conv_umma(const ConvUmma& c, cudaStream_t st) {
  if (c.N <= 0 || c.T <= 0) return cudaSuccess;
  if (c.Cp % BK || c.Kld % BK) return cudaErrorInvalidValue;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  const cuuint64_t adims[3] = {(cuuint64_t)c.Cp, (cuuint64_t)c.T, (cuuint64_t)c.N};
  const cuuint64_t astr[2] = {(cuuint64_t)c.Cp * 2, (cuuint64_t)c.T * c.Cp * 2};
  const cuuint32_t abox[3] = {BK, BM, 1};
  const cuuint64_t bdims[2] = {(cuuint64_t)c.Kld, (cuuint64_t)c.b_rows};
  const cuuint64_t bstr[1] = {(cuuint64_t)c.Kld * 2};
  const cuuint32_t bbox[2] = {BK, BN};
  const bool split = c.nsplit > 1;
  if (!make_map(&ma_hi, c.a_hi, 3, adims, astr, abox) || !make_map(&mb_hi, c.b_hi, 2, bdims, bstr, bbox) ||
      !make_map(&ma_lo, split ? c.a_lo : c.a_hi, 3, adims, astr, abox) ||
      !make_map(&mb_lo, split ? c.b_lo : c.b_hi, 2, bdims, bstr, bbox))
    return cudaErrorInvalidValue;
  UmmaArgs p;
  p.N = c.N; p.T = c.T; p.Cp = c.Cp; p.taps = c.taps; p.bank = c.bank; p.Cout = c.Cout;
  p.bias = c.bias; p.scale = c.scale; p.shift = c.shift; p.res = c.res; p.res_bs = c.res_bs; p.ldres = c.ldres;
  p.out = c.out; p.out_bs = c.out_bs; p.ldo = c.ldo; p.col_off = c.col_off; p.act = c.act; p.epi = c.epi;
  // Persistent grid: one CTA per SM (three 64 KB stages), each walking its share of the tiles.
  p.nx = c.N * ((c.T + BM - 1) / BM);
  p.ny = (c.Cout + BN - 1) / BN;
  const long long ntiles = (long long)p.nx * p.ny * (c.bank > 1 ? c.bank : 1);
  if (ntiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  p.ntiles = (int)ntiles;
  // Long-K launches (the 3 x 1024 -> 256 projection, the encoder bank) are bound by loads and MMAs: persistent CTAs,
  // one per SM with three 64 KB stages, overlap the epilogue of tile i with the main loop of tile i+1 (171 -> 140 us).
  // Short-K launches are bound by the epilogue (four warps per CTA drain TMEM and write 64 KB per tile): there one tile
  // per CTA with a shallow ring lets two or three CTAs share an SM, i.e. 8-12 epilogue warps (persistent: 158 -> 221 us
  // on the final dense, measured).
  const int nkb_max = c.taps * (c.Cp / BK);
  static const int force = [] { const char* e = getenv("TACO_UMMA_PERSISTENT"); return e ? atoi(e) : -1; }();
  const bool persistent = force >= 0 ? force != 0 : nkb_max >= 24;
  if (persistent) {
    p.stages = MAX_STAGES;
  } else {
    p.stages = nkb_max >= 6 ? 3 : (nkb_max >= 3 ? 2 : 1);
    if (ntiles >= 1024) p.stages = 1;
  }
  {   // developer switch: TACO_UMMA_STAGES_MAX caps the ring depth
    static const int cap = [] { const char* e = getenv("TACO_UMMA_STAGES_MAX"); return e ? atoi(e) : 0; }();
    if (cap >= 1 && p.stages > cap) p.stages = cap;
  }
  static int n_sm[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (n_sm[dev & 63] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n_sm[dev & 63] = v;
  }
  static const int cta_cap = [] { const char* e = getenv("TACO_UMMA_CTAS"); return e ? atoi(e) : 0; }();
  int nctas = persistent ? (cta_cap > 0 ? cta_cap : n_sm[dev & 63]) : p.ntiles;
  if (nctas > p.ntiles) nctas = p.ntiles;
  p.acc_cols = nctas < p.ntiles ? 256 : 128;
  dim3 grid(nctas, 1, 1);
  const int chn = c.epi == EPI_HIGHWAY ? 2 : 1;
  auto al4 = [](long long v) { return (v & 3) == 0; };
  p.vec_epi = (al4(c.ldo) && al4(c.col_off) && al4(c.out_bs) && al4(c.Cout / chn) && (c.Cout % (4 * chn) == 0) &&
               (reinterpret_cast<uintptr_t>(c.out) & 15) == 0 &&
               (c.res == nullptr || (al4(c.ldres) && al4(c.res_bs) && (reinterpret_cast<uintptr_t>(c.res) & 15) == 0)) &&
               (c.bias == nullptr || (reinterpret_cast<uintptr_t>(c.bias) & 15) == 0) &&
               (c.scale == nullptr || ((reinterpret_cast<uintptr_t>(c.scale) & 15) == 0 && (reinterpret_cast<uintptr_t>(c.shift) & 15) == 0)))
                  ? 1 : 0;
  if (c.epi == EPI_HIGHWAY && (c.bias == nullptr || c.res == nullptr)) return cudaErrorInvalidValue;
  const uint32_t smem = smem_bytes(p.stages, true);   // both epilogue paths stage through shared memory
  static bool attr_done[64] = {false};
  bool& attr_set = attr_done[dev & 63];
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(conv_umma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(MAX_STAGES, true));
    cudaError_t e2 = cudaFuncSetAttribute(conv_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(MAX_STAGES, true));
    if (e1 != cudaSuccess) return e1;
    if (e2 != cudaSuccess) return e2;
    attr_set = true;
  }
  if (split) conv_umma_kernel<3><<<grid, NTHREADS, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
  else conv_umma_kernel<1><<<grid, NTHREADS, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
  return cudaGetLastError();
}

}  // namespace taco
