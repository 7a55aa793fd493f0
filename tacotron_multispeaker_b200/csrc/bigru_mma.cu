// K6b (tensor-core variant): recurrence of the CBHG bidirectional GRU, EIGHT utterances per CTA.
//
// Same operator as bigru.cu (tf.nn.bidirectional_dynamic_rnn(GRUCell(128), GRUCell(128), x, sequence_length),
// reference models/modules.py:68-74; input halves of the kernels hoisted into one GEMM):
//     [r|u] = sigmoid(xg_t + h U_g)      U_g [128,256]
//     c     = tanh  (xc_t + (r*h) U_c)   U_c [128,128]   (reset BEFORE matmul: TF GRUCell)
//     h'    = u*h + (1-u)*c
// bigru.cu keeps one utterance per CTA with the weights in registers and is bound by instruction issue (201
// instructions per warp and step), so a batch of 32 holds 64 SMs for the 1000 steps of the post-net.  Here the
// mat-vecs run on mma.sync m16n8k16 with the eight accumulator columns = eight utterances (bf16 hi/lo split operands,
// hh + lh + hl products: fp32-class accuracy like the decoder) and the recurrent weights live in TENSOR MEMORY
// (192 chunk-tiles of 1 KB, fetched with tcgen05.ld.32x32b.x8 straight into the A registers), so one CTA per
// (direction, 8 utterances) does the work: 8 CTAs instead of 64 at about the same time per step (the step is bound by
// the tensor pipe: 576 HMMA / 4 sub-partitions x 8 clk).
//
//   gates     : warp w owns gate columns 16w..16w+15 (w < 8: r, else u), all 8 k-chunks          -> one barrier
//   candidate : warp w < 8 owns candidate columns 16w..16w+15, all 8 k-chunks of r*h             -> one barrier
// h and r*h sit in shared memory in MMA-fragment order X[k-chunk][utterance][16 words] (hi pairs, lo pairs), which is
// what the producing lanes hold after one shuffle with the neighbouring column.
#include <cuda_bf16.h>
#include <string.h>

#include "common.cuh"
#include "kernels.cuh"

namespace taco {

namespace {
constexpr int H = 128, XW = 768, NS = 8, NT = 512, NW = 16;
constexpr int RING = 11;                         // steps of hoisted projections in flight (also keeps GEMM CTAs, whose
                                                 // TMEM allocation would wait for ours, off this SM: 148 KB of smem)
constexpr uint32_t B_X = 0;                      // h fragments      [8 chunks][8 utterances][64 B]
constexpr uint32_t B_RH = 4096;                  // r*h fragments
constexpr uint32_t B_HS = 8192;                  // h fp32           [8][128]
constexpr uint32_t B_US = 12288;                 // update gate fp32 [8][128]
constexpr uint32_t B_MISC = 16384;               // TMEM base (4 B), lengths (32 B)
constexpr uint32_t B_RING = 16384 + 128;         // [RING][8][384] fp32
constexpr uint32_t SMEM_BYTES = B_RING + RING * NS * 384 * 4;

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float lds_f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint4& hi, uint4& lo) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w), "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4& hi, const uint4& lo) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(taddr), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
}
// registers written by tcgen05.ld are defined only after tcgen05.wait::ld: make the compiler see them produced there
__device__ __forceinline__ void tmem_wait_ld(uint4& a, uint4& b) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a.x), "+r"(a.y), "+r"(a.z), "+r"(a.w), "+r"(b.x), "+r"(b.y), "+r"(b.z), "+r"(b.w) :: "memory");
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// fp32 -> (bf16 hi) | (bf16 lo) << 16 with x ~= hi + lo
__device__ __forceinline__ uint32_t hilo(float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  return (uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(l) << 16);
}

// One 16-column tile times the eight utterances over all 8 k-chunks: A from tensor memory (chunk-tiles tc0 .. tc0+7 of
// this warp's slice, fetched one chunk ahead), B = fragment-order activations at xbase.  d = hh + (hl + lh).
__device__ __forceinline__ void tile_matvec(float (&d)[4], uint32_t tw, int tc0, uint32_t xbase, int g, int t) {
  float hh[4] = {0.f, 0.f, 0.f, 0.f}, hl[4] = {0.f, 0.f, 0.f, 0.f}, lh[4] = {0.f, 0.f, 0.f, 0.f};
  uint4 ahi[2], alo[2];
  tmem_ld8(tw + (uint32_t)tc0 * 8u, ahi[0], alo[0]);
  const uint32_t xl = xbase + (uint32_t)g * 64u + (uint32_t)t * 16u;
#pragma unroll
  for (int kc = 0; kc < 8; ++kc) {
    const int cur = kc & 1;
    tmem_wait_ld(ahi[cur], alo[cur]);
    if (kc + 1 < 8) tmem_ld8(tw + (uint32_t)(tc0 + kc + 1) * 8u, ahi[cur ^ 1], alo[cur ^ 1]);
    const uint4 xf = lds128(xl + (uint32_t)kc * 512u);
    mma16816(hh, ahi[cur], xf.x, xf.y);   // W_hi * x_hi
    mma16816(lh, alo[cur], xf.x, xf.y);   // W_lo * x_hi
    mma16816(hl, ahi[cur], xf.z, xf.w);   // W_hi * x_lo
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) d[k] = hh[k] + (hl[k] + lh[k]);
}

// Lane (g, t) holds v[0..3] = (column g, utterances 2t, 2t+1), (column g+8, utterances 2t, 2t+1) of a 16-column chunk:
// write them as bf16 hi/lo into the fragment-order buffer X[chunk][utterance][16 words].  Columns (2j, 2j+1) share a
// word, so neighbouring g exchange one utterance each: the even lane writes utterance 2t, the odd one 2t+1.
__device__ __forceinline__ void store_frag(uint32_t buf, int chunk, int g, int t, const float (&v)[4]) {
  const uint32_t pa = hilo(v[0]), pb = hilo(v[1]), qa = hilo(v[2]), qb = hilo(v[3]);
  const bool even = (g & 1) == 0;
  const uint32_t rp = __shfl_xor_sync(0xffffffffu, even ? pb : pa, 4);
  const uint32_t rq = __shfl_xor_sync(0xffffffffu, even ? qb : qa, 4);
  const uint32_t p0 = even ? pa : rp, p1 = even ? rp : pb;   // columns 2j, 2j+1
  const uint32_t q0 = even ? qa : rq, q1 = even ? rq : qb;   // columns 2j+8, 2j+9
  const int s = 2 * t + (even ? 0 : 1);
  sts128(buf + (uint32_t)chunk * 512u + (uint32_t)s * 64u + (uint32_t)(g >> 1) * 16u,
         (p0 & 0xffffu) | (p1 << 16), (q0 & 0xffffu) | (q1 << 16), (p0 >> 16) | (p1 & 0xffff0000u), (q0 >> 16) | (q1 & 0xffff0000u));
}

__global__ void __launch_bounds__(NT, 1)
bigru_mma_kernel(const float* __restrict__ xproj, const uint4* __restrict__ frag, const int32_t* __restrict__ lengths,
                 int N, int T, float* __restrict__ out, int64_t out_bs) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int g = lane >> 2, t = lane & 3;
  const int dir = blockIdx.y, n0 = blockIdx.x * NS;
  const uint32_t sbase = smem_u32(smem_raw);
  int* s_len = reinterpret_cast<int*>(smem_raw + B_MISC + 16);

  for (uint32_t i = tid * 4; i < B_MISC; i += NT * 4) *reinterpret_cast<uint32_t*>(smem_raw + i) = 0u;
  if (tid < NS) {
    int L = 0;
    if (n0 + tid < N) {
      L = lengths ? lengths[n0 + tid] : T;
      L = max(0, min(L, T));
    }
    s_len[tid] = L;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sbase + B_MISC) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sbase + B_MISC));
  // this warp's slice of tensor memory: lane quarter warp % 4 (hardware rule), 128 columns = 16 chunk-tiles:
  // 0-7 the gate tile of this warp, 8-15 (warps 0-7) its candidate tile
  const uint32_t tw = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 128u;
  {
    const uint4* src = frag + ((size_t)(dir * NW + warp) * 16) * 64 + lane;   // [dir][warp][16 chunk-tiles][hi 32 | lo 32] uint4
    const int ntiles = warp < 8 ? 16 : 8;
    for (int i0 = 0; i0 < ntiles; i0 += 4) {
      uint4 hi[4], lo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { hi[k] = __ldg(src + (size_t)(i0 + k) * 64); lo[k] = __ldg(src + (size_t)(i0 + k) * 64 + 32); }
#pragma unroll
      for (int k = 0; k < 4; ++k) tmem_st8(tw + (uint32_t)(i0 + k) * 8u, hi[k], lo[k]);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  int maxlen = 0;
#pragma unroll
  for (int s = 0; s < NS; ++s) maxlen = max(maxlen, s_len[s]);

  // outputs are zero for t >= len (dynamic_rnn zero_output); this direction's 128 columns
#pragma unroll 1
  for (int s = 0; s < NS; ++s) {
    if (n0 + s >= N) continue;
    float* o = out + (int64_t)(n0 + s) * out_bs + dir * H;
    for (int i = s_len[s] * H + tid; i < T * H; i += NT) o[(int64_t)(i >> 7) * (2 * H) + (i & 127)] = 0.f;
  }

  // hoisted projections: 8 utterances x 96 float4 per step through a RING-deep cp.async ring
  auto issue = [&](int step) {
    for (int i = tid; i < NS * 96; i += NT) {
      const int s = i / 96, q4 = i - s * 96;
      const int L = s_len[s];
      if (step < L) {
        const int pos = dir == 0 ? step : L - 1 - step;
        cp_async16(sbase + B_RING + (uint32_t)(((step % RING) * NS + s) * 384 + q4 * 4) * 4u,
                   xproj + ((int64_t)(n0 + s) * T + pos) * XW + dir * (3 * H) + q4 * 4);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll 1
  for (int p = 0; p < RING - 1; ++p) issue(p);
  asm volatile("cp.async.wait_group %0;" ::"n"(RING - 2) : "memory");
  __syncthreads();

  const int col0 = 16 * (warp & 7) + g;   // this lane's columns inside the r / u / candidate block: col0, col0 + 8
#pragma unroll 1
  for (int step = 0; step < maxlen; ++step) {
    issue(step + RING - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(RING - 2) : "memory");   // this thread's part of step+1 has landed; the
                                                                           // two barriers below publish it
    const uint32_t xr = sbase + B_RING + (uint32_t)((step % RING) * NS * 384) * 4u;
    // ---- gates ----
    {
      float d[4];
      tile_matvec(d, tw, 0, sbase + B_X, g, t);
      const uint32_t xa = xr + (uint32_t)(2 * t * 384 + (warp < 8 ? 0 : H) + col0) * 4u;   // utterance 2t; +384 floats: 2t+1
      float gate[4];
      gate[0] = sigmoid_f(d[0] + lds_f(xa));
      gate[1] = sigmoid_f(d[1] + lds_f(xa + 384 * 4));
      gate[2] = sigmoid_f(d[2] + lds_f(xa + 8 * 4));
      gate[3] = sigmoid_f(d[3] + lds_f(xa + (384 + 8) * 4));
      if (warp < 8) {   // reset gate: r * h in fragment order, chunk = this warp's 16 columns
        const uint32_t ha = sbase + B_HS + (uint32_t)(2 * t * H + col0) * 4u;
        float rh[4];
        rh[0] = gate[0] * lds_f(ha);
        rh[1] = gate[1] * lds_f(ha + H * 4);
        rh[2] = gate[2] * lds_f(ha + 8 * 4);
        rh[3] = gate[3] * lds_f(ha + (H + 8) * 4);
        store_frag(sbase + B_RH, warp, g, t, rh);
      } else {          // update gate, fp32
        const uint32_t ua = sbase + B_US + (uint32_t)(2 * t * H + col0) * 4u;
        sts_f(ua, gate[0]); sts_f(ua + H * 4, gate[1]); sts_f(ua + 8 * 4, gate[2]); sts_f(ua + (H + 8) * 4, gate[3]);
      }
    }
    __syncthreads();             // r*h and u complete
    // ---- candidate and state update (warps 0-7) ----
    if (warp < 8) {
      float d[4];
      tile_matvec(d, tw, 8, sbase + B_RH, g, t);
      const uint32_t xa = xr + (uint32_t)(2 * t * 384 + 2 * H + col0) * 4u;
      const uint32_t ha = sbase + B_HS + (uint32_t)(2 * t * H + col0) * 4u;
      const uint32_t ua = sbase + B_US + (uint32_t)(2 * t * H + col0) * 4u;
      float hn[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int s = 2 * t + (k & 1), cc = col0 + (k >> 1) * 8;
        const uint32_t o = (uint32_t)((k & 1) * H + (k >> 1) * 8) * 4u;
        const float c = tanh_f(d[k] + lds_f(xa + (uint32_t)((k & 1) * 384 + (k >> 1) * 8) * 4u));
        const float u = lds_f(ua + o), h = lds_f(ha + o);
        const int L = s_len[s];
        const bool on = step < L;
        hn[k] = on ? u * h + (1.0f - u) * c : h;     // past the end: state copied through, output stays zero
        if (on) {
          sts_f(ha + o, hn[k]);
          const int pos = dir == 0 ? step : L - 1 - step;
          out[(int64_t)(n0 + s) * out_bs + (int64_t)pos * (2 * H) + dir * H + cc] = hn[k];
        }
      }
      store_frag(sbase + B_X, warp, g, t, hn);
    }
    __syncthreads();             // new h complete before the next step's gate phase
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

}  // namespace

size_t bigru_mma_frag_words() { return (size_t)2 * NW * 16 * 32 * 8; }   // 32-bit words of one CBHG's fragment stream

// ug [2][128][256], uc [2][128][128] fp32 (recurrent halves of the TF kernels) -> A fragments (hi, lo) of W^T tiles:
// [dir][warp][16 chunk-tiles][hi: 32 lanes x uint4 | lo: 32 lanes x uint4]; chunk-tiles 0-7 = gate tile `warp`
// (k-chunks 0-7), 8-15 = candidate tile `warp` (warps 0-7; zero for the others).
void bigru_mma_pack(const float* ug, const float* uc, uint32_t* dst) {
  auto bf = [](float f) -> uint16_t {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
  };
  auto bff = [](uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; };
  for (int dir = 0; dir < 2; ++dir)
    for (int w = 0; w < NW; ++w)
      for (int ti = 0; ti < 16; ++ti) {
        uint32_t* hi = dst + ((((size_t)dir * NW + w) * 16 + ti) * 64) * 4;
        uint32_t* lo = hi + 32 * 4;
        const bool cand = ti >= 8;
        const int kc = ti & 7;
        for (int lane = 0; lane < 32; ++lane) {
          const int g = lane >> 2, t = lane & 3;
          const int rows[4] = {g, g + 8, g, g + 8}, ks[4] = {2 * t, 2 * t, 2 * t + 8, 2 * t + 8};
          for (int j = 0; j < 4; ++j) {
            float a0 = 0.f, a1 = 0.f;
            const int k = 16 * kc + ks[j], c = 16 * w + rows[j];
            if (!cand) { a0 = ug[((size_t)dir * H + k) * 256 + c]; a1 = ug[((size_t)dir * H + k + 1) * 256 + c]; }
            else if (w < 8) { a0 = uc[((size_t)dir * H + k) * H + c]; a1 = uc[((size_t)dir * H + k + 1) * H + c]; }
            const uint16_t h0 = bf(a0), h1 = bf(a1);
            const uint16_t l0 = bf(a0 - bff(h0)), l1 = bf(a1 - bff(h1));
            hi[lane * 4 + j] = (uint32_t)h0 | ((uint32_t)h1 << 16);
            lo[lane * 4 + j] = (uint32_t)l0 | ((uint32_t)l1 << 16);
          }
        }
      }
}

cudaError_t launch_bigru_mma(const float* xproj, const void* frag, const int32_t* lengths, int N, int T, float* out,
                             int64_t out_bs, cudaStream_t st) {
  if (N <= 0 || T <= 0) return cudaSuccess;
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(bigru_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_done[dev & 63] = true;
  }
  dim3 grid((N + NS - 1) / NS, 2);
  bigru_mma_kernel<<<grid, NT, SMEM_BYTES, st>>>(xproj, reinterpret_cast<const uint4*>(frag), lengths, N, T, out, out_bs);
  return cudaGetLastError();
}

}  // namespace taco
