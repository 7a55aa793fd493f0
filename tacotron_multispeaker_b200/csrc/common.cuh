// Shared device helpers for the sm_100a kernels of the Tacotron forward path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace taco {

// ---- activations ---------------------------------------------------------
// exp-based forms on the SFU (ex2.approx): abs error ~1e-6, far inside the
// 1e-3 parity budget, and ~5 instructions instead of ~25 for tanhf().
// Written on the raw SFU instructions: __expf / __fdividef wrap them in range fix-ups (FSETP + two predicated FMULs around
// EX2, a magnitude test around RCP) that these forms do not need -- 2^(+-big) saturates to inf / 0 and 1 / inf = 0, which is
// exactly the limit wanted -- and that sat on the critical chain of every gate phase.
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_approx1(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_approx1(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_f(float x) {
  // 1 - 2/(e^{2x}+1); saturates cleanly for |x| large (inf -> 1, 0 -> -1).
  return fmaf(-2.0f, rcp_approx1(ex2_approx(2.8853900817779268f * x) + 1.0f), 1.0f);
}
__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case 1: return fmaxf(v, 0.0f);
    case 2: return sigmoid_f(v);
    case 3: return tanh_f(v);
    default: return v;
  }
}

// ---- packed fp32 FMA (Blackwell FFMA2: two fp32 FMAs per issue slot) ---------
// (d0, d1) += (a0, a1) * (b0, b1)
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n.reg .b64 ra, rb, rc;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rc, {%0, %1};\n"
      "fma.rn.f32x2 rc, ra, rb, rc;\nmov.b64 {%0, %1}, rc;\n}"
      : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// ---- vector loads --------------------------------------------------------
__device__ __forceinline__ float4 ldg_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
// Streaming 128-bit load that does not allocate in L1 (weights read once per step).
__device__ __forceinline__ float4 ldg_f4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- thread-block cluster primitives (PTX; sm_90+) ------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r;
}
// Split-phase cluster barrier with release/acquire semantics at cluster scope:
// distributed-shared-memory stores issued before arrive are visible to every
// CTA of the cluster after wait.
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() { cluster_arrive(); cluster_wait(); }

// Map a local shared-memory address to the same offset in CTA `rank` of the cluster.
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace taco
