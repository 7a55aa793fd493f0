// K1: text-embedding gather fused with the speaker-embedding broadcast+concat.
// Replaces tf.nn.embedding_lookup x2 + expand_dims + tile + concat at reference
// models/tacotron.py:46-55.  HBM-bound: one float4 per thread, rows of the
// output are written fully coalesced.
#include "common.cuh"
#include "kernels.cuh"

namespace taco {

__global__ void __launch_bounds__(256)
gather_concat_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ spk,
                     const float* __restrict__ table, int V, int E,
                     const float* __restrict__ spk_table, int S, int Es,
                     int rows, int T, float* __restrict__ out, int* __restrict__ oob_flag) {
  const int W4 = (E + Es) >> 2;            // float4 per output row
  const int E4 = E >> 2;
  const int64_t total = (int64_t)rows * W4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / W4);
    const int q = (int)(i - (int64_t)row * W4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < E4) {
      const int id = __ldg(ids + row);
      if (id >= 0 && id < V) v = ldg_f4(table + (int64_t)id * E + q * 4);
      else if (q == 0) atomicOr(oob_flag, 1);
    } else {
      const int s = __ldg(spk + row / T);
      if (s >= 0 && s < S) v = ldg_f4(spk_table + (int64_t)s * Es + (q - E4) * 4);
      else if (q == E4) atomicOr(oob_flag, 2);
    }
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

void launch_gather_concat(const int32_t* ids, const int32_t* spk, const float* table, int V, int E,
                          const float* spk_table, int S, int Es, int N, int T, float* out,
                          int* oob_flag, cudaStream_t st) {
  if (spk == nullptr) Es = 0;
  const int rows = N * T;
  const int64_t total = (int64_t)rows * ((E + Es) / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  gather_concat_kernel<<<blocks, 256, 0, st>>>(ids, spk, table, V, E, spk_table, S, Es, rows, T, out,
                                               oob_flag);
}

}  // namespace taco
