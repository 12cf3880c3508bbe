// surya_b200 — the one kernel the ocr_error path (DistilBertForSequenceClassification) needs beyond the shared
// GEMM / attn_varlen / layernorm / small_head kernels.
//
//   embed_pos_layernorm   Embeddings.forward (surya/ocr_error/model/encoder.py:60-91): word_embeddings[id] + position_embeddings[pos]
//                         (one rounding of the sum, as the 16-bit tensor add of the reference), LayerNorm(eps) with fp32 statistics.
//                         Rows are the PACKED real tokens of a right-padded batch (pad positions are never computed), so ids / pos
//                         come as explicit int32 arrays built by the host plan.
#include "ops.cuh"
#include "sb_ptx.cuh"

namespace sb {

constexpr int EPL_MAXV = 8;   // 16-byte vectors per lane: C <= 8 * 32 * 8 = 2048

template <typename T>
__global__ void __launch_bounds__(128) embed_pos_layernorm_kernel(const int* __restrict__ ids, const int* __restrict__ pos,
                                                                  const T* __restrict__ word, const T* __restrict__ ptab,
                                                                  const T* __restrict__ w, const T* __restrict__ b,
                                                                  T* __restrict__ y, int rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;                                   // whole warps leave together
  const T* wr = word + static_cast<size_t>(ids[row]) * C;
  const T* pr = ptab + static_cast<size_t>(pos[row]) * C;
  uint4 v[EPL_MAXV];
  float s = 0.f;
#pragma unroll
  for (int it = 0; it < EPL_MAXV; ++it) {
    const int i = (it * 32 + lane) * 8;
    v[it] = make_uint4(0u, 0u, 0u, 0u);
    if (i < C) {
      const uint4 a = *reinterpret_cast<const uint4*>(wr + i);
      const uint4 p = *reinterpret_cast<const uint4*>(pr + i);
      const T *ae = reinterpret_cast<const T*>(&a), *pe = reinterpret_cast<const T*>(&p);
      T* e = reinterpret_cast<T*>(&v[it]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        e[j] = from_f<T>(to_f<T>(ae[j]) + to_f<T>(pe[j]));   // input_embeds + position_embeddings in the storage type (:88)
        s += to_f<T>(e[j]);
      }
    }
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int it = 0; it < EPL_MAXV; ++it) {
    const int i = (it * 32 + lane) * 8;
    if (i < C) {
      const T* e = reinterpret_cast<const T*>(&v[it]);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = to_f<T>(e[j]) - mean; q += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  T* yr = y + static_cast<size_t>(row) * C;
#pragma unroll
  for (int it = 0; it < EPL_MAXV; ++it) {
    const int i = (it * 32 + lane) * 8;
    if (i < C) {
      const uint4 wu = *reinterpret_cast<const uint4*>(w + i);
      const uint4 bu = *reinterpret_cast<const uint4*>(b + i);
      const T *e = reinterpret_cast<const T*>(&v[it]), *we = reinterpret_cast<const T*>(&wu), *be = reinterpret_cast<const T*>(&bu);
      uint4 o;
      T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
      for (int j = 0; j < 8; ++j) oe[j] = from_f<T>((to_f<T>(e[j]) - mean) * rstd * to_f<T>(we[j]) + to_f<T>(be[j]));
      *reinterpret_cast<uint4*>(yr + i) = o;
    }
  }
}

int embed_pos_layernorm(int dtype, const int* ids, const int* pos, const void* word, const void* ptab, const void* w, const void* b,
                        void* y, int rows, int C, float eps, cudaStream_t st) {
  if (rows <= 0) return 0;
  if (C % 8 || C > EPL_MAXV * 256) { set_error("embed_pos_layernorm: C = %d must be a multiple of 8 and <= %d", C, EPL_MAXV * 256); return -1; }
  if (!ids || !pos || !word || !ptab || !w || !b || !y) { set_error("embed_pos_layernorm: null argument"); return -2; }
  dim3 grid(static_cast<unsigned>((rows + 3) / 4)), block(128);
  if (dtype == DT_F16)
    embed_pos_layernorm_kernel<__half><<<grid, block, 0, st>>>(ids, pos, (const __half*)word, (const __half*)ptab, (const __half*)w,
                                                                (const __half*)b, (__half*)y, rows, C, eps);
  else
    embed_pos_layernorm_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(ids, pos, (const __nv_bfloat16*)word, (const __nv_bfloat16*)ptab,
                                                                       (const __nv_bfloat16*)w, (const __nv_bfloat16*)b,
                                                                       (__nv_bfloat16*)y, rows, C, eps);
  return launch_ok();
}

}  // namespace sb
