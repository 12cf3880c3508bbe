// surya_b200 — epilogue shared by the tcgen05 GEMM and implicit-GEMM convolution kernels.
// One call handles the 32 accumulator columns a thread just pulled from TMEM for its output row.
#pragma once
#include "gemm.cuh"
#include "sb_ptx.cuh"

namespace sb {

struct GemmKParams {
  int M, N, K;
  void* C;
  int ldc;
  const float* bias;
  const void* residual;
  int ldr;
  int act;
  int swiglu;
  int out_f32;
  int group_m;
  int group_k;  // grouped 1x1 conv: A column offset per n-block (0 = dense)
};

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_GELU_ERF: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
    case ACT_SILU: return x / (1.0f + expf(-x));
    case ACT_HARDSWISH: return x * fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) / 6.0f;
    case ACT_RELU: return fmaxf(x, 0.0f);
    case ACT_GELU_TANH: {
      const float k0 = 0.7978845608028654f, k1 = 0.044715f;
      float inner = k0 * (x + k1 * x * x * x);
      return 0.5f * x * (1.0f + tanhf(inner));
    }
    default: return x;
  }
}


// x = acc (+bias) -> round -> act -> round (+residual) -> round -> store; `row` is the global output row
// (already translated to an NHWC pixel index by the convolution kernel), `col0` the first of 32 columns.
template <typename T>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], const GemmKParams& p, int row, bool row_ok,
                                               int col0, bool vec_ok) {
  float x[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (col0 + j < p.N) x[j] += __ldg(p.bias + col0 + j);
    }
  }
  if (!row_ok) return;
  if (p.swiglu) {
    // columns (2i, 2i+1) = (gate_i, up_i)
    const int oc0 = col0 >> 1;
    const int n_out = p.N >> 1;
    T* crow = reinterpret_cast<T*>(p.C) + static_cast<size_t>(row) * p.ldc;
    T o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float g = rnd<T>(x[2 * i]);
      float u = rnd<T>(x[2 * i + 1]);
      float sact = rnd<T>(apply_act(g, p.act));
      o[i] = from_f<T>(sact * u);
    }
    if (vec_ok && oc0 + 16 <= n_out) {
      uint4* dst = reinterpret_cast<uint4*>(crow + oc0);
      const uint4* src = reinterpret_cast<const uint4*>(o);
      dst[0] = src[0];
      dst[1] = src[1];
    } else {
      for (int i = 0; i < 16; ++i)
        if (oc0 + i < n_out) crow[oc0 + i] = o[i];
    }
  } else if (p.out_f32) {
    float* crow = reinterpret_cast<float*>(p.C) + static_cast<size_t>(row) * p.ldc;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float y = apply_act(x[j], p.act);
      if (col0 + j < p.N) crow[col0 + j] = y;
    }
  } else {
    T* crow = reinterpret_cast<T*>(p.C) + static_cast<size_t>(row) * p.ldc;
    const T* rrow = p.residual ? reinterpret_cast<const T*>(p.residual) + static_cast<size_t>(row) * p.ldr : nullptr;
    const bool full = (col0 + 32 <= p.N) && vec_ok;
    T r[32];
    if (rrow) {
      if (full) {
        const uint4* src = reinterpret_cast<const uint4*>(rrow + col0);
        uint4* dst = reinterpret_cast<uint4*>(r);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = src[i];
      } else {
        for (int j = 0; j < 32; ++j) r[j] = (col0 + j < p.N) ? rrow[col0 + j] : from_f<T>(0.f);
      }
    }
    T o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float y = rnd<T>(x[j]);
      if (p.act != ACT_NONE) y = rnd<T>(apply_act(y, p.act));
      if (rrow) y = y + to_f<T>(r[j]);
      o[j] = from_f<T>(y);
    }
    if (full) {
      uint4* dst = reinterpret_cast<uint4*>(crow + col0);
      const uint4* src = reinterpret_cast<const uint4*>(o);
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = src[i];
    } else {
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) crow[col0 + j] = o[j];
    }
  }
}

}  // namespace sb
