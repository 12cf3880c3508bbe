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


// ---------------------------------------------------------------------------------------------------------------
// Epilogue v2: per-warp shared-memory staging so that every global access is a coalesced 128-byte row segment.
//
// A warp owns 32 accumulator rows.  For each group of GW (= min(BN, 64)) accumulator columns:
//   1. bias slice and (optionally) the residual sub-tile [32 rows x GWo cols] are fetched with coalesced 16-byte
//      loads into the warp's staging buffer;
//   2. each lane pulls its row's accumulators from TMEM (tcgen05.ld 32x32b.x32 per 32 columns), applies
//      bias -> round -> act -> round (+residual) -> round exactly like epilogue_chunk, and writes the packed
//      row back into the staging buffer (row pitch padded by 16 B: conflict-free for 16-byte accesses);
//   3. the warp streams the staging buffer to global memory, 8 lanes per 128-byte row segment.
// Requires ldc % 8 == 0, ldr % 8 == 0, N % 8 == 0 (N % 16 == 0 with SwiGLU) and 16-bit output; callers fall back to
// epilogue_chunk otherwise.
constexpr int EPI_PITCH = 64 * 2 + 16;                       // bytes per staged row (64 columns + pad)
constexpr int EPI_WARP_BYTES = 32 * EPI_PITCH + 64 * 4;      // staging tile + bias slice
constexpr int EPI_SMEM_BYTES = 4 * EPI_WARP_BYTES;

__device__ __forceinline__ bool epilogue_v2_ok(const GemmKParams& p) {
  if (p.out_f32 || (p.ldc & 7) || (p.residual && (p.ldr & 7))) return false;
  if (p.swiglu) return p.act == ACT_SILU && p.N % 16 == 0;
  return p.N % 8 == 0;
}

// Compile-time activation (keeps the per-element code of the hot loop small: a runtime switch inside the unrolled
// element loop replicated erff/tanhf/expf bodies 32x and made the epilogue instruction-fetch bound, see
// profiles/r01_gemm_prefill_ncu.md).
template <int ACT>
__device__ __forceinline__ float act_ct(float x) {
  if constexpr (ACT == ACT_GELU_ERF) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
  else if constexpr (ACT == ACT_SILU) return x / (1.0f + __expf(-x));
  else if constexpr (ACT == ACT_HARDSWISH) return x * fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f);
  else if constexpr (ACT == ACT_RELU) return fmaxf(x, 0.0f);
  else if constexpr (ACT == ACT_GELU_TANH) {
    const float inner = 0.7978845608028654f * (x + 0.044715f * x * x * x);
    return 0.5f * x * (1.0f + tanhf(inner));
  } else return x;
}

// taddr: TMEM address of (lane quarter, first column of this tile's accumulator); row_fn(r) = global output row of the
// warp's r-th accumulator row (0..31) or -1 when that row is outside the problem; n_col0 = first weight row of the tile.
template <typename T, int BN, int ACT, bool SWIGLU, typename RowFn>
__device__ __noinline__ void epilogue_tile_ct(uint32_t taddr, const GemmKParams& p, uint8_t* warp_smem, int lane,
                                              RowFn row_fn, int n_col0) {
  constexpr int GW = BN < 64 ? BN : 64;            // accumulator columns per group
  constexpr int GWO = SWIGLU ? GW / 2 : GW;        // output columns per group
  constexpr int CPR = GWO / 8;                     // 16-byte chunks per staged row
  uint8_t* stage = warp_smem;
  float* sbias = reinterpret_cast<float*>(warp_smem + 32 * EPI_PITCH);
  const int n_out = SWIGLU ? (p.N >> 1) : p.N;
  T* cbase = reinterpret_cast<T*>(p.C);
  const T* rbase = SWIGLU ? nullptr : reinterpret_cast<const T*>(p.residual);
  const bool has_bias = p.bias != nullptr;
#pragma unroll 1
  for (int g = 0; g < BN / GW; ++g) {
    const int acol0 = n_col0 + g * GW;             // accumulator (weight-row) column
    if (acol0 >= p.N) break;
    const int ocol0 = SWIGLU ? (acol0 >> 1) : acol0;
    __syncwarp();
    if (has_bias) {
#pragma unroll
      for (int j = lane; j < GW; j += 32) sbias[j] = (acol0 + j < p.N) ? __ldg(p.bias + acol0 + j) : 0.f;
    }
    if (rbase) {
#pragma unroll
      for (int it = 0; it < CPR; ++it) {
        const int idx = it * 32 + lane;
        const int r = idx / CPR, ch = idx % CPR;
        const int grow = row_fn(r);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (grow >= 0 && ocol0 + ch * 8 < n_out)
          v = *reinterpret_cast<const uint4*>(rbase + static_cast<size_t>(grow) * p.ldr + ocol0 + ch * 8);
        *reinterpret_cast<uint4*>(stage + r * EPI_PITCH + ch * 16) = v;
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int c = 0; c < GW / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(taddr + g * GW + c * 32, v);
      tmem_ld_wait();
      float x[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
      if (has_bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sbias + c * 32 + j);
          x[j] += b4.x; x[j + 1] += b4.y; x[j + 2] += b4.z; x[j + 3] += b4.w;
        }
      }
      if constexpr (SWIGLU) {
        T o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float gt = rnd<T>(x[2 * i]);
          const float up = rnd<T>(x[2 * i + 1]);
          o[i] = from_f<T>(rnd<T>(act_ct<ACT>(gt)) * up);
        }
        uint4* dst = reinterpret_cast<uint4*>(stage + lane * EPI_PITCH + c * 32);
        dst[0] = reinterpret_cast<const uint4*>(o)[0];
        dst[1] = reinterpret_cast<const uint4*>(o)[1];
      } else {
        T o[32];
        uint4* srow = reinterpret_cast<uint4*>(stage + lane * EPI_PITCH + c * 64);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float y = rnd<T>(x[j]);
          if constexpr (ACT != ACT_NONE) y = rnd<T>(act_ct<ACT>(y));
          x[j] = y;
        }
        if (rbase) {
          uint4 rv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) rv[i] = srow[i];
          const T* r = reinterpret_cast<const T*>(rv);
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] += to_f<T>(r[j]);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = from_f<T>(x[j]);
#pragma unroll
        for (int i = 0; i < 4; ++i) srow[i] = reinterpret_cast<const uint4*>(o)[i];
      }
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < CPR; ++it) {
      const int idx = it * 32 + lane;
      const int r = idx / CPR, ch = idx % CPR;
      const int grow = row_fn(r);
      if (grow >= 0 && ocol0 + ch * 8 < n_out)
        *reinterpret_cast<uint4*>(cbase + static_cast<size_t>(grow) * p.ldc + ocol0 + ch * 8) =
            *reinterpret_cast<const uint4*>(stage + r * EPI_PITCH + ch * 16);
    }
  }
}

// Runtime -> compile-time dispatch on (act, swiglu); done once per tile, outside every loop.
template <typename T, int BN, typename RowFn>
__device__ __forceinline__ void epilogue_tile_v2(uint32_t taddr, const GemmKParams& p, uint8_t* warp_smem, int lane,
                                                 RowFn row_fn, int n_col0) {
  if (p.swiglu) {
    epilogue_tile_ct<T, BN, ACT_SILU, true>(taddr, p, warp_smem, lane, row_fn, n_col0);
    return;
  }
  switch (p.act) {
    case ACT_NONE: epilogue_tile_ct<T, BN, ACT_NONE, false>(taddr, p, warp_smem, lane, row_fn, n_col0); break;
    case ACT_GELU_ERF: epilogue_tile_ct<T, BN, ACT_GELU_ERF, false>(taddr, p, warp_smem, lane, row_fn, n_col0); break;
    case ACT_HARDSWISH: epilogue_tile_ct<T, BN, ACT_HARDSWISH, false>(taddr, p, warp_smem, lane, row_fn, n_col0); break;
    case ACT_RELU: epilogue_tile_ct<T, BN, ACT_RELU, false>(taddr, p, warp_smem, lane, row_fn, n_col0); break;
    case ACT_SILU: epilogue_tile_ct<T, BN, ACT_SILU, false>(taddr, p, warp_smem, lane, row_fn, n_col0); break;
    default: epilogue_tile_ct<T, BN, ACT_GELU_TANH, false>(taddr, p, warp_smem, lane, row_fn, n_col0); break;
  }
}

}  // namespace sb
