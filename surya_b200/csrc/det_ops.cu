// surya_b200 — CUDA-core kernels of the detection path (NHWC fp16/bf16 activations; all HBM-bound).
//
//   stem_conv      Stem.in_conv: 3x3 s2 conv 3->32 + folded BN + Hardswish from the NCHW input
//                  (surya/detection/model/encoderdecoder.py:484-495)
//   dwconv         depthwise k x k conv (+bias, +Hardswish): MBConv.depth_conv (:202-211) and LiteMLA.aggreg[0] (:313-320)
//   lite_mla       ReLU linear attention core, fp32: kv = relu(k)^T [v,1]; out = relu(q) kv; out[:-1]/(out[-1]+eps)
//                  (:332-338, 352-359)
//   upsample_cat   DecodeHead: bilinear(align_corners=False) upsample of the 4 projected maps to the stage-0 size and
//                  channel concat in reversed order (:703-716)
//   classifier     1x1 conv 512->2 + bias + sigmoid, NCHW output (:720, :747)
//   upsample_nchw  DetectionPredictor's x4 bilinear upsample to fp32 (surya/detection/__init__.py:120-132)
#include "ops.cuh"
#include "sb_ptx.cuh"

namespace sb {

__device__ __forceinline__ float hardswish_f(float x) { return x * fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) / 6.0f; }

// ------------------------------------------------------------------------------------------------ stem conv
// in: NCHW [B,3,H,W] (T or float); w: fp32 [32][r][s][c] (27 per output channel, BN folded); out NHWC [B,H/2,W/2,CO].
template <typename T, typename InT, int CO>
__global__ void __launch_bounds__(128) stem_conv_kernel(const InT* __restrict__ in, const float* __restrict__ w,
                                                        const float* __restrict__ bias, T* __restrict__ out, int B,
                                                        int H, int W) {
  __shared__ float sw[CO * 27];
  __shared__ float sb_[CO];
  for (int i = threadIdx.x; i < CO * 27; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < CO; i += blockDim.x) sb_[i] = bias[i];
  __syncthreads();
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = static_cast<long long>(B) * Ho * Wo;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = idx % Wo;
    const int oy = (idx / Wo) % Ho;
    const int b = idx / (static_cast<long long>(Wo) * Ho);
    float x[27];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = oy * 2 + r - 1;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int ix = ox * 2 + s - 1;
        const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float v = 0.f;
          if (ok) {
            const InT* p = in + ((static_cast<size_t>(b) * 3 + c) * H + iy) * W + ix;
            if constexpr (sizeof(InT) == 4) v = rnd<T>(static_cast<float>(*p));  // .to(model dtype) of the pixel batch
            else v = to_f<InT>(*p);
          }
          x[(r * 3 + s) * 3 + c] = v;
        }
      }
    }
    T* o = out + idx * CO;
#pragma unroll
    for (int co8 = 0; co8 < CO; co8 += 8) {
      uint4 pack;
      T* pe = reinterpret_cast<T*>(&pack);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float* wr = sw + (co8 + j) * 27;
        float acc = sb_[co8 + j];
#pragma unroll
        for (int k = 0; k < 27; ++k) acc += x[k] * wr[k];
        pe[j] = from_f<T>(hardswish_f(rnd<T>(acc)));
      }
      *reinterpret_cast<uint4*>(o + co8) = pack;
    }
  }
}

int det_stem_conv(int dtype, const void* in, int in_f32, const float* w, const float* bias, void* out, int B, int H,
                  int W, int cout, cudaStream_t st) {
  if (cout != 32) { set_error("det_stem_conv: only 32 output channels are instantiated (got %d)", cout); return -1; }
  if ((H | W) & 1) { set_error("det_stem_conv: H and W must be even"); return -1; }
  const long long total = static_cast<long long>(B) * (H / 2) * (W / 2);
  int grid = static_cast<int>((total + 127) / 128);
  if (grid > num_sms() * 32) grid = num_sms() * 32;
  if (dtype == DT_F16) {
    if (in_f32) stem_conv_kernel<__half, float, 32><<<grid, 128, 0, st>>>((const float*)in, w, bias, (__half*)out, B, H, W);
    else stem_conv_kernel<__half, __half, 32><<<grid, 128, 0, st>>>((const __half*)in, w, bias, (__half*)out, B, H, W);
  } else {
    if (in_f32) stem_conv_kernel<__nv_bfloat16, float, 32><<<grid, 128, 0, st>>>((const float*)in, w, bias, (__nv_bfloat16*)out, B, H, W);
    else stem_conv_kernel<__nv_bfloat16, __nv_bfloat16, 32><<<grid, 128, 0, st>>>((const __nv_bfloat16*)in, w, bias, (__nv_bfloat16*)out, B, H, W);
  }
  return launch_ok();
}

// ------------------------------------------------------------------------------------------------ depthwise conv
// in NHWC [B,H,W,C]; w: T [k*k][C] (tap-major); bias fp32 [C] or null; one thread = 8 channels of one output pixel.
template <typename T, int KS>
__global__ void __launch_bounds__(256) dwconv_kernel(const T* __restrict__ in, const T* __restrict__ w,
                                                     const float* __restrict__ bias, T* __restrict__ out, int B, int H,
                                                     int W, int C, int stride, int pad, int act) {
  const int Ho = (H + 2 * pad - KS) / stride + 1, Wo = (W + 2 * pad - KS) / stride + 1;
  const int cv = C >> 3;
  const long long total = static_cast<long long>(B) * Ho * Wo * cv;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = (idx % cv) * 8;
    const long long pix = idx / cv;
    const int ox = pix % Wo;
    const int oy = (pix / Wo) % Ho;
    const int b = pix / (static_cast<long long>(Wo) * Ho);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int r = 0; r < KS; ++r) {
      const int iy = oy * stride + r - pad;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        const int ix = ox * stride + s - pad;
        if (ix < 0 || ix >= W) continue;
        const uint4 xv = *reinterpret_cast<const uint4*>(in + ((static_cast<size_t>(b) * H + iy) * W + ix) * C + c8);
        const uint4 wv = *reinterpret_cast<const uint4*>(w + static_cast<size_t>(r * KS + s) * C + c8);
        const T* xe = reinterpret_cast<const T*>(&xv);
        const T* we = reinterpret_cast<const T*>(&wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += to_f<T>(xe[j]) * to_f<T>(we[j]);
      }
    }
    uint4 pack;
    T* pe = reinterpret_cast<T*>(&pack);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float y = rnd<T>(acc[j] + (bias ? bias[c8 + j] : 0.f));
      if (act == ACT_HARDSWISH) y = hardswish_f(y);
      pe[j] = from_f<T>(y);
    }
    *reinterpret_cast<uint4*>(out + pix * C + c8) = pack;
  }
}

int det_dwconv(int dtype, const void* in, const void* w, const float* bias, void* out, int B, int H, int W, int C, int ks,
               int stride, int pad, int act, cudaStream_t st) {
  if (C % 8) { set_error("det_dwconv: C must be a multiple of 8"); return -1; }
  const int Ho = (H + 2 * pad - ks) / stride + 1, Wo = (W + 2 * pad - ks) / stride + 1;
  const long long total = static_cast<long long>(B) * Ho * Wo * (C / 8);
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > num_sms() * 16) grid = num_sms() * 16;
#define DW(T_, KS_) dwconv_kernel<T_, KS_><<<grid, 256, 0, st>>>((const T_*)in, (const T_*)w, bias, (T_*)out, B, H, W, C, stride, pad, act)
  if (dtype == DT_F16) {
    if (ks == 3) DW(__half, 3); else if (ks == 5) DW(__half, 5); else { set_error("det_dwconv: kernel size 3 or 5"); return -1; }
  } else {
    if (ks == 3) DW(__nv_bfloat16, 3); else if (ks == 5) DW(__nv_bfloat16, 5); else { set_error("det_dwconv: kernel size 3 or 5"); return -1; }
  }
#undef DW
  return launch_ok();
}

// ------------------------------------------------------------------------------------------------ LiteMLA core
// qkv_a / qkv_b: token-major [B*HW, 3*heads*DIM] (plain qkv and aggregated qkv); head h of scale s reads channels
// [h*3*DIM, (h+1)*3*DIM) = (q | k | v).  out: [B*HW, 2*heads*DIM], channel = (s*heads + h)*DIM + d.
// One CTA per (image, scale*heads + h).  fp32 throughout; rounded to T once at the end (reference: .float() ... .to(dtype)).
template <typename T, int DIM>
__global__ void __launch_bounds__(256) lite_mla_kernel(const T* __restrict__ qkv_a, const T* __restrict__ qkv_b,
                                                       T* __restrict__ out, int HW, int heads, float eps) {
  constexpr int DV = DIM + 1;
  __shared__ float kv[DIM * DV];
  __shared__ float tk[64][DIM + 1];
  __shared__ float tv[64][DIM + 1];
  const int b = blockIdx.x, hh = blockIdx.y;
  const int scale = hh / heads, h = hh % heads;
  const T* src = (scale == 0 ? qkv_a : qkv_b) + static_cast<size_t>(b) * HW * (3 * heads * DIM) + h * 3 * DIM;
  const int ld = 3 * heads * DIM;
  const int tid = threadIdx.x;
  // each thread owns up to ceil(DIM*DV/256) entries of kv
  float acc[(DIM * DV + 255) / 256];
#pragma unroll
  for (int i = 0; i < (DIM * DV + 255) / 256; ++i) acc[i] = 0.f;
  for (int t0 = 0; t0 < HW; t0 += 64) {
    __syncthreads();
    for (int i = tid; i < 64 * DIM; i += 256) {
      const int t = i / DIM, d = i % DIM;
      float kx = 0.f, vx = 0.f;
      if (t0 + t < HW) {
        const T* row = src + static_cast<size_t>(t0 + t) * ld;
        kx = fmaxf(to_f<T>(row[DIM + d]), 0.f);
        vx = to_f<T>(row[2 * DIM + d]);
      }
      tk[t][d] = kx;
      tv[t][d] = vx;
    }
    __syncthreads();
    const int nt = min(64, HW - t0);
#pragma unroll
    for (int i = 0; i < (DIM * DV + 255) / 256; ++i) {
      const int e = tid + i * 256;
      if (e < DIM * DV) {
        const int ki = e / DV, vj = e % DV;
        float a = acc[i];
        if (vj < DIM) {
          for (int t = 0; t < nt; ++t) a += tk[t][ki] * tv[t][vj];
        } else {
          for (int t = 0; t < nt; ++t) a += tk[t][ki];  // ones column
        }
        acc[i] = a;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < (DIM * DV + 255) / 256; ++i) {
    const int e = tid + i * 256;
    if (e < DIM * DV) kv[e] = acc[i];
  }
  __syncthreads();
  T* dst = out + static_cast<size_t>(b) * HW * (2 * heads * DIM) + (scale * heads + h) * DIM;
  const int ldo = 2 * heads * DIM;
  for (int t = tid; t < HW; t += 256) {
    const T* row = src + static_cast<size_t>(t) * ld;
    float q[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d += 8) {
      uint4 u = *reinterpret_cast<const uint4*>(row + d);
      const T* e = reinterpret_cast<const T*>(&u);
#pragma unroll
      for (int j = 0; j < 8; ++j) q[d + j] = fmaxf(to_f<T>(e[j]), 0.f);
    }
    float o[DV];
#pragma unroll
    for (int j = 0; j < DV; ++j) o[j] = 0.f;
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
#pragma unroll
      for (int j = 0; j < DV; ++j) o[j] += q[i] * kv[i * DV + j];
    }
    const float den = o[DIM] + eps;
#pragma unroll
    for (int d = 0; d < DIM; d += 8) {
      uint4 pack;
      T* pe = reinterpret_cast<T*>(&pack);
#pragma unroll
      for (int j = 0; j < 8; ++j) pe[j] = from_f<T>(o[d + j] / den);
      *reinterpret_cast<uint4*>(dst + static_cast<size_t>(t) * ldo + d) = pack;
    }
  }
}

int det_lite_mla(int dtype, const void* qkv_a, const void* qkv_b, void* out, int B, int HW, int heads, int dim, float eps,
                 cudaStream_t st) {
  if (dim != 32) { set_error("det_lite_mla: head dim 32 is instantiated (got %d)", dim); return -1; }
  dim3 grid(B, 2 * heads), block(256);
  if (dtype == DT_F16)
    lite_mla_kernel<__half, 32><<<grid, block, 0, st>>>((const __half*)qkv_a, (const __half*)qkv_b, (__half*)out, HW, heads, eps);
  else
    lite_mla_kernel<__nv_bfloat16, 32><<<grid, block, 0, st>>>((const __nv_bfloat16*)qkv_a, (const __nv_bfloat16*)qkv_b,
                                                               (__nv_bfloat16*)out, HW, heads, eps);
  return launch_ok();
}

// ------------------------------------------------------------------------------------------------ decode-head upsample + concat
// src maps NHWC [B, hs, ws, CS]; dst NHWC [B, HO, WO, n_src*CS], map i written at channel offset ch_off[i].
struct UpcatParams {
  const void* src[4];
  int hs[4], ws[4], ch_off[4];
  int n_src, CS, HO, WO, B;
};

template <typename T>
__global__ void __launch_bounds__(256) upsample_cat_kernel(const UpcatParams p, T* __restrict__ dst) {
  const int cv = p.CS >> 3;
  const int CT = p.n_src * p.CS;
  const long long total = static_cast<long long>(p.B) * p.HO * p.WO * p.n_src * cv;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = (idx % cv) * 8;
    long long r = idx / cv;
    const int si = r % p.n_src;
    r /= p.n_src;
    const int ox = r % p.WO;
    const int oy = (r / p.WO) % p.HO;
    const int b = r / (static_cast<long long>(p.WO) * p.HO);
    const T* src = reinterpret_cast<const T*>(p.src[si]);
    const int hs = p.hs[si], ws = p.ws[si];
    uint4 pack;
    if (hs == p.HO && ws == p.WO) {
      pack = *reinterpret_cast<const uint4*>(src + ((static_cast<size_t>(b) * hs + oy) * ws + ox) * p.CS + c8);
    } else {
      // torch upsample_bilinear2d, align_corners=False: src = (dst + 0.5) * (in/out) - 0.5, clamped at 0
      const float sy = fmaxf((oy + 0.5f) * (static_cast<float>(hs) / p.HO) - 0.5f, 0.f);
      const float sx = fmaxf((ox + 0.5f) * (static_cast<float>(ws) / p.WO) - 0.5f, 0.f);
      const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
      const int y1 = y0 + (y0 < hs - 1 ? 1 : 0), x1 = x0 + (x0 < ws - 1 ? 1 : 0);
      const float ly = sy - y0, lx = sx - x0, hy = 1.f - ly, hx = 1.f - lx;
      const size_t base = static_cast<size_t>(b) * hs * ws;
      const uint4 v00 = *reinterpret_cast<const uint4*>(src + (base + static_cast<size_t>(y0) * ws + x0) * p.CS + c8);
      const uint4 v01 = *reinterpret_cast<const uint4*>(src + (base + static_cast<size_t>(y0) * ws + x1) * p.CS + c8);
      const uint4 v10 = *reinterpret_cast<const uint4*>(src + (base + static_cast<size_t>(y1) * ws + x0) * p.CS + c8);
      const uint4 v11 = *reinterpret_cast<const uint4*>(src + (base + static_cast<size_t>(y1) * ws + x1) * p.CS + c8);
      const T *e00 = reinterpret_cast<const T*>(&v00), *e01 = reinterpret_cast<const T*>(&v01);
      const T *e10 = reinterpret_cast<const T*>(&v10), *e11 = reinterpret_cast<const T*>(&v11);
      T* pe = reinterpret_cast<T*>(&pack);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = hy * (hx * to_f<T>(e00[j]) + lx * to_f<T>(e01[j])) + ly * (hx * to_f<T>(e10[j]) + lx * to_f<T>(e11[j]));
        pe[j] = from_f<T>(v);
      }
    }
    *reinterpret_cast<uint4*>(dst + ((static_cast<size_t>(b) * p.HO + oy) * p.WO + ox) * CT + p.ch_off[si] + c8) = pack;
  }
}

int det_upsample_cat(int dtype, const void* const* src, const int* hs, const int* ws, const int* ch_off, int n_src, int CS,
                     void* dst, int B, int HO, int WO, cudaStream_t st) {
  if (n_src > 4 || CS % 8) { set_error("det_upsample_cat: at most 4 sources, CS multiple of 8"); return -1; }
  UpcatParams p;
  for (int i = 0; i < n_src; ++i) { p.src[i] = src[i]; p.hs[i] = hs[i]; p.ws[i] = ws[i]; p.ch_off[i] = ch_off[i]; }
  p.n_src = n_src; p.CS = CS; p.HO = HO; p.WO = WO; p.B = B;
  const long long total = static_cast<long long>(B) * HO * WO * n_src * (CS / 8);
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > num_sms() * 16) grid = num_sms() * 16;
  if (dtype == DT_F16) upsample_cat_kernel<__half><<<grid, 256, 0, st>>>(p, (__half*)dst);
  else upsample_cat_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p, (__nv_bfloat16*)dst);
  return launch_ok();
}

// ------------------------------------------------------------------------------------------------ classifier + sigmoid
// x NHWC [P, C]; w T [n_out][C]; b T [n_out]; out NCHW [B, n_out, HW] T: sigmoid(T(x.w + b)); one warp per pixel.
template <typename T, int NOUT>
__global__ void __launch_bounds__(256) classifier_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                         const T* __restrict__ b, T* __restrict__ out, long long P,
                                                         int C, int HW) {
  const long long pix = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pix >= P) return;
  const T* xr = x + pix * C;
  float acc[NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; ++o) acc[o] = 0.f;
  for (int c = lane * 8; c < C; c += 256) {
    const uint4 xv = *reinterpret_cast<const uint4*>(xr + c);
    const T* xe = reinterpret_cast<const T*>(&xv);
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
      const uint4 wv = *reinterpret_cast<const uint4*>(w + static_cast<size_t>(o) * C + c);
      const T* we = reinterpret_cast<const T*>(&wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[o] += to_f<T>(xe[j]) * to_f<T>(we[j]);
    }
  }
#pragma unroll
  for (int o = 0; o < NOUT; ++o) acc[o] = warp_sum(acc[o]);
  if (lane == 0) {
    const long long bimg = pix / HW, hw = pix % HW;
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
      float v = rnd<T>(acc[o] + to_f<T>(b[o]));
      out[(bimg * NOUT + o) * HW + hw] = from_f<T>(1.f / (1.f + expf(-v)));
    }
  }
}

int det_classifier(int dtype, const void* x, const void* w, const void* b, void* out, long long P, int C, int HW, int n_out,
                   cudaStream_t st) {
  if (n_out != 2 || C % 8) { set_error("det_classifier: 2 labels and C %% 8 == 0 are instantiated"); return -1; }
  const int grid = static_cast<int>((P + 7) / 8);
  if (dtype == DT_F16)
    classifier_kernel<__half, 2><<<grid, 256, 0, st>>>((const __half*)x, (const __half*)w, (const __half*)b, (__half*)out, P, C, HW);
  else
    classifier_kernel<__nv_bfloat16, 2><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w,
                                                              (const __nv_bfloat16*)b, (__nv_bfloat16*)out, P, C, HW);
  return launch_ok();
}

// ------------------------------------------------------------------------------------------------ x4 bilinear to fp32 (NCHW)
template <typename T>
__global__ void __launch_bounds__(256) upsample_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int planes,
                                                            int hs, int ws, int HO, int WO) {
  const long long total = static_cast<long long>(planes) * HO * WO;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = idx % WO;
    const int oy = (idx / WO) % HO;
    const int pl = idx / (static_cast<long long>(WO) * HO);
    const float sy = fmaxf((oy + 0.5f) * (static_cast<float>(hs) / HO) - 0.5f, 0.f);
    const float sx = fmaxf((ox + 0.5f) * (static_cast<float>(ws) / WO) - 0.5f, 0.f);
    const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
    const int y1 = y0 + (y0 < hs - 1 ? 1 : 0), x1 = x0 + (x0 < ws - 1 ? 1 : 0);
    const float ly = sy - y0, lx = sx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const T* p = in + static_cast<size_t>(pl) * hs * ws;
    float v = hy * (hx * to_f<T>(p[y0 * ws + x0]) + lx * to_f<T>(p[y0 * ws + x1])) +
              ly * (hx * to_f<T>(p[y1 * ws + x0]) + lx * to_f<T>(p[y1 * ws + x1]));
    out[idx] = rnd<T>(v);  // F.interpolate runs in the model dtype; .float() afterwards
  }
}

int det_upsample_nchw(int dtype, const void* in, float* out, int planes, int hs, int ws, int HO, int WO, cudaStream_t st) {
  const long long total = static_cast<long long>(planes) * HO * WO;
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > num_sms() * 16) grid = num_sms() * 16;
  if (dtype == DT_F16) upsample_nchw_kernel<__half><<<grid, 256, 0, st>>>((const __half*)in, out, planes, hs, ws, HO, WO);
  else upsample_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, out, planes, hs, ws, HO, WO);
  return launch_ok();
}

}  // namespace sb
