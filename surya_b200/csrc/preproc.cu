// surya_b200 — recognition crop preprocessing on the device: uint8 line crops -> fp32 image tiles of sb_rec_prefill.
//
// Replaces, per crop, the host chain of the reference (all OpenCV on float32 images):
//   SuryaOCRProcessor.scale_to_fit     cv2.resize(INTER_LANCZOS4) when the pixel count is outside [168*168, 1024*256]
//                                      (surya/common/surya/processor/__init__.py:140-178)
//   _process_and_tile                  cv2.resize(INTER_CUBIC) to the next multiple of patch*merge = 28, _image_processor
//                                      (x * (1/255) in double, (x - mean) / std in float), merge-block-major patch tiles
//                                      (processor/__init__.py:180-230)
// The host keeps what is integer / polygon work: slicing the crop out of the page and masking outside the polygon
// (surya/input/processing.py:57-101), and the output sizes (the same Python float arithmetic as the reference).  It uploads the crops
// as uint8 (3 B / pixel instead of 12 B / patch element of fp32 tiles).
//
// OpenCV is a third-party dependency of the reference (opencv-python 4.x; 4.13.0 in this image); its resize is restated here:
//   coordinate of destination index d on an axis:  f = (d + 0.5) * scale - 0.5,  scale = 1 / (dst / src)  (double), s = floor(f)
//   taps s - k/2 + 1 ... s + k/2 with the index clamped to the image (border replicate), k = 4 (cubic) / 8 (Lanczos)
//   INTER_LANCZOS4 (generic path, float coordinate): x = float(f) - s; interpolateLanczos4: weights cs[i] . (sin y0, cos y0) / y_i^2,
//       y_i = -(x + 3 - i) * pi / 4, normalised by their float sum; x < FLT_EPSILON -> the centre tap alone
//   INTER_CUBIC  (the IPP path this image's OpenCV takes for float32: coordinate and Keys weights, A = -0.75, in double)
//   two passes: the horizontal sums of the k source rows are float32 values, then the vertical sum.
// Float summation order differs between OpenCV's SIMD paths, so parity is to float32 rounding (tests state 1e-3 on the 0..255
// scale; measured ~1e-4), not bit-exact; oracle/preproc_oracle.py is the numpy restatement pinned to cv2.
#include "ops.cuh"
#include "sb_ptx.cuh"

#include <cfloat>

namespace sb {

namespace {

constexpr int PP_TX = 32, PP_TY = 8;      // output pixels per block

struct AxisTap { int s; float c[8]; };

__device__ __forceinline__ void lanczos_taps(int d, int src, int dst, AxisTap& t) {
  const double scale = 1.0 / (static_cast<double>(dst) / static_cast<double>(src));
  float fx = static_cast<float>((d + 0.5) * scale - 0.5);
  const int s = static_cast<int>(floorf(fx));
  fx -= static_cast<float>(s);
  t.s = s;
  if (fx < FLT_EPSILON) {
#pragma unroll
    for (int i = 0; i < 8; ++i) t.c[i] = 0.f;
    t.c[3] = 1.f;
    return;
  }
  const double s45 = 0.70710678118654752440084436210485;
  const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
  const double PI = 3.1415926535897932384626433832795;
  const double y0 = -(static_cast<double>(fx) + 3) * PI * 0.25;
  const double s0 = sin(y0), c0 = cos(y0);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double y = -(static_cast<double>(fx) + 3 - i) * PI * 0.25;
    t.c[i] = static_cast<float>((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
    sum += t.c[i];
  }
  sum = 1.f / sum;
#pragma unroll
  for (int i = 0; i < 8; ++i) t.c[i] *= sum;
}

__device__ __forceinline__ void cubic_taps(int d, int src, int dst, AxisTap& t) {
  const double scale = 1.0 / (static_cast<double>(dst) / static_cast<double>(src));
  const double f = (d + 0.5) * scale - 0.5;
  const double fl = floor(f);
  const double x = f - fl, A = -0.75;
  t.s = static_cast<int>(fl);
  const double c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
  const double c1 = ((A + 2) * x - (A + 3)) * x * x + 1;
  const double c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
  t.c[0] = static_cast<float>(c0); t.c[1] = static_cast<float>(c1); t.c[2] = static_cast<float>(c2);
  t.c[3] = static_cast<float>(1.0 - c0 - c1 - c2);
#pragma unroll
  for (int i = 4; i < 8; ++i) t.c[i] = 0.f;
}

__device__ __forceinline__ int clampi(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

// K-tap separable resample of one output pixel (3 channels) from an HWC image of SrcT (uint8 or float): the K horizontal sums are
// rounded to float32 one after the other (OpenCV's row buffers), then combined vertically.  __fmul_rn / __fadd_rn keep the compiler from
// contracting into FMAs, so the arithmetic is the plain float multiply-add sequence of the restatement.
template <int K, typename SrcT>
__device__ __forceinline__ void resample_px(const SrcT* __restrict__ img, int h, int w, const AxisTap& tx, const AxisTap& ty, float out[3]) {
  out[0] = out[1] = out[2] = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int yy = clampi(ty.s - (K / 2 - 1) + j, h - 1);
    const SrcT* row = img + static_cast<size_t>(yy) * w * 3;
    float hs[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const int xx = clampi(tx.s - (K / 2 - 1) + i, w - 1);
      const SrcT* px = row + xx * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) hs[c] = __fadd_rn(hs[c], __fmul_rn(static_cast<float>(px[c]), tx.c[i]));
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] = __fadd_rn(out[c], __fmul_rn(hs[c], ty.c[j]));
  }
}

// per-crop descriptor (int32 x PP_DESC, device memory; built by surya_b200.recognition.build_preprocess_plan)
enum { PD_SRC_OFF = 0, PD_H, PD_W, PD_NH, PD_NW, PD_HB, PD_WB, PD_SCRATCH_OFF, PD_TILE_ROW, PD_STRIDE };

// stage 1: scale_to_fit's INTER_LANCZOS4 resize, uint8 crop [h, w, 3] -> float32 [nh, nw, 3] in the scratch arena (crops that stay
// inside the size bounds have nh == h, nw == w and skip this kernel's work)
__global__ void __launch_bounds__(PP_TX * PP_TY) pp_lanczos_kernel(const unsigned char* __restrict__ crops, const int* __restrict__ desc,
                                                                   float* __restrict__ scratch) {
  const int* d = desc + static_cast<size_t>(blockIdx.z) * PD_STRIDE;
  const int h = d[PD_H], w = d[PD_W], nh = d[PD_NH], nw = d[PD_NW];
  if (nh == h && nw == w) return;
  const int x0 = blockIdx.x * PP_TX, y0 = blockIdx.y * PP_TY;
  if (x0 >= nw || y0 >= nh) return;
  __shared__ AxisTap sx[PP_TX], sy[PP_TY];
  const int tid = threadIdx.y * PP_TX + threadIdx.x;
  if (tid < PP_TX) { if (x0 + tid < nw) lanczos_taps(x0 + tid, w, nw, sx[tid]); }
  else if (tid < PP_TX + PP_TY) { const int j = tid - PP_TX; if (y0 + j < nh) lanczos_taps(y0 + j, h, nh, sy[j]); }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= nw || y >= nh) return;
  float v[3];
  resample_px<8, unsigned char>(crops + static_cast<unsigned int>(d[PD_SRC_OFF]), h, w, sx[threadIdx.x], sy[threadIdx.y], v);
  float* o = scratch + static_cast<size_t>(d[PD_SCRATCH_OFF]) + (static_cast<size_t>(y) * nw + x) * 3;
  o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
}

// stage 2: INTER_CUBIC resize to [hb, wb] (multiples of patch * merge; identity when already aligned), _image_processor, and the
// merge-block-major tile scatter: tiles[row0 + ((by * gwm + bx) * m + my) * m + mx][c * P * P + py * P + px]
__global__ void __launch_bounds__(PP_TX * PP_TY) pp_cubic_norm_tile_kernel(const unsigned char* __restrict__ crops, const int* __restrict__ desc,
                                                                          const float* __restrict__ scratch, float* __restrict__ tiles,
                                                                          int ld_tiles, int P, int m, float mean0, float mean1, float mean2,
                                                                          float std0, float std1, float std2) {
  const int* d = desc + static_cast<size_t>(blockIdx.z) * PD_STRIDE;
  const int h = d[PD_H], w = d[PD_W], nh = d[PD_NH], nw = d[PD_NW], hb = d[PD_HB], wb = d[PD_WB];
  const int x0 = blockIdx.x * PP_TX, y0 = blockIdx.y * PP_TY;
  if (x0 >= wb || y0 >= hb) return;
  const bool resized1 = !(nh == h && nw == w), resize2 = !(hb == nh && wb == nw);
  __shared__ AxisTap sx[PP_TX], sy[PP_TY];
  const int tid = threadIdx.y * PP_TX + threadIdx.x;
  if (resize2) {
    if (tid < PP_TX) { if (x0 + tid < wb) cubic_taps(x0 + tid, nw, wb, sx[tid]); }
    else if (tid < PP_TX + PP_TY) { const int j = tid - PP_TX; if (y0 + j < hb) cubic_taps(y0 + j, nh, hb, sy[j]); }
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= wb || y >= hb) return;
  const unsigned char* u8 = crops + static_cast<unsigned int>(d[PD_SRC_OFF]);
  const float* f32 = scratch + static_cast<size_t>(d[PD_SCRATCH_OFF]);
  float v[3];
  if (resize2) {
    if (resized1) resample_px<4, float>(f32, nh, nw, sx[threadIdx.x], sy[threadIdx.y], v);
    else resample_px<4, unsigned char>(u8, nh, nw, sx[threadIdx.x], sy[threadIdx.y], v);
  } else {
    const size_t o = (static_cast<size_t>(y) * nw + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = resized1 ? f32[o + c] : static_cast<float>(u8[o + c]);
  }
  const float mean[3] = {mean0, mean1, mean2}, sd[3] = {std0, std1, std2};
  const int gy = y / P, py = y % P, gx = x / P, px = x % P;
  const int gwm = (wb / P) / m;
  const int row = d[PD_TILE_ROW] + (((gy / m) * gwm + gx / m) * m + gy % m) * m + gx % m;
  float* t = tiles + static_cast<size_t>(row) * ld_tiles + py * P + px;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float r = static_cast<float>(static_cast<double>(v[c]) * (1.0 / 255.0));   // image.astype(float64) * rescale_factor -> float32
    t[c * P * P] = __fdiv_rn(__fsub_rn(r, mean[c]), sd[c]);                            // (x - image_mean) / image_std in float32
  }
}

}  // namespace

int rec_preprocess(const unsigned char* crops, const int* desc, int n_crops, int max_nh, int max_nw, int max_hb, int max_wb, int any_stage1,
                   float* scratch, float* tiles, int ld_tiles, int patch, int merge, const float* mean, const float* std3, cudaStream_t st) {
  if (n_crops <= 0) return 0;
  if (!crops || !desc || !tiles || !mean || !std3) { set_error("rec_preprocess: null argument"); return -1; }
  if (any_stage1 && !scratch) { set_error("rec_preprocess: a crop needs the scale_to_fit resize but no scratch arena was given"); return -2; }
  if (patch <= 0 || merge <= 0 || ld_tiles < 3 * patch * patch) { set_error("rec_preprocess: bad patch / merge / tile pitch"); return -3; }
  if (n_crops > 65535) { set_error("rec_preprocess: at most 65535 crops per call"); return -4; }
  const dim3 block(PP_TX, PP_TY);
  if (any_stage1) {
    const dim3 grid((max_nw + PP_TX - 1) / PP_TX, (max_nh + PP_TY - 1) / PP_TY, n_crops);
    pp_lanczos_kernel<<<grid, block, 0, st>>>(crops, desc, scratch);
    if (launch_ok()) return -5;
  }
  const dim3 grid((max_wb + PP_TX - 1) / PP_TX, (max_hb + PP_TY - 1) / PP_TY, n_crops);
  pp_cubic_norm_tile_kernel<<<grid, block, 0, st>>>(crops, desc, scratch, tiles, ld_tiles, patch, merge, mean[0], mean[1], mean[2], std3[0],
                                                    std3[1], std3[2]);
  return launch_ok();
}

}  // namespace sb
