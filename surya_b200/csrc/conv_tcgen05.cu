// surya_b200 — 3x3 (k x k) dense convolution as an implicit GEMM on tcgen05, NHWC activations.
//
//   out[n, oy, ox, :] = epilogue( sum_{r,s,c} in[n, oy*stride + r - pad, ox*stride + s - pad, c] * W[:, r, s, c] )
//
// Replaces the cuDNN implicit-GEMM convolutions of the detection backbone's ConvBlock / FusedMBConv
// (surya/detection/model/encoderdecoder.py:130-171, 228-270).  Same warp-specialised
// pipeline as gemm_tcgen05.cu; the only difference is the A operand: for k-block (tap r,s ; channel chunk c0)
// the producer issues ONE 4-D TMA tile load {BKC channels, TW pixels, TH rows, 1 image} at the tap-shifted
// coordinate.  TMA zero-fills out-of-bounds pixels (= the convolution's zero padding) and its elementStrides
// implement stride 2, so no im2col matrix is ever materialised.  BN is folded into W/bias at pack time.
#include "gemm.cuh"
#include "gemm_epilogue.cuh"
#include "ops.cuh"
#include "sb_ptx.cuh"

namespace sb {

int make_tma_2d_sw(CUtensorMap* map, int dtype, const void* base, int rows, int K, int ld, int box_k, int box_rows,
                   int swizzle_bytes);
int make_tma_nhwc(CUtensorMap* map, int dtype, const void* base, int N, int H, int W, int C, int box_c, int box_w,
                  int box_h, int stride, int swizzle_bytes);

struct ConvKParams {
  GemmKParams g;  // M unused; N = Cout
  int n_img, Ho, Wo, Cin, ksize, stride, pad;
  int tiles_x, tiles_y;
};

template <typename T, int BN, int BKC, int STAGES>
__global__ void __launch_bounds__(384, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const ConvKParams cp) {
  constexpr int BM = 128, TW = 16, TH = 8;
  constexpr uint32_t A_BYTES = BM * BKC * 2, B_BYTES = BN * BKC * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static_assert(BKC == 64 || BKC == 32, "channel chunk must be 64 (128B swizzle) or 32 (64B swizzle)");
  const GemmKParams& p = cp.g;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* epi_smem = smem + STAGES * STAGE_BYTES + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = cp.tiles_x * cp.tiles_y;
  const int m_blocks = cp.n_img * tiles_per_img;
  const int n_blocks = (p.N + BN - 1) / BN;
  const int cpb = cp.Cin / BKC;                       // channel chunks per tap
  const int k_blocks = cp.ksize * cp.ksize * cpb;
  const int num_tiles = m_blocks * n_blocks;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tma_a); tma_prefetch_desc(&tma_b); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], EPI_WARPS); }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto coords = [&](int tile, int& img, int& oy0, int& ox0, int& nb) {
    // n-block fastest inside a group of 8 m-blocks: the (small) weight matrix and the input halo stay L2-hot
    int mb;
    int per_group = p.group_m * n_blocks;
    int g = tile / per_group;
    int first = g * p.group_m;
    int gsz = min(m_blocks - first, p.group_m);
    int r = tile - g * per_group;
    mb = first + r % gsz;
    nb = r / gsz;
    img = mb / tiles_per_img;
    int t = mb - img * tiles_per_img;
    oy0 = (t / cp.tiles_x) * TH;
    ox0 = (t % cp.tiles_x) * TW;
  };

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int img, oy0, ox0, nb;
        coords(tile, img, oy0, ox0, nb);
        for (int kb = 0; kb < k_blocks; ++kb) {
          const int tap = kb / cpb, cc = kb - tap * cpb;
          const int r = tap / cp.ksize, sx = tap - r * cp.ksize;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          tma_load_4d(sa, &tma_a, &full_bar[s], cc * BKC, ox0 * cp.stride + sx - cp.pad, oy0 * cp.stride + r - cp.pad, img);
          tma_load_2d(sa + A_BYTES, &tma_b, &full_bar[s], kb * BKC, nb * BN);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(TypeInfo<T>::umma_fmt, BM, BN);
      int s = 0; uint32_t ph = 0; int as = 0; uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t da = (BKC == 64) ? umma_desc_k128(sa) : umma_desc_k64(sa);
          const uint64_t db = (BKC == 64) ? umma_desc_k128(sa + A_BYTES) : umma_desc_k64(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BKC / 16; ++k) umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull_bar[as]);
        as ^= 1;
        if (as == 0) aph ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int epi_tid = threadIdx.x - 128;
    uint8_t* stage = epi_smem + (warp - 4) * EPI_STAGE_BYTES;
    float* sbias = reinterpret_cast<float*>(epi_smem + EPI_WARPS * EPI_STAGE_BYTES);
    int as = 0; uint32_t aph = 0;
    const bool vec_ok = (p.ldc % 8 == 0) && (!p.residual || p.ldr % 8 == 0);
    const bool v2 = epilogue_v2_ok(p);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int img, oy0, ox0, nb;
      coords(tile, img, oy0, ox0, nb);
      if (v2) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        epilogue_stage_bias<BN>(p, sbias, epi_tid, nb * BN);
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      if (v2) {
        epilogue_tile_v2<T, BN>(tacc, p, stage, sbias, lane, half,
                                [&](int r) {
                                  const int mm = q * 32 + r;
                                  const int yy = oy0 + mm / TW, xx = ox0 + mm % TW;
                                  return (yy < cp.Ho && xx < cp.Wo) ? (img * cp.Ho + yy) * cp.Wo + xx : -1;
                                }, nb * BN);
      } else if (half == 0) {
        const int m = q * 32 + lane;
        const int oy = oy0 + m / TW, ox = ox0 + m % TW;
        const bool row_ok = oy < cp.Ho && ox < cp.Wo;
        const int row = (img * cp.Ho + oy) * cp.Wo + ox;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld_32x32(tacc + c * 32, v);
          tmem_ld_wait();
          const int col0 = nb * BN + c * 32;
          if (col0 >= p.N) continue;
          epilogue_chunk<T>(v, p, row, row_ok, col0, vec_ok);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

template <typename T, int BN, int BKC, int STAGES>
static int launch_conv(const ConvArgs& a, cudaStream_t stream) {
  constexpr uint32_t STAGE_BYTES = 128 * BKC * 2 + BN * BKC * 2;
  constexpr size_t SMEM = STAGES * STAGE_BYTES + 1024 + 256 + epi_smem_bytes<BN>();
  auto kern = conv_igemm_kernel<T, BN, BKC, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM) != cudaSuccess) {
      cudaGetLastError();
      set_error("conv: cudaFuncSetAttribute(smem=%zu) failed", SMEM);
      return -10;
    }
    attr_set = true;
  }
  const int Ho = (a.H + 2 * a.pad - a.ksize) / a.stride + 1, Wo = (a.W + 2 * a.pad - a.ksize) / a.stride + 1;
  CUtensorMap ma, mb;
  const int sw = BKC * 2;
  int rc = make_tma_nhwc(&ma, a.dtype, a.in, a.n_img, a.H, a.W, a.Cin, BKC, 16, 8, a.stride, sw);
  if (rc) return rc;
  rc = make_tma_2d_sw(&mb, a.dtype, a.weight, a.Cout, a.ksize * a.ksize * a.Cin, a.ksize * a.ksize * a.Cin, BKC, BN, sw);
  if (rc) return rc;
  ConvKParams cp{};
  cp.g.M = a.n_img * Ho * Wo; cp.g.N = a.Cout; cp.g.K = a.ksize * a.ksize * a.Cin;
  cp.g.C = a.out; cp.g.ldc = a.Cout; cp.g.bias = a.bias; cp.g.residual = a.residual; cp.g.ldr = a.Cout;
  cp.g.act = a.act; cp.g.swiglu = 0; cp.g.out_f32 = 0; cp.g.group_m = 8; cp.g.group_k = 0; cp.g.dbg = nullptr; cp.g.w_constant = 0;
  cp.n_img = a.n_img; cp.Ho = Ho; cp.Wo = Wo; cp.Cin = a.Cin; cp.ksize = a.ksize; cp.stride = a.stride; cp.pad = a.pad;
  cp.tiles_x = (Wo + 15) / 16; cp.tiles_y = (Ho + 7) / 8;
  const int tiles = a.n_img * cp.tiles_x * cp.tiles_y * ((a.Cout + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, 384, SMEM, stream>>>(ma, mb, cp);
  return launch_ok();
}

template <typename T>
static int conv_typed(const ConvArgs& a, cudaStream_t st) {
  const int bn = a.Cout >= 256 ? 256 : (a.Cout >= 128 ? 128 : (a.Cout >= 64 ? 64 : 32));
  if (a.Cin % 64 == 0) {
    switch (bn) {
      case 256: return launch_conv<T, 256, 64, 4>(a, st);
      case 128: return launch_conv<T, 128, 64, 6>(a, st);
      case 64: return launch_conv<T, 64, 64, 8>(a, st);
      default: return launch_conv<T, 32, 64, 8>(a, st);
    }
  }
  if (a.Cin % 32 == 0) {
    switch (bn) {
      case 256: return launch_conv<T, 256, 32, 6>(a, st);
      case 128: return launch_conv<T, 128, 32, 8>(a, st);
      case 64: return launch_conv<T, 64, 32, 8>(a, st);
      default: return launch_conv<T, 32, 32, 8>(a, st);
    }
  }
  set_error("conv_igemm: Cin must be a multiple of 32 (got %d)", a.Cin);
  return -20;
}

int conv_igemm(const ConvArgs& a, cudaStream_t st) {
  if (a.n_img <= 0) return 0;
  if (a.stride != 1 && a.stride != 2) { set_error("conv_igemm: stride must be 1 or 2"); return -21; }
  if (a.Cout % 8) { set_error("conv_igemm: Cout must be a multiple of 8"); return -22; }
  return a.dtype == DT_BF16 ? conv_typed<__nv_bfloat16>(a, st) : conv_typed<__half>(a, st);
}

}  // namespace sb
