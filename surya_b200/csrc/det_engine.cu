// surya_b200 — detection engine: executes the EfficientViT-seg op program over NHWC workspaces.
//
// Stands in for EfficientViTForSemanticSegmentation.forward (surya/detection/model/encoderdecoder.py:725-753):
// dense 3x3 convs -> conv_tcgen05.cu, 1x1 / grouped 1x1 convs -> gemm_tcgen05.cu, everything else -> det_ops.cu.
#include "../../include/surya_b200.h"
#include "ops.cuh"
#include "sb_ptx.cuh"

#include <vector>

using namespace sb;

#define CK(x)            \
  do {                   \
    int rc_ = (x);       \
    if (rc_) return rc_; \
  } while (0)

struct BufDim { int H = 0, W = 0, C = 0; };

struct sb_det_engine {
  int dtype = DT_F16;
  std::vector<sb_det_op> ops;
  std::vector<const void*> w;
  std::vector<long long> buf_elems;
  std::vector<void*> bufs;
  std::vector<BufDim> dims;
  int max_batch = 0;
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
};

static size_t al256(size_t v) { return (v + 255) & ~size_t(255); }

extern "C" {

int sb_det_create(int dtype, const sb_det_op* ops, int n_ops, const void* const* weights, int n_weights,
                  const long long* buf_elems, int n_bufs, int max_batch, sb_det_engine** out) {
  if (!ops || !weights || !buf_elems || !out) { set_error("sb_det_create: null argument"); return -1; }
  auto* e = new sb_det_engine();
  e->dtype = dtype;
  e->ops.assign(ops, ops + n_ops);
  e->w.assign(weights, weights + n_weights);
  e->buf_elems.assign(buf_elems, buf_elems + n_bufs);
  e->max_batch = max_batch;
  for (const auto& op : e->ops) {
    if (op.w >= n_weights || op.b >= n_weights) { set_error("sb_det_create: weight index out of range"); delete e; return -2; }
    if (op.dst >= n_bufs) { set_error("sb_det_create: buffer index out of range"); delete e; return -2; }
  }
  size_t total = 0;
  for (long long n : e->buf_elems) total += al256(static_cast<size_t>(n) * max_batch * 2);
  cudaError_t ce = cudaMalloc(&e->arena, total);
  if (ce != cudaSuccess) {
    cudaGetLastError();
    set_error("sb_det_create: cudaMalloc(%zu MiB) failed: %s", total >> 20, cudaGetErrorString(ce));
    delete e;
    return -3;
  }
  e->arena_bytes = total;
  uint8_t* p = e->arena;
  for (long long n : e->buf_elems) {
    e->bufs.push_back(p);
    p += al256(static_cast<size_t>(n) * max_batch * 2);
  }
  e->dims.resize(n_bufs);
  *out = e;
  return 0;
}

void sb_det_destroy(sb_det_engine* e) {
  if (!e) return;
  if (e->arena) cudaFree(e->arena);
  delete e;
}

size_t sb_det_workspace_bytes(const sb_det_engine* e) { return e ? e->arena_bytes : 0; }

int sb_det_forward(sb_det_engine* e, const void* pixel_values, int in_f32, int B, int H, int W, void* logits, void* stream) {
  if (!e) { set_error("sb_det_forward: null engine"); return -1; }
  if (B > e->max_batch) { set_error("sb_det_forward: batch %d exceeds capacity %d", B, e->max_batch); return -2; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int dt = e->dtype;
  auto fits = [&](int buf, int h, int w, int c) -> bool {
    return static_cast<long long>(h) * w * c <= e->buf_elems[buf];
  };
  for (size_t i = 0; i < e->ops.size(); ++i) {
    const sb_det_op& op = e->ops[i];
    const int s0 = op.src[0];
    BufDim in = (s0 >= 0) ? e->dims[s0] : BufDim{H, W, op.cin};
    const void* in_ptr = (s0 >= 0) ? e->bufs[s0] : pixel_values;
    BufDim od;
    const void* res = op.res >= 0 ? e->bufs[op.res] : nullptr;
    switch (op.op) {
      case SB_DOP_STEM: {
        od = {H / 2, W / 2, op.cout};
        if (!fits(op.dst, od.H, od.W, od.C)) { set_error("det: buffer %d too small for %dx%dx%d", op.dst, od.H, od.W, od.C); return -4; }
        CK(det_stem_conv(dt, pixel_values, in_f32, static_cast<const float*>(e->w[op.w]), static_cast<const float*>(e->w[op.b]),
                         e->bufs[op.dst], B, H, W, op.cout, st));
        break;
      }
      case SB_DOP_CONV: {
        od = {(in.H + 2 * op.pad - op.k) / op.stride + 1, (in.W + 2 * op.pad - op.k) / op.stride + 1, op.cout};
        if (!fits(op.dst, od.H, od.W, od.C)) { set_error("det: buffer %d too small", op.dst); return -4; }
        // FusedMBConv: 3x3 expand (Hardswish) directly followed by its 1x1 project -> one back-to-back GEMM kernel
        if (i + 1 < e->ops.size() && op.res < 0 && op.b >= 0) {
          const sb_det_op& pj = e->ops[i + 1];
          if (pj.op == SB_DOP_PW && pj.src[0] == op.dst && pj.cin == op.cout && pj.b >= 0 &&
              fmb_fused_ok(op.cin, op.cout, pj.cout, op.k, op.stride, op.pad, op.act, pj.act) && fits(pj.dst, od.H, od.W, pj.cout)) {
            CK(fmb_fused(dt, in_ptr, e->w[op.w], static_cast<const float*>(e->w[op.b]), e->w[pj.w], static_cast<const float*>(e->w[pj.b]),
                         pj.res >= 0 ? e->bufs[pj.res] : nullptr, e->bufs[pj.dst], B, in.H, in.W, op.cin, op.cout, pj.cout, op.stride, st));
            e->dims[op.dst] = od;
            e->dims[pj.dst] = {od.H, od.W, pj.cout};
            ++i;
            continue;
          }
        }
        ConvArgs a;
        a.dtype = dt; a.in = in_ptr; a.weight = e->w[op.w]; a.bias = op.b >= 0 ? static_cast<const float*>(e->w[op.b]) : nullptr;
        a.residual = res; a.out = e->bufs[op.dst]; a.n_img = B; a.H = in.H; a.W = in.W; a.Cin = op.cin; a.Cout = op.cout;
        a.ksize = op.k; a.stride = op.stride; a.pad = op.pad; a.act = op.act;
        CK(conv_igemm(a, st));
        break;
      }
      case SB_DOP_PW: {
        od = {in.H, in.W, op.cout};
        if (!fits(op.dst, od.H, od.W, od.C)) { set_error("det: buffer %d too small", op.dst); return -4; }
        GemmArgs g;
        g.dtype = dt; g.A = in_ptr; g.lda = op.cin; g.W = e->w[op.w]; g.ldw = op.cin; g.C = e->bufs[op.dst]; g.ldc = op.cout;
        g.M = B * in.H * in.W; g.N = op.cout; g.K = op.cin;
        g.bias = op.b >= 0 ? static_cast<const float*>(e->w[op.b]) : nullptr;
        g.residual = res; g.ldr = op.cout; g.act = op.act;
        CK(gemm_launch(g, st));
        break;
      }
      case SB_DOP_DW: {
        od = {(in.H + 2 * op.pad - op.k) / op.stride + 1, (in.W + 2 * op.pad - op.k) / op.stride + 1, op.cout};
        if (!fits(op.dst, od.H, od.W, od.C)) { set_error("det: buffer %d too small", op.dst); return -4; }
        CK(det_dwconv(dt, in_ptr, e->w[op.w], op.b >= 0 ? static_cast<const float*>(e->w[op.b]) : nullptr, e->bufs[op.dst], B,
                      in.H, in.W, op.cin, op.k, op.stride, op.pad, op.act, st));
        break;
      }
      case SB_DOP_GPW: {
        od = {in.H, in.W, op.cout};
        if (!fits(op.dst, od.H, od.W, od.C)) { set_error("det: buffer %d too small", op.dst); return -4; }
        const int gk = op.cin / op.groups, gn = op.cout / op.groups;
        if (gn != 32 || gk > 64) { set_error("det: grouped 1x1 conv supports 32 outputs and <= 64 inputs per group"); return -5; }
        GemmArgs g;
        g.dtype = dt; g.A = in_ptr; g.lda = op.cin; g.W = e->w[op.w]; g.ldw = 64; g.C = e->bufs[op.dst]; g.ldc = op.cout;
        g.M = B * in.H * in.W; g.N = op.cout; g.K = 64;
        g.group_k = gk; g.group_n = gn; g.a_cols = op.cin;
        CK(gemm_launch(g, st));
        break;
      }
      case SB_DOP_MLA: {
        od = {in.H, in.W, 2 * op.heads * op.dim};
        if (!fits(op.dst, od.H, od.W, od.C)) { set_error("det: buffer %d too small", op.dst); return -4; }
        CK(det_lite_mla(dt, in_ptr, e->bufs[op.src[1]], e->bufs[op.dst], B, in.H * in.W, op.heads, op.dim, op.eps, st));
        break;
      }
      case SB_DOP_UPCAT: {
        const void* srcs[4]; int hs[4], ws[4];
        for (int j = 0; j < op.n_src; ++j) { srcs[j] = e->bufs[op.src[j]]; hs[j] = e->dims[op.src[j]].H; ws[j] = e->dims[op.src[j]].W; }
        od = {hs[0], ws[0], op.cout};
        // decode head: UPCAT -> 1x1 fuse conv (ReLU) -> classifier collapse into one kernel when the shapes are the default head's
        if (i + 2 < e->ops.size()) {
          const sb_det_op& fz = e->ops[i + 1];
          const sb_det_op& cl = e->ops[i + 2];
          if (fz.op == SB_DOP_PW && fz.src[0] == op.dst && fz.act == ACT_RELU && fz.res < 0 && fz.b >= 0 && cl.op == SB_DOP_CLS &&
              cl.src[0] == fz.dst && fz.cin == op.cout &&
              det_head_fused_ok(op.n_src, op.cin, cl.cout, fz.cin, fz.cout, hs, ws, od.H, od.W)) {
            CK(det_head_fused(dt, srcs, hs, ws, op.src_off, op.n_src, op.cin, e->w[fz.w], static_cast<const float*>(e->w[fz.b]),
                              e->w[cl.w], e->w[cl.b], logits, B, od.H, od.W, st));
            e->dims[op.dst] = od;
            e->dims[fz.dst] = {od.H, od.W, fz.cout};
            i += 2;
            continue;
          }
        }
        if (!fits(op.dst, od.H, od.W, od.C)) { set_error("det: buffer %d too small", op.dst); return -4; }
        CK(det_upsample_cat(dt, srcs, hs, ws, op.src_off, op.n_src, op.cin, e->bufs[op.dst], B, od.H, od.W, st));
        break;
      }
      case SB_DOP_CLS: {
        od = {in.H, in.W, op.cout};
        CK(det_classifier(dt, in_ptr, e->w[op.w], e->w[op.b], logits, static_cast<long long>(B) * in.H * in.W, op.cin,
                          in.H * in.W, op.cout, st));
        break;
      }
      default: set_error("det: unknown op %d", op.op); return -6;
    }
    if (op.dst >= 0) e->dims[op.dst] = od;
  }
  return 0;
}

int sb_det_upsample(int dtype, const void* logits, float* out, int planes, int hs, int ws, int HO, int WO, void* stream) {
  return det_upsample_nchw(dtype, logits, out, planes, hs, ws, HO, WO, static_cast<cudaStream_t>(stream));
}

int sb_det_normalize_u8(int dtype, const unsigned char* pages_nhwc, void* out_nchw, int B, int H, int W, void* stream) {
  if (!pages_nhwc || !out_nchw) { set_error("sb_det_normalize_u8: null argument"); return -1; }
  return det_normalize_u8(dtype, pages_nhwc, out_nchw, B, H, W, static_cast<cudaStream_t>(stream));
}

int sb_det_text_front(int dtype, const void* logits, int n_labels, int B, int hs, int ws, int HO, int WO, void* map16,
                      unsigned char* mask, float* thresholds, unsigned int* hist_scratch, float text_threshold, float low_text,
                      void* stream) {
  if (!logits || !map16 || !mask || !thresholds || !hist_scratch) { set_error("sb_det_text_front: null argument"); return -1; }
  return det_text_front(dtype, logits, n_labels, B, hs, ws, HO, WO, map16, mask, thresholds, hist_scratch, text_threshold,
                        low_text, static_cast<cudaStream_t>(stream));
}

int sb_det_debug_copy(sb_det_engine* e, int buf, void* dst, size_t bytes, void* stream) {
  if (!e || buf < 0 || buf >= (int)e->bufs.size()) { set_error("sb_det_debug_copy: bad buffer"); return -1; }
  cudaError_t ce = cudaMemcpyAsync(dst, e->bufs[buf], bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream));
  if (ce != cudaSuccess) { cudaGetLastError(); set_error("sb_det_debug_copy: %s", cudaGetErrorString(ce)); return -2; }
  return 0;
}

int sb_conv2d_nhwc(int dtype, const void* in, const void* weight, const float* bias, const void* residual, void* out,
                   int n_img, int H, int W, int Cin, int Cout, int ksize, int stride, int pad, int act, void* stream) {
  ConvArgs a;
  a.dtype = dtype; a.in = in; a.weight = weight; a.bias = bias; a.residual = residual; a.out = out; a.n_img = n_img;
  a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ksize = ksize; a.stride = stride; a.pad = pad; a.act = act;
  return conv_igemm(a, static_cast<cudaStream_t>(stream));
}

int sb_dwconv_nhwc(int dtype, const void* in, const void* w, const float* bias, void* out, int B, int H, int W, int C,
                   int ks, int stride, int pad, int act, void* stream) {
  return det_dwconv(dtype, in, w, bias, out, B, H, W, C, ks, stride, pad, act, static_cast<cudaStream_t>(stream));
}

int sb_gemm_grouped(int dtype, const void* A, int lda, int a_cols, const void* W, int ldw, void* C, int ldc, int M, int N,
                    int Kpad, int group_k, int group_n, void* stream) {
  GemmArgs g;
  g.dtype = dtype; g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = Kpad;
  g.group_k = group_k; g.group_n = group_n; g.a_cols = a_cols;
  return gemm_launch(g, static_cast<cudaStream_t>(stream));
}

int sb_lite_mla(int dtype, const void* qkv_a, const void* qkv_b, void* out, int B, int HW, int heads, int dim, float eps,
                void* stream) {
  return det_lite_mla(dtype, qkv_a, qkv_b, out, B, HW, heads, dim, eps, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
