// surya_b200 — layout / table_rec engine: Donut-Swin encoder + ADETR box decoder behind the C ABI (sb_layout_*).
//
// Same kernels and the same op order as the op-by-op path in surya_b200/layout.py (impl="ops": one C-ABI call per kernel; it stays as the readable
// statement of the sequence and as a cross-check: tests require bit-identical results); here the layer loops, the
// workspaces, the K/V caches and the greedy decode loop (one CUDA graph per 8 steps) live in C++ so that nothing on the
// forward path depends on the host between the first and the last kernel.
//
// Reference: DonutSwinLayoutModel.forward / DonutSwinModel.forward (surya/layout/model/encoder.py:33-81,
// surya/table_rec/model/encoder.py:35-87), DonutSwinEncoder/Stage/Layer (surya/common/donut/encoder.py:534-931),
// SuryaADETRDecoderModel + layers (surya/common/adetr/decoder.py:360-652), SuryaLayoutDecoder / SuryaTableRecDecoder heads
// (surya/layout/model/decoder.py:95-126, surya/table_rec/model/decoder.py:121-155), per-step token logic
// (surya/layout/__init__.py:111-137, surya/table_rec/__init__.py:62-131).
#include "../../include/surya_b200.h"
#include "ops.cuh"

#include <cmath>
#include <cstring>
#include <vector>

using namespace sb;

#define CK(call)                 \
  do {                           \
    int rc_ = (call);            \
    if (rc_) return rc_;         \
  } while (0)

namespace {
constexpr int GRAPH_GROUP = 8;
}

struct sb_layout_engine {
  sb_layout_config c;
  std::vector<const void*> w;
  size_t esz = 2;
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  // encoder workspaces
  void *x = nullptr, *h = nullptr, *qkv = nullptr, *ao = nullptr, *mlp = nullptr;
  // decoder state (persistent: the step graphs point at these)
  void *dx = nullptr, *dn = nullptr, *dq = nullptr, *da = nullptr, *dcross = nullptr, *dres = nullptr, *dqkv = nullptr,
       *dm = nullptr, *dh = nullptr;
  void* enc = nullptr;                 // [B, Lk, enc_hidden]
  std::vector<void*> ckv, kc, vc;      // per decoder layer
  int* slot = nullptr;
  long long* tok = nullptr;            // [B, ncol]
  int* pos = nullptr;                  // [B]
  int* base = nullptr;                 // [1]
  float* out_bbox = nullptr;           // [B, 6]
  float* out_head[4] = {nullptr, nullptr, nullptr, nullptr};
  unsigned char* done = nullptr;
  int Lk = 0;
  // graphs
  cudaGraphExec_t g_group = nullptr, g_one = nullptr;
  const void* graph_key[8] = {nullptr};
  int graph_batch = 0;
  long long launches_per_step = 0;
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;

  int ncol() const { return 6 + c.n_out_heads; }
  // ---- weight table (see include/surya_b200.h)
  const void* WF(int i) const { return w[i]; }
  int stage_base(int s) const {      // index of the stage's position table
    int idx = SB_LW_ENC_FIXED;
    for (int i = 0; i < s; ++i) idx += 1 + c.depths[i] * SB_LW_ENC_LAYER + (i < c.n_stages - 1 ? SB_LW_ENC_MERGE : 0);
    return idx;
  }
  const void* WL(int s, int l, int k) const { return w[stage_base(s) + 1 + l * SB_LW_ENC_LAYER + k]; }
  const void* WM(int s, int k) const { return w[stage_base(s) + 1 + c.depths[s] * SB_LW_ENC_LAYER + k]; }
  int dec_base() const { return stage_base(c.n_stages); }
  int n_tables() const { return c.kind == 1 ? 13 : 15; }
  const void* WT(int i) const { return w[dec_base() + i]; }
  const void* WD(int l, int k) const { return w[dec_base() + n_tables() + l * SB_LW_DEC_LAYER + k]; }
  const void* WX(int k) const { return w[dec_base() + n_tables() + c.dec_layers * SB_LW_DEC_LAYER + k]; }
  int n_weights() const { return dec_base() + n_tables() + c.dec_layers * SB_LW_DEC_LAYER + SB_LW_DEC_TAIL + (c.kind == 1 ? 5 : 3); }
};

static size_t al256(size_t v) { return (v + 255) & ~size_t(255); }

static int lin(const sb_layout_engine* e, const void* A, int lda, const void* Wt, int ldw, void* C, int ldc, int M, int N, int K,
               const void* bias_f32, const void* residual, int ldr, int act, int swiglu, cudaStream_t st, int allow_splitk = 0) {
  GemmArgs a;
  a.dtype = e->c.dtype;
  a.A = A; a.lda = lda; a.W = Wt; a.ldw = ldw; a.C = C; a.ldc = ldc;
  a.M = M; a.N = N; a.K = K;
  a.bias = static_cast<const float*>(bias_f32);
  a.residual = residual; a.ldr = ldr; a.act = act; a.swiglu = swiglu;
  a.w_constant = 1;
  a.allow_splitk = allow_splitk;
  return gemm_launch(a, st);
}

// ------------------------------------------------------------------------------------------------ Swin encoder
static int run_encoder(sb_layout_engine* e, const void* pixels, int is_f32, int B, void* out, cudaStream_t st) {
  const sb_layout_config& c = e->c;
  const int dt = c.dtype;
  int H = c.img_h / c.patch, W = c.img_w / c.patch;
  const int K0 = c.in_ch * c.patch * c.patch;
  const int Kp = (K0 + 63) / 64 * 64;
  CK(patch_gather(dt, pixels, is_f32, e->h, B, c.in_ch, c.img_h, c.img_w, c.patch, Kp, st));
  long long rows = static_cast<long long>(B) * H * W;
  int C = c.embed_dim;
  CK(lin(e, e->h, Kp, e->WF(SB_LW_PE_W), Kp, e->x, C, static_cast<int>(rows), C, Kp, e->WF(SB_LW_PE_B), nullptr, 0, ACT_NONE, 0, st));
  CK(layernorm(dt, e->x, e->WF(SB_LW_PE_LN_W), e->WF(SB_LW_PE_LN_B), e->x, static_cast<int>(rows), C, 1e-5f, st));
  void* x = e->x;
  void* alt = e->ao;     // patch merging ping-pongs between the two C-wide buffers
  for (int s = 0; s < c.n_stages; ++s) {
    const int nh = c.heads[s];
    const int R = static_cast<int>(rows);
    CK(add_bcast_rows(dt, x, e->w[e->stage_base(s)], rows, H * W, C, st));
    for (int l = 0; l < c.depths[s]; ++l) {
      const int shift = (l % 2 == 1 && (H < W ? H : W) > c.window) ? c.window / 2 : 0;
      void* attn_out = (x == e->x) ? e->ao : e->x;   // any C-wide buffer that is not x
      CK(layernorm(dt, x, e->WL(s, l, SB_LWE_LN1_W), e->WL(s, l, SB_LWE_LN1_B), e->h, R, C, c.enc_ln_eps, st));
      CK(lin(e, e->h, C, e->WL(s, l, SB_LWE_QKV_W), C, e->qkv, 3 * C, R, 3 * C, C, e->WL(s, l, SB_LWE_QKV_B), nullptr, 0, ACT_NONE,
             0, st));
      CK(swin_window_attn(dt, e->qkv, static_cast<const float*>(e->WL(s, l, SB_LWE_QKV_B)), e->WL(s, l, SB_LWE_RPB), attn_out, B, H, W, C, nh,
                          shift, st));
      // x = attn_out @ Wo^T + bo + x   (written over h first, then swapped in: the GEMM must not alias its residual's rows
      // with a different pitch; here both are [R, C] so in-place on x is the same read-then-write per element)
      CK(lin(e, attn_out, C, e->WL(s, l, SB_LWE_O_W), C, x, C, R, C, C, e->WL(s, l, SB_LWE_O_B), x, C, ACT_NONE, 0, st));
      CK(layernorm(dt, x, e->WL(s, l, SB_LWE_LN2_W), e->WL(s, l, SB_LWE_LN2_B), e->h, R, C, c.enc_ln_eps, st));
      CK(lin(e, e->h, C, e->WL(s, l, SB_LWE_FC1_W), C, e->mlp, 4 * C, R, 4 * C, C, e->WL(s, l, SB_LWE_FC1_B), nullptr, 0,
             ACT_GELU_ERF, 0, st));
      CK(lin(e, e->mlp, 4 * C, e->WL(s, l, SB_LWE_FC2_W), 4 * C, x, C, R, C, 4 * C, e->WL(s, l, SB_LWE_FC2_B), x, C, ACT_NONE, 0, st));
    }
    if (s < c.n_stages - 1) {
      if ((H | W) & 1) {
        // the reference pads here (encoder.py:281-287) and then fails on its floor-sized stage position table (:730-735, 773-776)
        set_error("sb_layout_encode: stage %d has an odd %dx%d token grid (the reference's stage position table cannot follow it)", s, H, W);
        return -31;
      }
      CK(patch_merge_gather(dt, x, e->mlp, B, H, W, C, st));
      rows /= 4;
      H /= 2; W /= 2;
      CK(layernorm(dt, e->mlp, e->WM(s, 0), e->WM(s, 1), e->mlp, static_cast<int>(rows), 4 * C, 1e-5f, st));
      void* nx = (x == e->x) ? alt : e->x;
      CK(lin(e, e->mlp, 4 * C, e->WM(s, 2), 4 * C, nx, 2 * C, static_cast<int>(rows), 2 * C, 4 * C, nullptr, nullptr, 0, ACT_NONE, 0, st));
      x = nx;
      C *= 2;
    }
  }
  CK(add_bcast_rows(dt, x, e->WF(SB_LW_ENC_POS), rows, H * W, C, st));
  if (cudaMemcpyAsync(out, x, static_cast<size_t>(rows) * C * e->esz, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
    cudaGetLastError();
    set_error("sb_layout_encode: copy-out failed");
    return -30;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ ADETR decoder
// One q_len = 1 call on the persistent state: embedding of e->tok rows -> layers -> heads into e->out_*.
static int run_token(sb_layout_engine* e, const long long* tok, const int* pos, int B, cudaStream_t st) {
  const sb_layout_config& c = e->c;
  const int dt = c.dtype, Hd = c.hidden, nh = c.n_heads, nkv = c.n_kv, hd = c.head_dim;
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  const void* tables[15];
  for (int i = 0; i < e->n_tables(); ++i) tables[i] = e->WT(i);
  if (c.kind == 1) CK(label_embed(dt, tok, tables, e->dx, B, c.box_w, c.prop_w, c.bbox_size, c.vocab, st));
  else CK(bbox_embed_sum(dt, tok, tables, e->dx, B, Hd, c.bbox_size, st));
  const int QW = (nh + 2 * nkv) * hd, KVW = 2 * nkv * hd;
  for (int l = 0; l < c.dec_layers; ++l) {
    // cross attention over the cached encoder K/V
    CK(rmsnorm(dt, e->dx, Hd, e->WD(l, SB_LWD_CROSS_NORM), e->dn, Hd, B, Hd, c.rms_eps, nullptr, st, 1));
    CK(lin(e, e->dn, Hd, e->WD(l, SB_LWD_CQ_W), Hd, e->dq, nh * hd, B, nh * hd, Hd, nullptr, nullptr, 0, ACT_NONE, 0, st));
    const uint8_t* kv = static_cast<const uint8_t*>(e->ckv[l]);
    CK(attn_single_query(dt, e->dq, nh * hd, kv, kv + static_cast<size_t>(nkv) * hd * e->esz, static_cast<long long>(e->Lk) * KVW, hd,
                         KVW, e->da, nh * hd, B, nh, nkv, hd, e->Lk, scale, st));
    CK(lin(e, e->da, nh * hd, e->WD(l, SB_LWD_CO_W), nh * hd, e->dcross, Hd, B, Hd, nh * hd, e->WD(l, SB_LWD_CO_B), e->dx, Hd, ACT_NONE,
           0, st));
    // causal self attention (RoPE + cache append fused into the attention kernel)
    CK(rmsnorm(dt, e->dcross, Hd, e->WD(l, SB_LWD_SELF_NORM), e->dn, Hd, B, Hd, c.rms_eps, nullptr, st, 1));
    CK(lin(e, e->dn, Hd, e->WD(l, SB_LWD_SQKV_W), Hd, e->dqkv, QW, B, QW, Hd, nullptr, nullptr, 0, ACT_NONE, 0, st));
    DecodeAttnArgs a;
    a.dtype = dt; a.qkv = e->dqkv; a.ld = QW; a.kcache = e->kc[l]; a.vcache = e->vc[l]; a.slot = e->slot; a.pos = pos;
    a.inv_freq = static_cast<const float*>(e->WX(SB_LWX_INV_FREQ));
    a.out = e->da; a.ldo = nh * hd; a.batch = B; a.n_heads = nh; a.n_kv_heads = nkv; a.head_dim = hd; a.s_max = c.s_max;
    a.scale = scale;
    CK(decode_attn(a, st));
    const void* res_src = c.double_residual ? e->dx : e->dcross;
    CK(lin(e, e->da, nh * hd, e->WD(l, SB_LWD_SO_W), nh * hd, e->dres, Hd, B, Hd, nh * hd, e->WD(l, SB_LWD_SO_B), res_src, Hd, ACT_NONE,
           0, st));
    // GeGLU MLP
    CK(rmsnorm(dt, e->dres, Hd, e->WD(l, SB_LWD_MLP_NORM), e->dn, Hd, B, Hd, c.rms_eps, nullptr, st, 1));
    CK(lin(e, e->dn, Hd, e->WD(l, SB_LWD_GU_W), Hd, e->dm, c.inter, B, 2 * c.inter, Hd, nullptr, nullptr, 0, ACT_GELU_TANH, 1, st));
    CK(lin(e, e->dm, c.inter, e->WD(l, SB_LWD_DOWN_W), c.inter, e->dx, Hd, B, Hd, c.inter, nullptr, e->dres, Hd, ACT_NONE, 0, st, 1));
  }
  CK(rmsnorm(dt, e->dx, Hd, e->WX(SB_LWX_FINAL_NORM), e->dn, Hd, B, Hd, c.rms_eps, nullptr, st, 1));
  CK(layernorm(dt, e->dn, e->WX(SB_LWX_OUT_LN_W), e->WX(SB_LWX_OUT_LN_B), e->dh, B, Hd, c.dec_ln_eps, st));
  const int hb = e->dec_base() + e->n_tables() + c.dec_layers * SB_LW_DEC_LAYER + SB_LW_DEC_TAIL;
  if (c.kind == 1) {   // table: bbox (sigmoid), category, merges, colspan, is_header — all bias-free
    CK(small_head(dt, e->dh, Hd, e->w[hb + 0], nullptr, B, Hd, 6, 1, e->out_bbox, nullptr, 0.f, st));
    for (int k = 0; k < 4; ++k)
      CK(small_head(dt, e->dh, Hd, e->w[hb + 1 + k], nullptr, B, Hd, c.head_n[k], 0, e->out_head[k], nullptr, 0.f, st));
  } else {             // layout: bbox head with bias + sigmoid, class head
    CK(small_head(dt, e->dh, Hd, e->w[hb + 0], e->w[hb + 1], B, Hd, 6, 1, e->out_bbox, nullptr, 0.f, st));
    CK(small_head(dt, e->dh, Hd, e->w[hb + 2], nullptr, B, Hd, c.head_n[0], 0, e->out_head[0], nullptr, 0.f, st));
  }
  return 0;
}

static int run_next(sb_layout_engine* e, int B, int hist_T, long long* hist_tok, float* hist_bbox, float* const* hist_heads,
                    unsigned char* hist_done, cudaStream_t st) {
  const sb_layout_config& c = e->c;
  const float* heads[4];
  int hn[4], hm[4];
  for (int k = 0; k < c.n_out_heads; ++k) { heads[k] = e->out_head[k]; hn[k] = c.head_n[k]; hm[k] = 0; }
  if (c.kind == 1) hm[2] = 1;   // colspan: round(max(v, 1))
  return box_next_token(e->out_bbox, heads, hn, hm, c.n_out_heads, static_cast<float>(c.bbox_size), e->tok, e->done,
                        c.kind == 1 ? 0 : -1, c.eos, c.pad, B, e->pos, e->base, hist_T, hist_tok, hist_bbox, hist_heads, hist_done,
                        st);
}

extern "C" {

int sb_layout_create(const sb_layout_config* cfg, const void* const* weights, int n_weights, sb_layout_engine** out) {
  if (!cfg || !weights || !out) { set_error("sb_layout_create: null argument"); return -1; }
  auto* e = new sb_layout_engine();
  e->c = *cfg;
  const sb_layout_config& c = e->c;
  if (c.n_stages < 1 || c.n_stages > 4 || c.n_out_heads < 1 || c.n_out_heads > 4 || c.head_dim != 64) {
    set_error("sb_layout_create: 1..4 stages, 1..4 output heads and head_dim 64 are supported");
    delete e;
    return -2;
  }
  if (c.patch <= 0 || c.img_h % c.patch || c.img_w % c.patch) {
    // the reference would zero pad the pixels (encoder.py:232-239) and then fail on its floor-sized position tables
    set_error("sb_layout_create: image size %dx%d is not a multiple of the patch size %d", c.img_h, c.img_w, c.patch);
    delete e;
    return -2;
  }
  if (n_weights != e->n_weights()) {
    set_error("sb_layout_create: expected %d weight pointers, got %d", e->n_weights(), n_weights);
    delete e;
    return -3;
  }
  for (int i = 0; i < n_weights; ++i)
    if (!weights[i]) { set_error("sb_layout_create: weight pointer %d is null", i); delete e; return -4; }
  e->w.assign(weights, weights + n_weights);
  const size_t es = e->esz;
  const size_t B = c.max_batch;
  const size_t rows0 = B * (c.img_h / c.patch) * (c.img_w / c.patch);
  const size_t C0 = c.embed_dim;
  const int K0 = c.in_ch * c.patch * c.patch, Kp = (K0 + 63) / 64 * 64;
  const size_t hcols = C0 > static_cast<size_t>(Kp) ? C0 : Kp;
  int Hl = c.img_h / c.patch, Wl = c.img_w / c.patch;
  for (int s = 0; s < c.n_stages - 1; ++s) { Hl /= 2; Wl /= 2; }
  const size_t Lk = static_cast<size_t>(Hl) * Wl;
  const size_t Hd = c.hidden, QW = (c.n_heads + 2 * c.n_kv) * c.head_dim, KVW = 2 * c.n_kv * c.head_dim;
  const size_t cache = B * c.n_kv * c.s_max * c.head_dim * es;
  std::vector<size_t> sizes = {
      rows0 * C0 * es, rows0 * hcols * es, rows0 * 3 * C0 * es, rows0 * C0 * es, rows0 * 4 * C0 * es,           // x h qkv ao mlp
      B * Hd * es, B * Hd * es, B * c.n_heads * c.head_dim * es, B * c.n_heads * c.head_dim * es, B * Hd * es,  // dx dn dq da dcross
      B * Hd * es, B * QW * es, B * c.inter * es, B * Hd * es,                                                  // dres dqkv dm dh
      B * Lk * c.enc_hidden * es,                                                                               // enc
      B * sizeof(int), B * 10 * sizeof(long long), B * sizeof(int), 256, B * 6 * sizeof(float), B,              // slot tok pos base bbox done
  };
  for (int k = 0; k < 4; ++k) sizes.push_back(B * (k < c.n_out_heads ? c.head_n[k] : 1) * sizeof(float));
  for (int l = 0; l < c.dec_layers; ++l) { sizes.push_back(B * Lk * KVW * es); sizes.push_back(cache); sizes.push_back(cache); }
  size_t total = 0;
  for (size_t s : sizes) total += al256(s);
  if (cudaMalloc(&e->arena, total) != cudaSuccess) {
    cudaGetLastError();
    set_error("sb_layout_create: cudaMalloc(%zu) failed", total);
    delete e;
    return -5;
  }
  cudaMemset(e->arena, 0, total);
  e->arena_bytes = total;
  uint8_t* p = e->arena;
  size_t si = 0;
  auto take = [&]() { void* r = p; p += al256(sizes[si++]); return r; };
  e->x = take(); e->h = take(); e->qkv = take(); e->ao = take(); e->mlp = take();
  e->dx = take(); e->dn = take(); e->dq = take(); e->da = take(); e->dcross = take();
  e->dres = take(); e->dqkv = take(); e->dm = take(); e->dh = take();
  e->enc = take();
  e->slot = static_cast<int*>(take()); e->tok = static_cast<long long*>(take()); e->pos = static_cast<int*>(take());
  e->base = static_cast<int*>(take()); e->out_bbox = static_cast<float*>(take()); e->done = static_cast<unsigned char*>(take());
  for (int k = 0; k < 4; ++k) e->out_head[k] = static_cast<float*>(take());
  for (int l = 0; l < c.dec_layers; ++l) { e->ckv.push_back(take()); e->kc.push_back(take()); e->vc.push_back(take()); }
  std::vector<int> slots(B);
  for (size_t i = 0; i < B; ++i) slots[i] = static_cast<int>(i);
  cudaMemcpy(e->slot, slots.data(), B * sizeof(int), cudaMemcpyHostToDevice);
  e->Lk = static_cast<int>(Lk);
  *out = e;
  return 0;
}

void sb_layout_destroy(sb_layout_engine* e) {
  if (!e) return;
  if (e->g_group) cudaGraphExecDestroy(e->g_group);
  if (e->g_one) cudaGraphExecDestroy(e->g_one);
  if (e->ev_in) cudaEventDestroy(e->ev_in);
  if (e->ev_out) cudaEventDestroy(e->ev_out);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  if (e->arena) cudaFree(e->arena);
  delete e;
}

size_t sb_layout_workspace_bytes(const sb_layout_engine* e) { return e ? e->arena_bytes : 0; }

int sb_layout_encode(sb_layout_engine* e, const void* pixels, int pixels_f32, int batch, void* enc_out, void* stream) {
  if (!e || !pixels || !enc_out) { set_error("sb_layout_encode: null argument"); return -1; }
  if (batch <= 0 || batch > e->c.max_batch) { set_error("sb_layout_encode: batch %d outside 1..%d", batch, e->c.max_batch); return -2; }
  return run_encoder(e, pixels, pixels_f32, batch, enc_out, static_cast<cudaStream_t>(stream));
}

int sb_layout_decode(sb_layout_engine* e, const void* enc, int batch, const long long* prompt, int q_len, int n_steps,
                     long long* hist_tok, float* hist_bbox, float* const* hist_heads, unsigned char* hist_done, int use_graph,
                     void* stream) {
  if (!e || !enc || !prompt || !hist_tok) { set_error("sb_layout_decode: null argument"); return -1; }
  const sb_layout_config& c = e->c;
  if (batch <= 0 || batch > c.max_batch) { set_error("sb_layout_decode: batch %d outside 1..%d", batch, c.max_batch); return -2; }
  if (q_len < 1 || n_steps < 1 || q_len - 1 + n_steps > c.s_max) { set_error("sb_layout_decode: prompt + steps exceed s_max %d", c.s_max); return -3; }
  cudaStream_t caller = static_cast<cudaStream_t>(stream);
  cudaStream_t st = caller;
  const bool hop = use_graph && (caller == nullptr || caller == cudaStreamLegacy || caller == cudaStreamPerThread);
  if (hop) {
    if (!e->own_stream) {
      if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        set_error("sb_layout_decode: could not create the engine stream");
        return -8;
      }
    }
    cudaEventRecord(e->ev_in, caller);
    cudaStreamWaitEvent(e->own_stream, e->ev_in, 0);
    st = e->own_stream;
  }
  struct Rejoin {
    bool on; cudaStream_t own, caller; cudaEvent_t ev;
    ~Rejoin() { if (on) { cudaEventRecord(ev, own); cudaStreamWaitEvent(caller, ev, 0); } }
  } rejoin{hop, e->own_stream, caller, e->ev_out};

  const int B = batch, ncol = e->ncol();
  const size_t es = e->esz;
  // encoder states -> persistent buffer; cross-attention K/V of every layer (SuryaADETRDecoderSdpaCrossAttention caches them at
  // the first call, adetr/decoder.py:150-194)
  if (cudaMemcpyAsync(e->enc, enc, static_cast<size_t>(B) * e->Lk * c.enc_hidden * es, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
    cudaGetLastError(); set_error("sb_layout_decode: encoder copy failed"); return -4;
  }
  const int KVW = 2 * c.n_kv * c.head_dim;
  for (int l = 0; l < c.dec_layers; ++l)
    CK(lin(e, e->enc, c.enc_hidden, e->WD(l, SB_LWD_CKV_W), c.enc_hidden, e->ckv[l], KVW, B * e->Lk, KVW, c.enc_hidden, nullptr, nullptr, 0,
           ACT_NONE, 0, st));
  // prompt: all but the last token, position by position (identical to the batched prefill under the causal mask)
  std::vector<int> hpos(B);
  for (int j = 0; j < q_len; ++j) {
    // gather column j of the prompt [B, q_len, ncol] into the token buffer
    if (cudaMemcpy2DAsync(e->tok, ncol * sizeof(long long), prompt + static_cast<size_t>(j) * ncol, static_cast<size_t>(q_len) * ncol * sizeof(long long),
                          ncol * sizeof(long long), B, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      cudaGetLastError(); set_error("sb_layout_decode: prompt copy failed"); return -5;
    }
    for (int b = 0; b < B; ++b) hpos[b] = j;
    if (cudaMemcpyAsync(e->pos, hpos.data(), B * sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess) {
      cudaGetLastError(); set_error("sb_layout_decode: position copy failed"); return -5;
    }
    if (j == q_len - 1) break;
    cudaStreamSynchronize(st);     // hpos is reused by the next iteration (pageable source)
    CK(run_token(e, e->tok, e->pos, B, st));
  }
  const int base = q_len - 1;
  if (cudaMemcpyAsync(e->base, &base, sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess) {
    cudaGetLastError(); set_error("sb_layout_decode: base copy failed"); return -5;
  }
  cudaStreamSynchronize(st);       // host sources above are stack / vector memory

  auto body = [&](cudaStream_t s) -> int {
    CK(run_token(e, e->tok, e->pos, B, s));
    return run_next(e, B, n_steps, hist_tok, hist_bbox, hist_heads, hist_done, s);
  };
  CK(body(st));                    // first step eagerly (sets kernel attributes before any capture)
  int left = n_steps - 1;
  if (left <= 0) return 0;
  if (!use_graph) {
    for (int i = 0; i < left; ++i) CK(body(st));
    return 0;
  }
  const void* key[8] = {hist_tok, hist_bbox, hist_done, hist_heads ? hist_heads[0] : nullptr,
                        reinterpret_cast<const void*>(static_cast<uintptr_t>(n_steps)), nullptr, nullptr, nullptr};
  const bool hit = e->g_one && e->graph_batch == B && std::memcmp(key, e->graph_key, sizeof(key)) == 0;
  if (!hit) {
    if (e->g_group) { cudaGraphExecDestroy(e->g_group); e->g_group = nullptr; }
    if (e->g_one) { cudaGraphExecDestroy(e->g_one); e->g_one = nullptr; }
    for (int which = 0; which < 2; ++which) {
      const int k = which == 0 ? GRAPH_GROUP : 1;
      cudaGraph_t graph = nullptr;
      cudaError_t cb = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
      if (cb != cudaSuccess) { cudaGetLastError(); set_error("sb_layout_decode: capture begin failed: %s", cudaGetErrorString(cb)); return -6; }
      const long long before = launch_count();
      int rc = 0;
      for (int i = 0; i < k && !rc; ++i) rc = body(st);
      e->launches_per_step = (launch_count() - before) / k;
      count_launches(before - launch_count());   // captured, not executed
      cudaError_t ce = cudaStreamEndCapture(st, &graph);
      if (rc || ce != cudaSuccess) { cudaGetLastError(); set_error("sb_layout_decode: capture failed: rc=%d %s", rc, cudaGetErrorString(ce)); return -6; }
      ce = cudaGraphInstantiate(which == 0 ? &e->g_group : &e->g_one, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) { cudaGetLastError(); set_error("sb_layout_decode: instantiate failed: %s", cudaGetErrorString(ce)); return -7; }
    }
    e->graph_batch = B;
    std::memcpy(e->graph_key, key, sizeof(key));
  }
  const long long per_step = e->launches_per_step;   // kernels per step as counted while capturing (launch accounting)
  while (left >= GRAPH_GROUP) {
    if (cudaGraphLaunch(e->g_group, st) != cudaSuccess) { cudaGetLastError(); set_error("sb_layout_decode: graph launch failed"); return -9; }
    count_launches(per_step * GRAPH_GROUP);
    left -= GRAPH_GROUP;
  }
  for (; left > 0; --left) {
    if (cudaGraphLaunch(e->g_one, st) != cudaSuccess) { cudaGetLastError(); set_error("sb_layout_decode: graph launch failed"); return -9; }
    count_launches(per_step);
  }
  return 0;
}

}  // extern "C"
