// surya_b200 — recognition engine: vision tower + decoder prefill + decode steps as one C++ object.
//
// Stands in for SuryaModel.forward (surya/common/surya/__init__.py:274-338) under RecognitionPredictor.prefill /
// decode (surya/recognition/__init__.py:326-352, 354-471).  The layer loops live here (not in Python) so that a
// decode step is ~90 back-to-back kernel launches on one stream, capturable into a CUDA graph and replayed with
// token feedback / position increment done on the device (sb_rec_decode_steps).
#include "../../include/surya_b200.h"
#include "ops.cuh"
#include "sb_ptx.cuh"

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace sb;

#define CK(x)                    \
  do {                           \
    int rc_ = (x);               \
    if (rc_) return rc_;         \
  } while (0)

struct sb_rec_engine {
  sb_rec_config c;
  std::vector<const void*> w;
  size_t esz = 2;
  // workspaces (device)
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  void *x = nullptr, *nbuf = nullptr, *qkv = nullptr, *ao = nullptr, *act = nullptr, *x0 = nullptr, *m1 = nullptr,
       *feat = nullptr, *xl = nullptr, *logits = nullptr;
  void* kcache = nullptr;  // [layer][slot][kvh][s_max][d]
  void* vcache = nullptr;
  // staging for the device-side decode loop
  long long* st_tok = nullptr; float* st_score = nullptr; long long* st_bbox = nullptr; unsigned char* st_done = nullptr;
  long long* st_next = nullptr; int* st_step = nullptr;
  int qkv_w_enc = 0, qkv_w_dec = 0;
  // CUDA graph cache for decode_steps
  cudaGraphExec_t graph_exec = nullptr;
  int graph_batch = 0;
  long long graph_nodes = 0;
  // the legacy default stream cannot be captured: decode_steps hops onto an engine-owned stream, ordered by events
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  // decode chains: the batch is cut into row groups whose (strictly sequential, latency-bound) kernel chains run side by side
  // on forked streams inside one step / one CUDA graph
  static constexpr int MAX_CHAINS = 4;
  int n_chains = 1;
  cudaStream_t chain_stream[MAX_CHAINS] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_CHAINS] = {nullptr, nullptr, nullptr, nullptr};
  const void* graph_key[8] = {nullptr};

  const void* W(int idx) const { return w[idx]; }
  const void* WE(int layer, int k) const { return w[SB_RW_ENC_BASE + layer * SB_RWE_STRIDE + k]; }
  const void* WD(int layer, int k) const {
    return w[SB_RW_ENC_BASE + c.enc_depth * SB_RWE_STRIDE + layer * SB_RWD_STRIDE + k];
  }
  void* kc(int layer) const {
    return static_cast<uint8_t*>(kcache) + static_cast<size_t>(layer) * c.max_slots * c.dec_kv_heads * c.s_max * c.dec_head_dim * esz;
  }
  void* vc(int layer) const {
    return static_cast<uint8_t*>(vcache) + static_cast<size_t>(layer) * c.max_slots * c.dec_kv_heads * c.s_max * c.dec_head_dim * esz;
  }
};

static size_t al256(size_t v) { return (v + 255) & ~size_t(255); }

static int linear(const sb_rec_engine* e, const void* A, int lda, const void* Wt, int ldw, void* C, int ldc, int M, int N,
                  int K, const void* bias_f32, const void* residual, int ldr, int act, int swiglu, cudaStream_t st,
                  int allow_splitk = 0) {
  GemmArgs a;
  a.dtype = e->c.dtype;
  a.A = A; a.lda = lda; a.W = Wt; a.ldw = ldw; a.C = C; a.ldc = ldc;
  a.M = M; a.N = N; a.K = K;
  a.bias = static_cast<const float*>(bias_f32);
  a.residual = residual; a.ldr = ldr; a.act = act; a.swiglu = swiglu;
  a.w_constant = 1;   // engine weights are never written after packing
  a.allow_splitk = allow_splitk;
  return gemm_launch(a, st);
}

// ------------------------------------------------------------------------------------------------ vision tower
static int run_vision(sb_rec_engine* e, const void* tiles, int tiles_f32, int n, const int* perm, const int* pos_rc,
                      const int* win_start, const int* win_len, int n_win, int max_win, const int* img_start,
                      const int* img_len, int n_img, int max_img, cudaStream_t st) {
  const sb_rec_config& c = e->c;
  const int H = c.enc_hidden, nh = c.enc_heads, hd = H / nh;
  const int dt = c.dtype;
  CK(gather_pad_rows(dt, tiles, tiles_f32, c.patch_dim, perm, e->x0, c.patch_dim_pad, n, c.patch_dim, c.patch_dim_pad, st));
  CK(linear(e, e->x0, c.patch_dim_pad, e->W(SB_RW_PATCH_W), c.patch_dim_pad, e->x, H, n, H, c.patch_dim_pad, nullptr,
            nullptr, 0, ACT_NONE, 0, st));
  const int ldq = 3 * H;
  uint8_t* qkv8 = static_cast<uint8_t*>(e->qkv);
  for (int l = 0; l < c.enc_depth; ++l) {
    CK(rmsnorm(dt, e->x, H, e->WE(l, SB_RWE_NORM1), e->nbuf, H, n, H, 1e-6f, nullptr, st));
    CK(linear(e, e->nbuf, H, e->WE(l, SB_RWE_QKV_W), H, e->qkv, ldq, n, 3 * H, H, e->WE(l, SB_RWE_QKV_B), nullptr, 0,
              ACT_NONE, 0, st));
    CK(rope_vision(dt, e->qkv, ldq, pos_rc, static_cast<const float*>(e->W(SB_RW_ENC_INV_FREQ)), n, nh, hd, st));
    AttnArgs a;
    a.dtype = dt;
    a.q = qkv8; a.ldq = ldq;
    a.k = qkv8 + static_cast<size_t>(H) * e->esz; a.ldk = ldq;
    a.v = qkv8 + static_cast<size_t>(2 * H) * e->esz; a.ldv = ldq;
    a.out = e->ao; a.ldo = H;
    const bool full = (c.fullatt_mask >> l) & 1u;
    a.seq_start = full ? img_start : win_start;
    a.seq_len = full ? img_len : win_len;
    a.n_seq = full ? n_img : n_win;
    a.max_len = full ? max_img : max_win;
    a.n_heads = nh; a.n_kv_heads = nh; a.head_dim = hd; a.causal = 0;
    a.scale = 1.0f / sqrtf(static_cast<float>(hd));
    CK(attn_varlen(a, st));
    CK(linear(e, e->ao, H, e->WE(l, SB_RWE_PROJ_W), H, e->x, H, n, H, H, e->WE(l, SB_RWE_PROJ_B), e->x, H, ACT_NONE, 0, st));
    CK(rmsnorm(dt, e->x, H, e->WE(l, SB_RWE_NORM2), e->nbuf, H, n, H, 1e-6f, nullptr, st));
    CK(linear(e, e->nbuf, H, e->WE(l, SB_RWE_GU_W), H, e->act, c.enc_inter_pad, n, 2 * c.enc_inter_pad, H,
              e->WE(l, SB_RWE_GU_B), nullptr, 0, ACT_SILU, 1, st));
    CK(linear(e, e->act, c.enc_inter_pad, e->WE(l, SB_RWE_DOWN_W), c.enc_inter_pad, e->x, H, n, H, c.enc_inter_pad,
              e->WE(l, SB_RWE_DOWN_B), e->x, H, ACT_NONE, 0, st));
  }
  // merger: RMSNorm -> view [n/unit, unit*H] -> Linear+GELU -> Linear
  const int mu = c.merge_unit, MH = mu * H, nm = n / mu;
  CK(rmsnorm(dt, e->x, H, e->W(SB_RW_MERGER_LN), e->nbuf, H, n, H, 1e-6f, nullptr, st));
  CK(linear(e, e->nbuf, MH, e->W(SB_RW_MERGER_W0), MH, e->m1, MH, nm, MH, MH, e->W(SB_RW_MERGER_B0), nullptr, 0,
            ACT_GELU_ERF, 0, st));
  CK(linear(e, e->m1, MH, e->W(SB_RW_MERGER_W2), MH, e->feat, c.enc_out_hidden, nm, c.enc_out_hidden, MH,
            e->W(SB_RW_MERGER_B2), nullptr, 0, ACT_NONE, 0, st));
  return 0;
}

// ------------------------------------------------------------------------------------------------ decoder pieces
static int dec_mlp_block(sb_rec_engine* e, int l, int rows, cudaStream_t st) {
  const sb_rec_config& c = e->c;
  const int D = c.dec_hidden;
  CK(rmsnorm(c.dtype, e->x, D, e->WD(l, SB_RWD_POST_NORM), e->nbuf, D, rows, D, c.rms_eps, nullptr, st));
  CK(linear(e, e->nbuf, D, e->WD(l, SB_RWD_GU_W), D, e->act, c.dec_inter_pad, rows, 2 * c.dec_inter_pad, D, nullptr,
            nullptr, 0, ACT_SILU, 1, st));
  CK(linear(e, e->act, c.dec_inter_pad, e->WD(l, SB_RWD_DOWN_W), c.dec_inter_pad, e->x, D, rows, D, c.dec_inter_pad,
            nullptr, e->x, D, ACT_NONE, 0, st));
  return 0;
}

static int run_heads(sb_rec_engine* e, const void* hidden, int rows, void* logits_out, long long* tok, float* score,
                     long long* bbox, float* bbox_sig, unsigned char* done, long long* next_ids, cudaStream_t st) {
  const sb_rec_config& c = e->c;
  const int D = c.dec_hidden;
  void* lg = logits_out ? logits_out : e->logits;
  CK(linear(e, hidden, D, e->W(SB_RW_EMBED), D, lg, c.vocab, rows, c.vocab, D, e->W(SB_RW_LM_BIAS), nullptr, 0, ACT_NONE,
            0, st));
  if (tok || score || done || next_ids) {
    CK(argmax_score(c.dtype, lg, c.vocab, rows, c.vocab, tok ? tok : e->st_tok, score ? score : e->st_score, done,
                    next_ids, c.eos_id, c.pad_id, st));
  }
  if (bbox || bbox_sig) {
    CK(small_head(c.dtype, hidden, D, e->W(SB_RW_BBOX_W), e->W(SB_RW_BBOX_B), rows, D, 6, 1, bbox_sig, bbox, c.bbox_size, st));
  }
  return 0;
}

static int run_decoder_prefill(sb_rec_engine* e, const long long* ids, int n_tok, const int* feat_row, const int* hidx,
                               const int* widx, const int* tok_pos, const int* tok_slot, const int* seq_start,
                               const int* seq_len, int n_seq, int max_len, const int* last_tok, void* logits,
                               long long* tok, float* score, long long* bbox, float* bbox_sig, unsigned char* done,
                               long long* next_ids, cudaStream_t st) {
  const sb_rec_config& c = e->c;
  const int D = c.dec_hidden, nh = c.dec_heads, nkv = c.dec_kv_heads, hd = c.dec_head_dim;
  const int Q = (nh + 2 * nkv) * hd;
  const int dt = c.dtype;
  CK(embed_splice(dt, ids, feat_row, hidx, widx, e->W(SB_RW_EMBED), e->feat, c.enc_out_hidden, e->W(SB_RW_H_EMBED),
                  e->W(SB_RW_W_EMBED), e->x, D, n_tok, D, st));
  uint8_t* qkv8 = static_cast<uint8_t*>(e->qkv);
  for (int l = 0; l < c.dec_layers; ++l) {
    CK(rmsnorm(dt, e->x, D, e->WD(l, SB_RWD_IN_NORM), e->nbuf, D, n_tok, D, c.rms_eps, nullptr, st));
    CK(linear(e, e->nbuf, D, e->WD(l, SB_RWD_QKV_W), D, e->qkv, Q, n_tok, Q, D, e->WD(l, SB_RWD_QKV_B), nullptr, 0,
              ACT_NONE, 0, st));
    CK(rope_kv_append(dt, e->qkv, Q, tok_pos, tok_slot, static_cast<const float*>(e->W(SB_RW_DEC_INV_FREQ)), e->kc(l),
                      e->vc(l), n_tok, nh, nkv, hd, c.s_max, st));
    AttnArgs a;
    a.dtype = dt;
    a.q = qkv8; a.ldq = Q;
    a.k = qkv8 + static_cast<size_t>(nh * hd) * e->esz; a.ldk = Q;
    a.v = qkv8 + static_cast<size_t>((nh + nkv) * hd) * e->esz; a.ldv = Q;
    a.out = e->ao; a.ldo = nh * hd;
    a.seq_start = seq_start; a.seq_len = seq_len; a.n_seq = n_seq; a.max_len = max_len;
    a.n_heads = nh; a.n_kv_heads = nkv; a.head_dim = hd; a.causal = 1;
    a.scale = 1.0f / sqrtf(static_cast<float>(hd));
    CK(attn_varlen(a, st));
    CK(linear(e, e->ao, nh * hd, e->WD(l, SB_RWD_O_W), nh * hd, e->x, D, n_tok, D, nh * hd, nullptr, e->x, D, ACT_NONE, 0, st));
    CK(dec_mlp_block(e, l, n_tok, st));
  }
  CK(rmsnorm(dt, e->x, D, e->W(SB_RW_DEC_NORM), e->xl, D, n_seq, D, c.rms_eps, last_tok, st));
  return run_heads(e, e->xl, n_seq, logits, tok, score, bbox, bbox_sig, done, next_ids, st);
}

// One greedy decode step for batch rows [r0, r0 + B): every workspace is row-major, so a row group is a pointer offset.
static int run_decode_step(sb_rec_engine* e, const long long* ids, const int* slot, const int* pos, int r0, int B, void* logits,
                           long long* tok, float* score, long long* bbox, float* bbox_sig, unsigned char* done,
                           long long* next_ids, cudaStream_t st) {
  const sb_rec_config& c = e->c;
  const int D = c.dec_hidden, nh = c.dec_heads, nkv = c.dec_kv_heads, hd = c.dec_head_dim;
  const int Q = (nh + 2 * nkv) * hd;
  const int dt = c.dtype;
  auto rows = [&](void* base, size_t width) { return static_cast<void*>(static_cast<uint8_t*>(base) + static_cast<size_t>(r0) * width * e->esz); };
  void* x = rows(e->x, D);
  void* nbuf = rows(e->nbuf, D);
  void* qkv = rows(e->qkv, Q);
  void* ao = rows(e->ao, nh * hd);
  void* act = rows(e->act, c.dec_inter_pad);
  void* xl = rows(e->xl, D);
  ids += r0; slot += r0; pos += r0;
  CK(embed_rows(dt, ids, e->W(SB_RW_EMBED), x, D, B, D, st));
  for (int l = 0; l < c.dec_layers; ++l) {
    CK(rmsnorm(dt, x, D, e->WD(l, SB_RWD_IN_NORM), nbuf, D, B, D, c.rms_eps, nullptr, st));
    CK(linear(e, nbuf, D, e->WD(l, SB_RWD_QKV_W), D, qkv, Q, B, Q, D, e->WD(l, SB_RWD_QKV_B), nullptr, 0, ACT_NONE, 0, st));
    DecodeAttnArgs a;
    a.dtype = dt; a.qkv = qkv; a.ld = Q; a.kcache = e->kc(l); a.vcache = e->vc(l); a.slot = slot; a.pos = pos;
    a.inv_freq = static_cast<const float*>(e->W(SB_RW_DEC_INV_FREQ));
    a.out = ao; a.ldo = nh * hd; a.batch = B; a.n_heads = nh; a.n_kv_heads = nkv; a.head_dim = hd; a.s_max = c.s_max;
    a.scale = 1.0f / sqrtf(static_cast<float>(hd));
    CK(decode_attn(a, st));
    CK(linear(e, ao, nh * hd, e->WD(l, SB_RWD_O_W), nh * hd, x, D, B, D, nh * hd, nullptr, x, D, ACT_NONE, 0, st));
    CK(rmsnorm(dt, x, D, e->WD(l, SB_RWD_POST_NORM), nbuf, D, B, D, c.rms_eps, nullptr, st));
    CK(linear(e, nbuf, D, e->WD(l, SB_RWD_GU_W), D, act, c.dec_inter_pad, B, 2 * c.dec_inter_pad, D, nullptr, nullptr, 0,
              ACT_SILU, 1, st));
    CK(linear(e, act, c.dec_inter_pad, e->WD(l, SB_RWD_DOWN_W), c.dec_inter_pad, x, D, B, D, c.dec_inter_pad, nullptr, x, D,
              ACT_NONE, 0, st, /*allow_splitk=*/1));
  }
  CK(rmsnorm(dt, x, D, e->W(SB_RW_DEC_NORM), xl, D, B, D, c.rms_eps, nullptr, st));
  void* lg = logits ? static_cast<void*>(static_cast<uint8_t*>(logits) + static_cast<size_t>(r0) * c.vocab * e->esz)
                    : rows(e->logits, c.vocab);
  return run_heads(e, xl, B, lg, tok ? tok + r0 : nullptr, score ? score + r0 : nullptr, bbox ? bbox + static_cast<size_t>(r0) * 6 : nullptr,
                   bbox_sig ? bbox_sig + static_cast<size_t>(r0) * 6 : nullptr, done ? done + r0 : nullptr,
                   next_ids ? next_ids + r0 : nullptr, st);
}

// The whole batch as n_chains row groups on forked streams (joined back into `st`); works eagerly and under stream capture.
static int run_decode_step_chained(sb_rec_engine* e, const long long* ids, const int* slot, const int* pos, int B, void* logits,
                                   long long* tok, float* score, long long* bbox, float* bbox_sig, unsigned char* done,
                                   long long* next_ids, cudaStream_t st) {
  int nc = e->n_chains;
  if (nc > sb_rec_engine::MAX_CHAINS) nc = sb_rec_engine::MAX_CHAINS;
  while (nc > 1 && B / nc < 32) --nc;            // do not cut below one warp-row group
  if (nc <= 1) return run_decode_step(e, ids, slot, pos, 0, B, logits, tok, score, bbox, bbox_sig, done, next_ids, st);
  if (!e->ev_fork) {
    if (cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); set_error("chain event"); return -20; }
    for (int i = 1; i < sb_rec_engine::MAX_CHAINS; ++i) {
      if (cudaStreamCreateWithFlags(&e->chain_stream[i], cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&e->ev_join[i], cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError(); set_error("chain stream"); return -20;
      }
    }
  }
  const int per = ((B + nc - 1) / nc + 7) & ~7;   // row groups in multiples of 8
  if (cudaEventRecord(e->ev_fork, st) != cudaSuccess) { cudaGetLastError(); set_error("chain fork"); return -21; }
  for (int i = 0; i < nc; ++i) {
    const int r0 = i * per;
    const int n = (r0 + per <= B) ? per : B - r0;
    if (n <= 0) break;
    cudaStream_t s = i == 0 ? st : e->chain_stream[i];
    if (i) cudaStreamWaitEvent(s, e->ev_fork, 0);
    CK(run_decode_step(e, ids, slot, pos, r0, n, logits, tok, score, bbox, bbox_sig, done, next_ids, s));
    if (i) {
      cudaEventRecord(e->ev_join[i], s);
      cudaStreamWaitEvent(st, e->ev_join[i], 0);
    }
  }
  return 0;
}

// Device-side bookkeeping between two greedy steps: history append, token feedback, position increment.
__global__ void record_step_kernel(int* step, int B, const long long* tok, const float* score, const long long* bbox,
                                   const unsigned char* done, const long long* next, long long* tok_hist,
                                   float* score_hist, long long* bbox_hist, unsigned char* done_hist,
                                   long long* ids_io, int* pos_io) {
  pdl_trigger();
  pdl_wait();
  const int s = *step;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    size_t o = static_cast<size_t>(s) * B + b;
    if (tok_hist) tok_hist[o] = tok[b];
    if (score_hist) score_hist[o] = score[b];
    if (done_hist) done_hist[o] = done[b];
    if (bbox_hist)
      for (int j = 0; j < 6; ++j) bbox_hist[o * 6 + j] = bbox[b * 6 + j];
    ids_io[b] = next[b];
    pos_io[b] += 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) *step = s + 1;
}

extern "C" {

int sb_rec_create(const sb_rec_config* cfg, const void* const* weights, int n_weights, sb_rec_engine** out) {
  if (!cfg || !weights || !out) { set_error("sb_rec_create: null argument"); return -1; }
  const int need = SB_RW_ENC_BASE + cfg->enc_depth * SB_RWE_STRIDE + cfg->dec_layers * SB_RWD_STRIDE;
  if (n_weights != need) { set_error("sb_rec_create: expected %d weight pointers, got %d", need, n_weights); return -2; }
  if (cfg->dec_hidden != cfg->enc_out_hidden) { set_error("decoder hidden must equal encoder out_hidden"); return -3; }
  if (cfg->patch_dim_pad % 8 || cfg->enc_inter_pad % 8 || cfg->dec_inter_pad % 8 || cfg->enc_hidden % 8 ||
      cfg->dec_hidden % 8 || cfg->vocab % 8) {
    set_error("sb_rec_create: padded dims and vocab must be multiples of 8 (16-byte TMA pitch rule)");
    return -4;
  }
  for (int i = 0; i < n_weights; ++i)
    if (!weights[i]) { set_error("sb_rec_create: weight pointer %d is null", i); return -5; }
  auto* e = new sb_rec_engine();
  e->c = *cfg;
  if (const char* ev = getenv("SB_DECODE_CHAINS")) {
    const int v = atoi(ev);
    if (v >= 1 && v <= sb_rec_engine::MAX_CHAINS) e->n_chains = v;
  }
  e->w.assign(weights, weights + n_weights);
  const sb_rec_config& c = e->c;
  const size_t es = e->esz;
  const size_t R = static_cast<size_t>(c.max_patches > c.max_tokens ? c.max_patches : c.max_tokens);
  const size_t Hm = c.enc_hidden > c.dec_hidden ? c.enc_hidden : c.dec_hidden;
  const int Q = (c.dec_heads + 2 * c.dec_kv_heads) * c.dec_head_dim;
  const size_t qkvw = (size_t)(3 * c.enc_hidden > Q ? 3 * c.enc_hidden : Q);
  const size_t aow = (size_t)(c.enc_hidden > c.dec_heads * c.dec_head_dim ? c.enc_hidden : c.dec_heads * c.dec_head_dim);
  const size_t actw = (size_t)(c.enc_inter_pad > c.dec_inter_pad ? c.enc_inter_pad : c.dec_inter_pad);
  const size_t nm = (size_t)c.max_patches / c.merge_unit + 1;
  const size_t kv_bytes = (size_t)c.dec_layers * c.max_slots * c.dec_kv_heads * c.s_max * c.dec_head_dim * es;
  const size_t rows_out = (size_t)(c.max_seqs > c.max_slots ? c.max_seqs : c.max_slots);
  size_t sizes[] = {
      R * Hm * es, R * Hm * es, R * qkvw * es, R * aow * es, R * actw * es,           // x nbuf qkv ao act
      (size_t)c.max_patches * c.patch_dim_pad * es, nm * c.merge_unit * c.enc_hidden * es,  // x0 m1
      nm * c.enc_out_hidden * es, rows_out * c.dec_hidden * es, rows_out * c.vocab * es,   // feat xl logits
      kv_bytes, kv_bytes,
      rows_out * 8, rows_out * 4, rows_out * 6 * 8, rows_out, rows_out * 8, 256};
  size_t total = 0;
  for (size_t s : sizes) total += al256(s);
  cudaError_t ce = cudaMalloc(&e->arena, total);
  if (ce != cudaSuccess) {
    set_error("sb_rec_create: cudaMalloc(%zu MiB) failed: %s", total >> 20, cudaGetErrorString(ce));
    delete e;
    return -6;
  }
  e->arena_bytes = total;
  uint8_t* p = e->arena;
  void** slots[] = {&e->x, &e->nbuf, &e->qkv, &e->ao, &e->act, &e->x0, &e->m1, &e->feat, &e->xl, &e->logits,
                    &e->kcache, &e->vcache, (void**)&e->st_tok, (void**)&e->st_score, (void**)&e->st_bbox,
                    (void**)&e->st_done, (void**)&e->st_next, (void**)&e->st_step};
  for (size_t i = 0; i < sizeof(sizes) / sizeof(sizes[0]); ++i) {
    *slots[i] = p;
    p += al256(sizes[i]);
  }
  cudaMemset(e->arena, 0, total);
  *out = e;
  return 0;
}

void sb_rec_destroy(sb_rec_engine* e) {
  if (!e) return;
  if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
  if (e->ev_in) cudaEventDestroy(e->ev_in);
  if (e->ev_out) cudaEventDestroy(e->ev_out);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  for (int i = 1; i < sb_rec_engine::MAX_CHAINS; ++i) {
    if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
    if (e->chain_stream[i]) cudaStreamDestroy(e->chain_stream[i]);
  }
  if (e->arena) cudaFree(e->arena);
  delete e;
}

size_t sb_rec_workspace_bytes(const sb_rec_engine* e) { return e ? e->arena_bytes : 0; }

int sb_rec_prefill(sb_rec_engine* e, const void* tiles, int tiles_f32, int n_patches, const int* patch_perm,
                   const int* patch_pos_rc, const int* win_start, const int* win_len, int n_win, int max_win_len,
                   const int* img_start, const int* img_len, int n_img, int max_img_len, const long long* input_ids,
                   int n_tok, const int* tok_feat_row, const int* tok_hidx, const int* tok_widx, const int* tok_pos,
                   const int* tok_slot, const int* seq_start, const int* seq_len, int n_seq, int max_seq_len,
                   const int* last_tok, void* logits, long long* tok, float* score, long long* bbox, float* bbox_sig,
                   unsigned char* done, long long* next_ids, void* stream) {
  if (!e) { set_error("sb_rec_prefill: null engine"); return -1; }
  const sb_rec_config& c = e->c;
  if (n_patches > c.max_patches || n_tok > c.max_tokens || n_seq > c.max_seqs) {
    set_error("sb_rec_prefill: batch exceeds engine capacity (patches %d/%d tokens %d/%d seqs %d/%d)", n_patches,
              c.max_patches, n_tok, c.max_tokens, n_seq, c.max_seqs);
    return -2;
  }
  if (n_patches % c.merge_unit) { set_error("sb_rec_prefill: patch count must be a multiple of the merge unit"); return -3; }
  if (max_seq_len > c.s_max) { set_error("sb_rec_prefill: sequence longer than s_max"); return -4; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_patches > 0)
    CK(run_vision(e, tiles, tiles_f32, n_patches, patch_perm, patch_pos_rc, win_start, win_len, n_win, max_win_len,
                  img_start, img_len, n_img, max_img_len, st));
  return run_decoder_prefill(e, input_ids, n_tok, tok_feat_row, tok_hidx, tok_widx, tok_pos, tok_slot, seq_start,
                             seq_len, n_seq, max_seq_len, last_tok, logits, tok, score, bbox, bbox_sig, done, next_ids, st);
}

int sb_rec_decode(sb_rec_engine* e, const long long* input_ids, const int* slot, const int* pos, int batch, void* logits,
                  long long* tok, float* score, long long* bbox, float* bbox_sig, unsigned char* done,
                  long long* next_ids, void* stream) {
  if (!e) { set_error("sb_rec_decode: null engine"); return -1; }
  if (batch > e->c.max_slots || batch > e->c.max_tokens) { set_error("sb_rec_decode: batch %d exceeds capacity", batch); return -2; }
  return run_decode_step(e, input_ids, slot, pos, 0, batch, logits, tok, score, bbox, bbox_sig, done, next_ids,
                         static_cast<cudaStream_t>(stream));
}

int sb_rec_decode_steps(sb_rec_engine* e, long long* ids_io, const int* slot, int* pos_io, int batch, int n_steps,
                        long long* tok_hist, float* score_hist, long long* bbox_hist, unsigned char* done_hist,
                        int use_graph, void* stream) {
  if (!e) { set_error("sb_rec_decode_steps: null engine"); return -1; }
  if (batch > e->c.max_slots || batch > e->c.max_tokens) { set_error("sb_rec_decode_steps: batch exceeds capacity"); return -2; }
  cudaStream_t caller = static_cast<cudaStream_t>(stream);
  cudaStream_t st = caller;
  const bool hop = use_graph && (caller == nullptr || caller == cudaStreamLegacy || caller == cudaStreamPerThread);
  if (hop) {
    if (!e->own_stream) {
      if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        set_error("sb_rec_decode_steps: could not create the engine stream");
        return -8;
      }
    }
    cudaEventRecord(e->ev_in, caller);
    cudaStreamWaitEvent(e->own_stream, e->ev_in, 0);
    st = e->own_stream;
  }
  struct Rejoin {  // whatever happens below, order the caller's stream after the engine stream
    bool on; cudaStream_t own, caller; cudaEvent_t ev;
    ~Rejoin() { if (on) { cudaEventRecord(ev, own); cudaStreamWaitEvent(caller, ev, 0); } }
  } rejoin{hop, e->own_stream, caller, e->ev_out};
  if (cudaMemsetAsync(e->st_step, 0, sizeof(int), st) != cudaSuccess) { cudaGetLastError(); set_error("memset failed"); return -3; }
  auto one_step = [&](cudaStream_t s) -> int {
    CK(run_decode_step_chained(e, ids_io, slot, pos_io, batch, nullptr, e->st_tok, e->st_score, e->st_bbox, nullptr,
                               e->st_done, e->st_next, s));
    launch_pdl(record_step_kernel, dim3(1), dim3(256), 0, s, e->st_step, batch, (const long long*)e->st_tok,
               (const float*)e->st_score, (const long long*)e->st_bbox, (const unsigned char*)e->st_done,
               (const long long*)e->st_next, tok_hist, score_hist, bbox_hist, done_hist, ids_io, pos_io);
    return launch_ok();
  };
  if (!use_graph) {
    for (int i = 0; i < n_steps; ++i) CK(one_step(st));
    return 0;
  }
  const void* key[8] = {ids_io, slot, pos_io, tok_hist, score_hist, bbox_hist, done_hist, nullptr};
  bool hit = e->graph_exec && e->graph_batch == batch && std::memcmp(key, e->graph_key, sizeof(key)) == 0;
  if (!hit) {
    if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
    // one eager step first: sets kernel attributes (not capturable) and validates the launch parameters
    CK(one_step(st));
    n_steps -= 1;
    if (n_steps <= 0) return 0;
    cudaGraph_t graph = nullptr;
    cudaError_t cb = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    if (cb != cudaSuccess) { cudaGetLastError(); set_error("graph capture begin failed: %s", cudaGetErrorString(cb)); return -4; }
    const long long before = launch_count();
    int rc = one_step(st);
    e->graph_nodes = launch_count() - before;
    count_launches(-e->graph_nodes);  // captured, not executed
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc || ce != cudaSuccess) { cudaGetLastError(); set_error("graph capture failed: rc=%d %s", rc, cudaGetErrorString(ce)); return -5; }
    ce = cudaGraphInstantiate(&e->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { cudaGetLastError(); set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce)); e->graph_exec = nullptr; return -6; }
    e->graph_batch = batch;
    std::memcpy(e->graph_key, key, sizeof(key));
  }
  for (int i = 0; i < n_steps; ++i) {
    if (cudaGraphLaunch(e->graph_exec, st) != cudaSuccess) { cudaGetLastError(); set_error("cudaGraphLaunch failed"); return -7; }
    count_launches(e->graph_nodes);
  }
  return 0;
}

int sb_rec_set_decode_chains(sb_rec_engine* e, int n_chains) {
  if (!e) { set_error("sb_rec_set_decode_chains: null engine"); return -1; }
  if (n_chains < 1 || n_chains > sb_rec_engine::MAX_CHAINS) { set_error("sb_rec_set_decode_chains: 1..%d", sb_rec_engine::MAX_CHAINS); return -2; }
  if (n_chains != e->n_chains && e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
  e->n_chains = n_chains;
  return 0;
}

int sb_rec_debug_copy(sb_rec_engine* e, const char* name, void* dst, size_t bytes, void* stream) {
  if (!e || !name || !dst) { set_error("sb_rec_debug_copy: null argument"); return -1; }
  std::string n(name);
  const void* src = nullptr;
  if (n == "feat") src = e->feat;
  else if (n == "x") src = e->x;
  else if (n == "xl") src = e->xl;
  else if (n == "logits") src = e->logits;
  else if (n == "qkv") src = e->qkv;
  else if (n == "kcache") src = e->kcache;
  else if (n == "vcache") src = e->vcache;
  if (!src) { set_error("sb_rec_debug_copy: unknown workspace '%s'", name); return -2; }
  cudaError_t ce = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream));
  if (ce != cudaSuccess) { set_error("sb_rec_debug_copy: %s", cudaGetErrorString(ce)); return -3; }
  return 0;
}

}  // extern "C"
