/* surya_b200 — C ABI of libsurya_b200.so (sm_100a).
 *
 * The reference (VikParuchuri/surya v0.14.6) has no FFI: its seam is Python attribute access on
 * `predictor.model` (surya/common/predictor.py:20-29).  Every entry point below replaces one group of
 * library calls (cuBLAS / cuDNN / flash-attn / torch eager) that the reference's nn.Modules make on the
 * hot path; the comment on each cites the reference code it stands in for.  The Python mirrors in
 * surya_b200/*.py bind these through ctypes.
 *
 * Conventions: plain pointers and sizes only; `*_dev` / unnamed tensor pointers are DEVICE pointers owned by
 * the caller, row-major; `stream` is a cudaStream_t passed as void*; dtype: 0 = bf16, 1 = fp16.
 * Return 0 on success, negative on error (message via sb_last_error()).  No hidden global state besides the
 * per-thread error string; one engine per GPU/thread.
 */
#ifndef SURYA_B200_H
#define SURYA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_DT_BF16 0
#define SB_DT_F16 1

#define SB_ACT_NONE 0
#define SB_ACT_GELU_ERF 1
#define SB_ACT_SILU 2
#define SB_ACT_HARDSWISH 3
#define SB_ACT_RELU 4
#define SB_ACT_GELU_TANH 5

const char* sb_last_error(void);
int sb_version(void);
/* Number of kernels this library has launched in the calling process (all entry points). */
long long sb_launch_count(void);

/* ------------------------------------------------------------------------------------------------ ops
 * nn.Linear (+bias, activation, residual, SwiGLU) : torch.nn.functional.linear -> cuBLAS in the reference,
 * e.g. surya/common/surya/decoder/__init__.py:37-50,147-158; encoder/__init__.py:22-35,140-141,115-119.
 * C[M,Nout] = epi(A[M,K] @ W[N,K]^T).  bias is fp32 [N] or NULL; residual [M,Nout] or NULL.
 * swiglu=1: W rows interleaved (gate_i, up_i), Nout = N/2, out = act(gate)*up.  out_f32: C is float32. */
int sb_gemm(int dtype, const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
            const float* bias, const void* residual, int ldr, int act, int swiglu, int out_f32, int force_bn,
            void* stream);

/* Qwen2RMSNorm (decoder/__init__.py:241-258, encoder/__init__.py:90-104); src_rows optional gather. */
int sb_rmsnorm(int dtype, const void* x, int ldx, const void* w, void* y, int ldy, int rows, int H, float eps,
               const int* src_rows, void* stream);

/* Row gather + zero pad (+ fp32 -> 16-bit conversion): window re-ordering of patch rows
 * (encoder/__init__.py:622-627) fused with the K-padding the TMA pitch rule needs (588 -> 592). */
int sb_gather_pad_rows(int dtype, const void* src, int src_is_f32, int lds, const int* perm, void* dst, int ldd,
                       int rows, int K, int Kp, void* stream);

/* apply_rotary_pos_emb_vision on a fused qkv buffer (encoder/__init__.py:188-199, 523-550). pos_rc = int2 (row,col). */
int sb_rope_vision(int dtype, void* qkv, int ld, const int* pos_rc, const float* inv_freq, int n_tok, int nh, int d,
                   void* stream);

/* apply_rotary_pos_emb (decoder/__init__.py:60-84, 346-361) + DynamicCache.update without the torch.cat
 * (decoder/__init__.py:190-195): rotates q,k in place and writes k,v into cache[slot][kv_head][pos][:]. */
int sb_rope_kv_append(int dtype, void* qkv, int ld, const int* tok_pos, const int* tok_slot, const float* inv_freq,
                      void* kcache, void* vcache, int n_tok, int nh, int nkv, int d, int s_max, void* stream);

/* flash_attn_varlen_func / SDPA over packed sequences (encoder/__init__.py:174-176, 400-406;
 * flash_attn_utils.py:106-154).  Head h of q at column q_col0 + h*d of a row-major matrix with pitch ldq. */
int sb_attn_varlen(int dtype, const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out,
                   int ldo, const int* seq_start, const int* seq_len, int n_seq, int max_len, int n_heads,
                   int n_kv_heads, int head_dim, int causal, float scale, void* stream);

/* flash_attn_with_kvcache / SDPA decode step (flash_attn_utils.py:157-188) fused with RoPE and cache append. */
int sb_decode_attn(int dtype, const void* qkv, int ld, void* kcache, void* vcache, const int* slot, const int* pos,
                   const float* inv_freq, void* out, int ldo, int batch, int n_heads, int n_kv_heads, int head_dim,
                   int s_max, float scale, void* stream);

/* embed_ids_boxes_images: token embedding + 2-D learned image position embedding + masked_scatter
 * (surya/common/surya/__init__.py:197-272). feat_row[t] < 0 -> plain token embedding. */
int sb_embed_splice(int dtype, const long long* ids, const int* feat_row, const int* hidx, const int* widx,
                    const void* embed, const void* feat, int ldf, const void* h_embed, const void* w_embed,
                    void* out, int ldo, int n_tok, int H, void* stream);
int sb_embed_rows(int dtype, const long long* ids, const void* embed, void* out, int ldo, int n, int H, void* stream);

/* RecognitionPredictor.process_outputs (surya/recognition/__init__.py:294-324): argmax, max-softmax score,
 * done mask, next input id. */
int sb_argmax_score(int dtype, const void* logits, int ld, int rows, int V, long long* tok, float* score,
                    unsigned char* done, long long* next_ids, int eos, int pad, void* stream);

/* bbox_head + sigmoid (+ trunc(sig * bbox_size)) (surya/common/surya/__init__.py:329; recognition/__init__.py:315-316). */
int sb_small_head(int dtype, const void* x, int ldx, const void* w, const void* b, int rows, int H, int n_out,
                  int sigmoid, float* out_f, long long* out_box, float box_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SURYA_B200_H */
