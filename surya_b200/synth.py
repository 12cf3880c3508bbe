"""Deterministic synthetic weights, keyed by the reference's state_dict names.

No checkpoints are available offline, so benchmarks and parity tests use seeded random weights.  Each tensor
is drawn from its own CPU generator seeded by crc32(name) ^ seed, so the reference modules (this container,
oracle/make_golden.py), the CPU oracle and the CUDA engine (GPU box) all see bit-identical fp32 values
without shipping a checkpoint.  Biases and norm weights are randomised as well so that they are exercised.
"""
from __future__ import annotations

import zlib
from typing import Dict

import torch

from .config import RecConfig


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _normal(name, shape, std, seed):
    return torch.randn(shape, generator=_gen(name, seed), dtype=torch.float32) * std


def _norm_w(name, n, seed):
    return 1.0 + 0.1 * torch.randn(n, generator=_gen(name, seed), dtype=torch.float32)


def rec_state_dict(cfg: RecConfig, seed: int = 0, std: float | None = None) -> Dict[str, torch.Tensor]:
    """fp32 state dict for SuryaModel (names as in surya/common/surya/__init__.py + encoder/ + decoder/).

    Recipe (round 2): plain N(0, 0.02) weights make an untrained decoder emit ONE token for ever (the 0.02-RMS token embedding
    drowns in the first attention output, so the final hidden state never turns) — useless as a token-parity input.  Chosen so
    that greedy decoding walks through the vocabulary (24 distinct ids in 24 steps on SYN-REC, 27-31 of 32 on the tiny config)
    while logits stay O(1-3):
      * linear weights N(0, std), std = 0.02 * sqrt(1280 / hidden) per tower (0.02 at SYN-REC width): unit-RMS inputs give
        ~0.7-RMS outputs at every width;
      * decoder gate/up projections x4: the token-dependent MLP path outweighs the context-averaging attention path;
      * token embedding: first half of the channels N(0, 0.3) (the token survives in the residual stream), second half N(0, std);
        final RMSNorm weight is ZERO on the first half, so the tied lm_head reads only channels where the embedding is small —
        otherwise logit[tok] = |E_tok| * cos(h, E_tok) dominates and greedy decoding repeats its own input;
      * lm_head bias -6 on the special ids except EOS (they are never valid outputs)."""
    sd: Dict[str, torch.Tensor] = {}
    e, d = cfg.vision_encoder, cfg.decoder
    H = e.hidden_size
    std_dec = std if std is not None else 0.02 * (1280.0 / d.hidden_size) ** 0.5
    std = std if std is not None else 0.02 * (1280.0 / H) ** 0.5
    p = "vision_encoder."
    sd[p + "patch_embed.proj.weight"] = _normal(p + "patch_embed.proj.weight",
                                                (H, e.in_channels, e.temporal_patch_size, e.patch_size, e.patch_size),
                                                std, seed)
    for i in range(e.depth):
        b = f"{p}blocks.{i}."
        sd[b + "norm1.weight"] = _norm_w(b + "norm1.weight", H, seed)
        sd[b + "norm2.weight"] = _norm_w(b + "norm2.weight", H, seed)
        sd[b + "attn.qkv.weight"] = _normal(b + "attn.qkv.weight", (3 * H, H), std, seed)
        sd[b + "attn.qkv.bias"] = _normal(b + "attn.qkv.bias", (3 * H,), std, seed)
        sd[b + "attn.proj.weight"] = _normal(b + "attn.proj.weight", (H, H), std, seed)
        sd[b + "attn.proj.bias"] = _normal(b + "attn.proj.bias", (H,), std, seed)
        for nm, shp in (("gate_proj", (e.intermediate_size, H)), ("up_proj", (e.intermediate_size, H)),
                        ("down_proj", (H, e.intermediate_size))):
            sd[b + f"mlp.{nm}.weight"] = _normal(b + f"mlp.{nm}.weight", shp, std, seed)
            sd[b + f"mlp.{nm}.bias"] = _normal(b + f"mlp.{nm}.bias", (shp[0],), std, seed)
    m = e.spatial_merge_size ** 2 * H
    sd[p + "merger.ln_q.weight"] = _norm_w(p + "merger.ln_q.weight", H, seed)
    sd[p + "merger.mlp.0.weight"] = _normal(p + "merger.mlp.0.weight", (m, m), std, seed)
    sd[p + "merger.mlp.0.bias"] = _normal(p + "merger.mlp.0.bias", (m,), std, seed)
    sd[p + "merger.mlp.2.weight"] = _normal(p + "merger.mlp.2.weight", (e.out_hidden_size, m), std, seed)
    sd[p + "merger.mlp.2.bias"] = _normal(p + "merger.mlp.2.bias", (e.out_hidden_size,), std, seed)

    D = d.hidden_size
    hd = d.head_dim
    std = std_dec
    for i in range(d.num_hidden_layers):
        b = f"decoder.layers.{i}."
        sd[b + "input_layernorm.weight"] = _norm_w(b + "input_layernorm.weight", D, seed)
        sd[b + "post_attention_layernorm.weight"] = _norm_w(b + "post_attention_layernorm.weight", D, seed)
        for nm, rows in (("q_proj", d.num_attention_heads * hd), ("k_proj", d.num_key_value_heads * hd),
                         ("v_proj", d.num_key_value_heads * hd)):
            sd[b + f"self_attn.{nm}.weight"] = _normal(b + f"self_attn.{nm}.weight", (rows, D), std, seed)
            sd[b + f"self_attn.{nm}.bias"] = _normal(b + f"self_attn.{nm}.bias", (rows,), std, seed)
        sd[b + "self_attn.o_proj.weight"] = _normal(b + "self_attn.o_proj.weight", (D, d.num_attention_heads * hd),
                                                    std, seed)
        sd[b + "mlp.gate_proj.weight"] = _normal(b + "mlp.gate_proj.weight", (d.intermediate_size, D), 4 * std, seed)
        sd[b + "mlp.up_proj.weight"] = _normal(b + "mlp.up_proj.weight", (d.intermediate_size, D), 4 * std, seed)
        sd[b + "mlp.down_proj.weight"] = _normal(b + "mlp.down_proj.weight", (D, d.intermediate_size), std, seed)
    sd["decoder.norm.weight"] = _norm_w("decoder.norm.weight", D, seed)
    sd["decoder.norm.weight"][: D // 2] = 0.0

    emb = _normal("embedder.token_embed.weight", (cfg.vocab_size, D), 1.0, seed)
    emb[:, : D // 2] *= 0.3
    emb[:, D // 2:] *= std
    sd["embedder.token_embed.weight"] = emb
    sd["lm_head.weight"] = sd["embedder.token_embed.weight"]  # tied (surya/common/surya/__init__.py:111-116)
    sd["lm_head.bias"] = _normal("lm_head.bias", (cfg.vocab_size,), std, seed)
    # special ids other than EOS are never valid outputs (feeding IMAGE back trips an assert in the reference,
    # surya/common/surya/__init__.py:227); a trained head suppresses them, the synthetic one does it through the bias
    sd["lm_head.bias"][2:16] = -6.0
    sd["lm_head.bias"][0] = -6.0
    sd["bbox_head.weight"] = _normal("bbox_head.weight", (6, D), std, seed)
    sd["bbox_head.bias"] = _normal("bbox_head.bias", (6,), std, seed)
    sd["img_h_embed.weight"] = _normal("img_h_embed.weight", (cfg.image_embed_encoding_size, D), std, seed)
    sd["img_w_embed.weight"] = _normal("img_w_embed.weight", (cfg.image_embed_encoding_size, D), std, seed)
    return sd


def rec_synthetic_crops(n: int, height: int = 48, width: int = 512, seed: int = 1234):
    """BASELINE config 2 input: uint8 [n, H, W, 3] uniform noise."""
    import numpy as np

    return np.random.default_rng(seed).integers(0, 256, size=(n, height, width, 3), dtype=np.uint8)


# ------------------------------------------------------------------------------------------------ detection
def _bn_params(name: str, n: int, seed: int) -> Dict[str, torch.Tensor]:
    """Randomised BatchNorm affine + running stats so that folding is actually exercised."""
    g = _gen(name, seed)
    return {
        f"{name}.weight": 0.5 + torch.rand(n, generator=g),
        f"{name}.bias": 0.1 * torch.randn(n, generator=g),
        f"{name}.running_mean": 0.1 * torch.randn(n, generator=g),
        f"{name}.running_var": 0.5 + torch.rand(n, generator=g),
    }


def det_state_dict(cfg, seed: int = 0) -> Dict[str, torch.Tensor]:
    """fp32 state dict for EfficientViTForSemanticSegmentation (names: SURVEY.md §9.8 / det_arch.py).

    Conv weights use a fan-in scaled normal (std = 1/sqrt(fan_in)) instead of the reference's 0.02 so that
    activations keep O(1) magnitude through ~40 layers and the fp16 parity test exercises real dynamic range."""
    from .det_arch import det_blocks, det_head_specs

    sd: Dict[str, torch.Tensor] = {}

    def conv(spec):
        fan_in = (spec.cin // spec.groups) * spec.k * spec.k
        sd[spec.wkey] = _normal(spec.wkey, (spec.cout, spec.cin // spec.groups, spec.k, spec.k), fan_in ** -0.5, seed)
        if spec.bias:
            sd[spec.bkey] = _normal(spec.bkey, (spec.cout,), 0.1, seed)
        if spec.norm:
            sd.update(_bn_params(f"{spec.name}.norm", spec.cout, seed))

    for blk in det_blocks(cfg):
        for c in blk.convs:
            conv(c)
        for c in blk.mla or []:
            conv(c)
    hs = det_head_specs(cfg)
    for name, cin, cout in hs["linear_c"]:
        sd[f"{name}.weight"] = _normal(f"{name}.weight", (cout, cin), cin ** -0.5, seed)
        sd[f"{name}.bias"] = _normal(f"{name}.bias", (cout,), 0.1, seed)
    name, cin, cout = hs["fuse"]
    sd[f"{name}.weight"] = _normal(f"{name}.weight", (cout, cin, 1, 1), cin ** -0.5, seed)
    sd.update(_bn_params(hs["bn"], cout, seed))
    name, cin, cout = hs["cls"]
    sd[f"{name}.weight"] = _normal(f"{name}.weight", (cout, cin, 1, 1), cin ** -0.5, seed)
    sd[f"{name}.bias"] = _normal(f"{name}.bias", (cout,), 0.1, seed)
    return sd


def det_synthetic_pages(n: int, size: int = 1024, seed: int = 1234, text_like: bool = False):
    """BASELINE config 3 input: uint8 [n, size, size, 3]; uniform noise, or white pages with random dark boxes."""
    import numpy as np

    rng = np.random.default_rng(seed)
    if not text_like:
        return rng.integers(0, 256, size=(n, size, size, 3), dtype=np.uint8)
    pages = np.full((n, size, size, 3), 255, dtype=np.uint8)
    for p in pages:
        for _ in range(40):
            h = int(rng.integers(8, 25))
            w = int(rng.integers(40, size // 2))
            y = int(rng.integers(0, size - h))
            x = int(rng.integers(0, size - w))
            p[y:y + h, x:x + w] = rng.integers(0, 80)
    return pages


def det_normalize(pages) -> torch.Tensor:
    """SegformerImageProcessor: rescale 1/255 + ImageNet mean/std, NCHW fp32 (surya/detection/processor.py:94-95, 140-146)."""
    import numpy as np

    mean = np.array((0.485, 0.456, 0.406), dtype=np.float32)
    std = np.array((0.229, 0.224, 0.225), dtype=np.float32)
    x = (pages.astype(np.float32) * (1 / 255.0) - mean) / std
    return torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))


# ------------------------------------------------------------------------------------------------ layout (Swin + ADETR)
def _ln(name, n, seed):
    g = _gen(name, seed)
    return {f"{name}.weight": 1.0 + 0.1 * torch.randn(n, generator=g), f"{name}.bias": 0.05 * torch.randn(n, generator=g)}


def swin_state_dict(cfg, seed: int = 0, prefix: str = "") -> Dict[str, torch.Tensor]:
    """fp32 state dict for DonutSwinLayoutModel / DonutSwinModel (names: SURVEY.md §9.8)."""
    sd: Dict[str, torch.Tensor] = {}
    p = prefix
    C0 = cfg.embed_dim
    sd[p + "embeddings.patch_embeddings.projection.weight"] = _normal(p + "pe.w", (C0, cfg.num_channels, cfg.patch_size, cfg.patch_size), (cfg.num_channels * cfg.patch_size ** 2) ** -0.5, seed)
    sd[p + "embeddings.patch_embeddings.projection.bias"] = _normal(p + "pe.b", (C0,), 0.05, seed)
    sd.update(_ln(p + "embeddings.norm", C0, seed))
    ws = cfg.window_size
    for s, (depth, nh) in enumerate(zip(cfg.depths, cfg.num_heads)):
        C = C0 * 2 ** s
        for b in range(depth):
            q = f"{p}encoder.layers.{s}.blocks.{b}."
            sd.update(_ln(q + "layernorm_before", C, seed))
            sd.update(_ln(q + "layernorm_after", C, seed))
            sd[q + "attention.self.relative_position_bias_table"] = _normal(q + "rpb", ((2 * ws - 1) ** 2, nh), 0.5, seed)
            for nm in ("query", "key", "value"):
                sd[q + f"attention.self.{nm}.weight"] = _normal(q + nm + ".w", (C, C), C ** -0.5, seed)
                sd[q + f"attention.self.{nm}.bias"] = _normal(q + nm + ".b", (C,), 0.05, seed)
            sd[q + "attention.output.dense.weight"] = _normal(q + "ao.w", (C, C), 0.5 * C ** -0.5, seed)
            sd[q + "attention.output.dense.bias"] = _normal(q + "ao.b", (C,), 0.02, seed)
            I = int(cfg.mlp_ratio * C)
            sd[q + "intermediate.dense.weight"] = _normal(q + "fc1.w", (I, C), C ** -0.5, seed)
            sd[q + "intermediate.dense.bias"] = _normal(q + "fc1.b", (I,), 0.05, seed)
            sd[q + "output.dense.weight"] = _normal(q + "fc2.w", (C, I), 0.5 * I ** -0.5, seed)
            sd[q + "output.dense.bias"] = _normal(q + "fc2.b", (C,), 0.02, seed)
        if s < len(cfg.depths) - 1:
            q = f"{p}encoder.layers.{s}.downsample."
            sd[q + "reduction.weight"] = _normal(q + "red.w", (2 * C, 4 * C), (4 * C) ** -0.5, seed)
            sd.update(_ln(q + "norm", 4 * C, seed))
    sd[p + "position_embeddings"] = _normal(p + "pos", (1, cfg.encoder_length, cfg.hidden_size), 0.1, seed)
    return sd


LAYOUT_EMBED_TABLES = ("w", "h", "cx", "cy", "xskew", "yskew", "x1", "y1", "x2", "y2", "x3", "y3", "x4", "y4")


def adetr_layout_state_dict(cfg, seed: int = 0) -> Dict[str, torch.Tensor]:
    """fp32 state dict for SuryaLayoutDecoder (surya/layout/model/decoder.py + surya/common/adetr/decoder.py)."""
    sd: Dict[str, torch.Tensor] = {}
    H, I = cfg.hidden_size, cfg.intermediate_size
    nh, nkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    for t in LAYOUT_EMBED_TABLES:
        sd[f"model.embed_tokens.{t}_embed.weight"] = _normal(f"emb.{t}", (cfg.vocab_size, H), 0.25, seed)
    sd["model.embed_tokens.label_embed.weight"] = _normal("emb.label", (cfg.label_count, H), 0.25, seed)
    for l in range(cfg.num_hidden_layers):
        q = f"model.layers.{l}."
        for nm in ("cross_pre_norm", "temporal_pre_norm", "channel_pre_norm"):
            sd[q + nm + ".weight"] = 0.1 * torch.randn(H, generator=_gen(q + nm, seed))
        for blk, kin in (("temporal_block", H), ("cross_attn_block", cfg.encoder_hidden_size)):
            sd[q + f"{blk}.q_proj.weight"] = _normal(q + blk + ".q", (nh * hd, H), H ** -0.5, seed)
            sd[q + f"{blk}.k_proj.weight"] = _normal(q + blk + ".k", (nkv * hd, kin), kin ** -0.5, seed)
            sd[q + f"{blk}.v_proj.weight"] = _normal(q + blk + ".v", (nkv * hd, kin), kin ** -0.5, seed)
            sd[q + f"{blk}.o_proj.weight"] = _normal(q + blk + ".o", (H, nh * hd), 0.5 * H ** -0.5, seed)
            sd[q + f"{blk}.o_proj.bias"] = _normal(q + blk + ".ob", (H,), 0.02, seed)
        sd[q + "mlp_block.gate_proj.weight"] = _normal(q + "gate", (I, H), H ** -0.5, seed)
        sd[q + "mlp_block.up_proj.weight"] = _normal(q + "up", (I, H), H ** -0.5, seed)
        sd[q + "mlp_block.down_proj.weight"] = _normal(q + "down", (H, I), 0.5 * I ** -0.5, seed)
    sd["model.final_norm.weight"] = 0.1 * torch.randn(H, generator=_gen("final_norm", seed))
    sd["lm_head.weight"] = _normal("lm_head", (cfg.label_count, H), H ** -0.5, seed)
    sd["bbox_head.weight"] = _normal("bbox_head.w", (6, H), H ** -0.5, seed)
    sd["bbox_head.bias"] = _normal("bbox_head.b", (6,), 0.1, seed)
    sd.update(_ln("pre_output_norm", H, seed))
    return sd


TABLE_BOX_TABLES = ("w", "h", "cx", "cy", "xskew", "yskew", "x1", "y1", "x2", "y2", "x3", "y3", "x4", "y4")
TABLE_HEADS = ("bbox", "category", "merges", "colspan", "is_header")


def table_head_sizes(cfg) -> Dict[str, int]:
    return {"bbox": 6, "category": cfg.category_count, "merges": cfg.merge_count, "colspan": 1, "is_header": cfg.header_count}


def adetr_table_state_dict(cfg, seed: int = 0) -> Dict[str, torch.Tensor]:
    """fp32 state dict for SuryaTableRecDecoder (surya/table_rec/model/decoder.py:12-119): the layout decoder's layer
    stack with LabelEmbedding tables and five bias-free property heads.  The category / merge embedding tables carry the
    special-token offset twice (decoder.py:36-37 adds it on top of get_box_property's), the heads once."""
    base = adetr_layout_state_dict(cfg, seed)
    sd = {k: v for k, v in base.items() if k.startswith("model.layers.") or k.startswith("model.final_norm") or k.startswith("pre_output_norm")}
    for t in TABLE_BOX_TABLES:
        sd[f"model.embed_tokens.{t}_embed.weight"] = _normal(f"temb.{t}", (cfg.vocab_size, cfg.box_embed_size), 0.25, seed)
    sd["model.embed_tokens.category_embed.weight"] = _normal("temb.cat", (cfg.category_count + cfg.special_token_count, cfg.property_embed_size), 0.25, seed)
    sd["model.embed_tokens.merge_embed.weight"] = _normal("temb.merge", (cfg.merge_count + cfg.special_token_count, cfg.property_embed_size), 0.25, seed)
    sd["model.embed_tokens.colspan_embed.weight"] = _normal("temb.colspan", (cfg.vocab_size, cfg.property_embed_size), 0.25, seed)
    H = cfg.hidden_size
    for k, n in table_head_sizes(cfg).items():
        sd[f"box_property_heads.{k}.weight"] = _normal(f"thead.{k}", (n, H), H ** -0.5, seed)
    return sd


def table_query_tokens(cfg, n: int, seed: int = 0) -> torch.Tensor:
    """The row/column-pass prompt SuryaTableRecProcessor builds (surya/table_rec/processor.py:65-76): [bos x10], the query
    box (whole-image table polygon -> cx, cy, w, h, skews, category 'Table' + specials, 0 + specials x3), [query_end x10]."""
    lab = [512, 512, 1024, 1024, 512, 512, 4 + cfg.special_token_count, cfg.special_token_count, 0, cfg.special_token_count]
    tok = torch.tensor([[cfg.bos_token_id] * 10, lab, [cfg.query_end_token_id] * 10], dtype=torch.long)
    return tok.unsqueeze(0).repeat(n, 1, 1)


def layout_synthetic_pages(n: int, size=(768, 768), seed: int = 1234) -> torch.Tensor:
    """BASELINE config 4 input: uint8 noise pages through the Donut processor's rescale + normalise
    (surya/common/donut/processor.py:61-116: resize is a no-op at the native size); NCHW fp32."""
    import numpy as np

    rng = np.random.default_rng(seed)
    x = rng.integers(0, 256, size=(n, size[0], size[1], 3), dtype=np.uint8).astype(np.float32) * (1 / 255.0)
    mean = np.array((0.485, 0.456, 0.406), dtype=np.float32)
    std = np.array((0.229, 0.224, 0.225), dtype=np.float32)
    return torch.from_numpy(np.ascontiguousarray(((x - mean) / std).transpose(0, 3, 1, 2)))


# ------------------------------------------------------------------------------------------------ ocr_error (DistilBERT)
def ocr_error_state_dict(cfg, seed: int = 0) -> Dict[str, torch.Tensor]:
    """fp32 state dict for DistilBertForSequenceClassification (names as in surya/ocr_error/model/encoder.py:48-58, 94-117,
    381-389, 408-420, 697-706).  Linear weights N(0, 0.7 / sqrt(fan_in)) keep every sub-layer's output near unit RMS so the two
    class logits are O(1) and differ between texts; embeddings N(0, 1) / N(0, 0.3)."""
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, n_out, n_in, scale=0.7):
        sd[name + ".weight"] = _normal(name + ".weight", (n_out, n_in), scale / n_in ** 0.5, seed)
        sd[name + ".bias"] = _normal(name + ".bias", (n_out,), 0.05, seed)

    e = "distilbert.embeddings."
    sd[e + "word_embeddings.weight"] = _normal(e + "word_embeddings.weight", (cfg.vocab_size, cfg.dim), 1.0, seed)
    sd[e + "word_embeddings.weight"][cfg.pad_token_id] = 0.0          # nn.Embedding(padding_idx=pad) row
    sd[e + "position_embeddings.weight"] = _normal(e + "position_embeddings.weight", (cfg.max_position_embeddings, cfg.dim), 0.3, seed)
    sd.update(_ln(e + "LayerNorm", cfg.dim, seed))
    for i in range(cfg.n_layers):
        b = f"distilbert.transformer.layer.{i}."
        for n in ("q_lin", "k_lin", "v_lin", "out_lin"):
            lin(b + "attention." + n, cfg.dim, cfg.dim, 2.0 if n in ("q_lin", "k_lin") else 1.5)
        sd.update(_ln(b + "sa_layer_norm", cfg.dim, seed))
        lin(b + "ffn.lin1", cfg.hidden_dim, cfg.dim)
        lin(b + "ffn.lin2", cfg.dim, cfg.hidden_dim)
        sd.update(_ln(b + "output_layer_norm", cfg.dim, seed))
    lin("pre_classifier", cfg.dim, cfg.dim, 1.0)
    lin("classifier", cfg.num_labels, cfg.dim, 4.0)
    # zero-sum classifier rows: the post-ReLU pooled vector has a large common mean that would otherwise pick one label for every text
    sd["classifier.weight"] -= sd["classifier.weight"].mean(dim=1, keepdim=True)
    return sd


def ocr_error_synthetic_batch(cfg, n: int, max_len: int = 64, seed: int = 0, min_len: int = 3):
    """Right-padded (input_ids int64 [n, L], attention_mask int64 [n, L]) like the tokenizer call of OCRErrorPredictor
    (surya/ocr_error/__init__.py:28-30: padding='longest'): seeded lengths in [min_len, max_len], at least one row of full length,
    [CLS]-like id 101 % vocab first, pad id after the end."""
    g = _gen("ocr_error_batch", seed)
    lens = torch.randint(min_len, max_len + 1, (n,), generator=g)
    lens[int(torch.randint(0, n, (1,), generator=g))] = max_len
    ids = torch.randint(1, cfg.vocab_size, (n, max_len), generator=g, dtype=torch.int64)
    ids[:, 0] = 101 % cfg.vocab_size
    mask = (torch.arange(max_len)[None, :] < lens[:, None]).to(torch.int64)
    ids = torch.where(mask.bool(), ids, torch.full_like(ids, cfg.pad_token_id))
    return ids, mask
