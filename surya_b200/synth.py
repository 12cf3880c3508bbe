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


def rec_state_dict(cfg: RecConfig, seed: int = 0, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """fp32 state dict for SuryaModel (names as in surya/common/surya/__init__.py + encoder/ + decoder/)."""
    sd: Dict[str, torch.Tensor] = {}
    e, d = cfg.vision_encoder, cfg.decoder
    H = e.hidden_size
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
        sd[b + "mlp.gate_proj.weight"] = _normal(b + "mlp.gate_proj.weight", (d.intermediate_size, D), std, seed)
        sd[b + "mlp.up_proj.weight"] = _normal(b + "mlp.up_proj.weight", (d.intermediate_size, D), std, seed)
        sd[b + "mlp.down_proj.weight"] = _normal(b + "mlp.down_proj.weight", (D, d.intermediate_size), std, seed)
    sd["decoder.norm.weight"] = _norm_w("decoder.norm.weight", D, seed)

    sd["embedder.token_embed.weight"] = _normal("embedder.token_embed.weight", (cfg.vocab_size, D), std, seed)
    sd["lm_head.weight"] = sd["embedder.token_embed.weight"]  # tied (surya/common/surya/__init__.py:111-116)
    sd["lm_head.bias"] = _normal("lm_head.bias", (cfg.vocab_size,), std, seed)
    sd["bbox_head.weight"] = _normal("bbox_head.weight", (6, D), std, seed)
    sd["bbox_head.bias"] = _normal("bbox_head.bias", (6,), std, seed)
    sd["img_h_embed.weight"] = _normal("img_h_embed.weight", (cfg.image_embed_encoding_size, D), std, seed)
    sd["img_w_embed.weight"] = _normal("img_w_embed.weight", (cfg.image_embed_encoding_size, D), std, seed)
    return sd


def rec_synthetic_crops(n: int, height: int = 48, width: int = 512, seed: int = 1234):
    """BASELINE config 2 input: uint8 [n, H, W, 3] uniform noise."""
    import numpy as np

    return np.random.default_rng(seed).integers(0, 256, size=(n, height, width, 3), dtype=np.uint8)
