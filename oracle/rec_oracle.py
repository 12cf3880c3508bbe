"""CPU ORACLE (test infrastructure, not product code) — recognition path of VikParuchuri/surya v0.14.6.

A functional, plain-PyTorch restatement of the reference's recognition forward pass, written against a flat
state dict with the reference's parameter names.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product (surya_b200/) never does.

Pinning: the reference ships no tensor-level golden vectors, so this restatement is pinned
against outputs of the reference's own nn.Modules, generated in the build container by oracle/make_golden.py
(committed under tests/golden/) and re-checked by tests/test_oracle_golden.py.

Every function cites the reference code it restates (paths relative to the reference checkout).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------------------------ helpers
def cast_sd(sd: SD, dtype: torch.dtype) -> SD:
    """model.to(dtype): every floating parameter is rounded to the compute dtype once."""
    out = {}
    seen = {}
    for k, v in sd.items():
        key = v.data_ptr()
        if key not in seen:
            seen[key] = v.to(dtype)
        out[k] = seen[key]
    return out


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """Qwen2RMSNorm — surya/common/surya/decoder/__init__.py:250-255, encoder/__init__.py:99-104."""
    dt = x.dtype
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    return w * xf.to(dt)


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    """surya/common/surya/decoder/__init__.py:53-57 (same in encoder/__init__.py:181-185)."""
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


# ------------------------------------------------------------------------------------------------ vision tower
def vision_rot_pos_ids(grid_thw: np.ndarray, merge: int) -> np.ndarray:
    """rot_pos_emb position ids — surya/common/surya/encoder/__init__.py:523-546. Returns [N, 2] (row, col)."""
    out = []
    for t, h, w in grid_thw:
        hp = np.arange(h)[:, None].repeat(w, 1).reshape(h // merge, merge, w // merge, merge)
        hp = hp.transpose(0, 2, 1, 3).reshape(-1)
        wp = np.arange(w)[None, :].repeat(h, 0).reshape(h // merge, merge, w // merge, merge)
        wp = wp.transpose(0, 2, 1, 3).reshape(-1)
        out.append(np.tile(np.stack([hp, wp], -1), (t, 1)))
    return np.concatenate(out, 0)


def vision_window_index(grid_thw: np.ndarray, window_size: int, merge: int, patch: int):
    """get_window_index — surya/common/surya/encoder/__init__.py:552-597 (+ unique_consecutive :620)."""
    window_index: List[np.ndarray] = []
    cu = [0]
    base = 0
    vws = window_size // merge // patch
    unit = merge * merge
    for t, h, w in grid_thw:
        lh, lw = h // merge, w // merge
        index = np.arange(t * lh * lw).reshape(t, lh, lw)
        pad_h = vws - lh % vws
        pad_w = vws - lw % vws
        nwh, nww = (lh + pad_h) // vws, (lw + pad_w) // vws
        padded = np.pad(index, ((0, 0), (0, pad_h), (0, pad_w)), constant_values=-100)
        padded = padded.reshape(t, nwh, vws, nww, vws).transpose(0, 1, 3, 2, 4).reshape(t, nwh * nww, vws, vws)
        seqlens = (padded != -100).sum((2, 3)).reshape(-1)
        flat = padded.reshape(-1)
        window_index.append(flat[flat != -100] + base)
        cu.extend((np.cumsum(seqlens) * unit + cu[-1]).tolist())
        base += int(t * lh * lw)
    cu_arr = np.array(cu, dtype=np.int64)
    keep = np.ones(len(cu_arr), dtype=bool)
    keep[1:] = cu_arr[1:] != cu_arr[:-1]
    return np.concatenate(window_index), cu_arr[keep]


def _block_attention(q, k, v, cu_seqlens: np.ndarray, scale_div: float):
    """Qwen2_5_VLVisionAttention eager path — encoder/__init__.py:238-261 (block-diagonal mask, fp32 softmax)."""
    n = q.shape[0]
    mask = torch.full((1, n, n), torch.finfo(q.dtype).min, dtype=q.dtype, device=q.device)
    for i in range(1, len(cu_seqlens)):
        a, b = int(cu_seqlens[i - 1]), int(cu_seqlens[i])
        mask[..., a:b, a:b] = 0
    q, k, v = q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1)
    w = torch.matmul(q, k.transpose(1, 2)) / scale_div
    w = w + mask
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    o = torch.matmul(w, v)
    return o.transpose(0, 1).reshape(n, -1)


# bench.py's `gpu_eager_baseline` flips this: equal-length segments go through ONE batched F.scaled_dot_product_attention call
# and the decoder uses SDPA with enable_gqa — a PyTorch-eager path at least as fast as what the reference runs on a GPU (its
# SDPA branch pads windows in Python loops, encoder/__init__.py:298-330).  Parity tests never set it.
FAST_ATTENTION = False


def _block_attention_chunked(q, k, v, cu_seqlens: np.ndarray, scale_div: float):
    """Same math as _block_attention evaluated per sequence (identical results, O(sum L^2) instead of O(N^2))."""
    lens = np.diff(np.asarray(cu_seqlens))
    if FAST_ATTENTION and len(lens) and (lens == lens[0]).all():
        L, nseg = int(lens[0]), len(lens)
        nh, hd = q.shape[1], q.shape[2]
        qq, kk, vv = (t.reshape(nseg, L, nh, hd).transpose(1, 2) for t in (q, k, v))
        o = F.scaled_dot_product_attention(qq, kk, vv, scale=1.0 / scale_div)
        return o.transpose(1, 2).reshape(nseg * L, nh * hd)
    outs = []
    for i in range(1, len(cu_seqlens)):
        a, b = int(cu_seqlens[i - 1]), int(cu_seqlens[i])
        qq, kk, vv = q[a:b].transpose(0, 1), k[a:b].transpose(0, 1), v[a:b].transpose(0, 1)
        w = torch.matmul(qq, kk.transpose(1, 2)) / scale_div
        w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
        outs.append(torch.matmul(w, vv).transpose(0, 1).reshape(b - a, -1))
    return torch.cat(outs, 0)


def vision_tower(sd: SD, cfg, tiles: torch.Tensor, grid_thw: np.ndarray, chunked_attn: bool = True) -> torch.Tensor:
    """Qwen2_5_VisionTransformerPretrainedModel.forward — encoder/__init__.py:599-672.
    tiles [N, C*T*P*P] in the compute dtype; returns merged tokens [N/4, out_hidden] in original order."""
    e = cfg.vision_encoder
    p = "vision_encoder."
    dt = sd[p + "patch_embed.proj.weight"].dtype
    H, nh = e.hidden_size, e.num_heads
    hd = H // nh
    unit = e.spatial_merge_size ** 2
    # patch embed: Conv3d with kernel == stride == one GEMM (:53-73)
    wpe = sd[p + "patch_embed.proj.weight"].reshape(H, -1)
    x = F.linear(tiles.to(dt), wpe)
    # rotary table (:76-87, :523-550) and window permutation (:614-634)
    pos = vision_rot_pos_ids(grid_thw, e.spatial_merge_size)
    dev = tiles.device
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, hd // 2, 2, dtype=torch.float, device=dev) / (hd // 2)))
    max_grid = int(grid_thw[:, 1:].max())
    freqs_full = torch.outer(torch.arange(max_grid, dtype=torch.float, device=dev), inv_freq)
    rot = freqs_full[torch.from_numpy(pos).to(dev)].flatten(1)  # [N, hd/2]
    widx, cu_win = vision_window_index(grid_thw, e.window_size, e.spatial_merge_size, e.patch_size)
    widx_t = torch.from_numpy(widx).to(dev)
    n = x.shape[0]
    x = x.reshape(n // unit, unit, -1)[widx_t].reshape(n, -1)
    rot = rot.reshape(n // unit, unit, -1)[widx_t].reshape(n, -1)
    emb = torch.cat((rot, rot), dim=-1)
    cos, sin = emb.cos(), emb.sin()
    cu_full = np.concatenate([[0], np.cumsum(np.repeat(grid_thw[:, 1] * grid_thw[:, 2], grid_thw[:, 0]))])
    attn = _block_attention_chunked if chunked_attn else _block_attention
    for li in range(e.depth):
        b = f"{p}blocks.{li}."
        cu = cu_full if li in e.fullatt_block_indexes else cu_win
        hn = rms_norm(x, sd[b + "norm1.weight"], 1e-6)
        qkv = F.linear(hn, sd[b + "attn.qkv.weight"], sd[b + "attn.qkv.bias"])
        q, k, v = qkv.reshape(n, 3, nh, -1).permute(1, 0, 2, 3).unbind(0)
        # apply_rotary_pos_emb_vision (:188-199): fp32, one rounding
        c32, s32 = cos.unsqueeze(-2).float(), sin.unsqueeze(-2).float()
        qf, kf = q.float(), k.float()
        q = ((qf * c32) + (rotate_half(qf) * s32)).to(dt)
        k = ((kf * c32) + (rotate_half(kf) * s32)).to(dt)
        a = attn(q, k, v, cu, math.sqrt(hd))
        x = x + F.linear(a, sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"])
        hn = rms_norm(x, sd[b + "norm2.weight"], 1e-6)
        g = F.linear(hn, sd[b + "mlp.gate_proj.weight"], sd[b + "mlp.gate_proj.bias"])
        u = F.linear(hn, sd[b + "mlp.up_proj.weight"], sd[b + "mlp.up_proj.bias"])
        x = x + F.linear(F.silu(g) * u, sd[b + "mlp.down_proj.weight"], sd[b + "mlp.down_proj.bias"])
    # merger (:110-123) then inverse window permutation (:668-670)
    m = rms_norm(x, sd[p + "merger.ln_q.weight"], 1e-6).view(-1, unit * H)
    m = F.linear(m, sd[p + "merger.mlp.0.weight"], sd[p + "merger.mlp.0.bias"])
    m = F.gelu(m)
    m = F.linear(m, sd[p + "merger.mlp.2.weight"], sd[p + "merger.mlp.2.bias"])
    return m[torch.argsort(widx_t)]


def learned_2d_embeddings(sd: SD, cfg, grid_thw: np.ndarray) -> torch.Tensor:
    """get_2d_learned_embeddings — surya/common/surya/__init__.py:233-272."""
    outs = []
    mult = cfg.image_embed_encoding_multiplier
    for _, gh, gw in grid_thw:
        lh, lw = int(gh) // cfg.merge_size, int(gw) // cfg.merge_size
        dev = sd["img_h_embed.weight"].device
        hv = torch.arange(lh) / max(1, lh - 1) * mult
        wv = torch.arange(lw) / max(1, lw - 1) * mult
        he = sd["img_h_embed.weight"][hv.to(torch.long).to(dev)]
        we = sd["img_w_embed.weight"][wv.to(torch.long).to(dev)]
        outs.append((he[:, None] + we[None, :]).flatten(0, 1))
    return torch.cat(outs, 0)


def embed_inputs(sd: SD, cfg, input_ids: torch.Tensor, tiles: Optional[torch.Tensor], grid_thw: Optional[np.ndarray]):
    """embed_ids_boxes_images — surya/common/surya/__init__.py:197-231 (encoder chunking :137-170 is a no-op
    on the values: chunks split at image boundaries only)."""
    emb = sd["embedder.token_embed.weight"][input_ids]
    if tiles is not None:
        feats = vision_tower(sd, cfg, tiles, grid_thw) + learned_2d_embeddings(sd, cfg, grid_thw)
        mask = (input_ids == cfg.image_token_id).unsqueeze(-1).expand_as(emb)
        emb = emb.masked_scatter(mask, feats.to(emb.dtype))
    return emb


# ------------------------------------------------------------------------------------------------ decoder
@dataclass
class OracleCache:
    """DynamicCache semantics (transformers): per layer, concatenate on the sequence axis."""
    k: List[Optional[torch.Tensor]] = field(default_factory=list)
    v: List[Optional[torch.Tensor]] = field(default_factory=list)

    def seq_len(self) -> int:
        return 0 if not self.k or self.k[0] is None else self.k[0].shape[-2]

    def update(self, layer: int, k: torch.Tensor, v: torch.Tensor):
        while len(self.k) <= layer:
            self.k.append(None)
            self.v.append(None)
        if self.k[layer] is None:
            self.k[layer], self.v[layer] = k, v
        else:
            self.k[layer] = torch.cat([self.k[layer], k], dim=-2)
            self.v[layer] = torch.cat([self.v[layer], v], dim=-2)
        return self.k[layer], self.v[layer]


def causal_padding_mask(attention_mask: torch.Tensor, q_len: int, past_len: int, dtype: torch.dtype) -> torch.Tensor:
    """_prepare_4d_causal_attention_mask_with_cache_position — decoder/__init__.py:554-631.
    key j visible to query at cache position p iff j <= p and attention_mask[b, j] == 1; masked = finfo.min."""
    B, target = attention_mask.shape
    min_v = torch.finfo(dtype).min
    dev = attention_mask.device
    cache_position = torch.arange(past_len, past_len + q_len, device=dev)
    m = torch.full((q_len, target), min_v, dtype=dtype, device=dev)
    m = m * (torch.arange(target, device=dev) > cache_position.reshape(-1, 1))
    m = m[None, None].expand(B, 1, -1, -1).clone()
    pad = (m + attention_mask[:, None, None, :].to(dtype)) == 0
    return m.masked_fill(pad, min_v)


def decoder_forward(sd: SD, cfg, x: torch.Tensor, attention_mask: torch.Tensor, position_ids: torch.Tensor,
                    cache: OracleCache) -> torch.Tensor:
    """SuryaDecoderModel.forward — decoder/__init__.py:417-490 with Qwen2DecoderLayer :272-316,
    Qwen2Attention :161-238 (eager attention :101-128), Qwen2MLP :48-50, rotary :346-361, :60-84."""
    d = cfg.decoder
    dt = x.dtype
    B, q_len, _ = x.shape
    nh, nkv, hd = d.num_attention_heads, d.num_key_value_heads, d.head_dim
    past = cache.seq_len()
    mask = causal_padding_mask(attention_mask, q_len, past, dt)
    inv_freq = 1.0 / (d.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64, device=x.device).float() / hd))
    freqs = (inv_freq[None, :, None].float().expand(B, -1, 1) @ position_ids[:, None, :].float()).transpose(1, 2)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos().to(dt).unsqueeze(1), emb.sin().to(dt).unsqueeze(1)
    for li in range(d.num_hidden_layers):
        b = f"decoder.layers.{li}."
        hn = rms_norm(x, sd[b + "input_layernorm.weight"], d.rms_norm_eps)
        q = F.linear(hn, sd[b + "self_attn.q_proj.weight"], sd[b + "self_attn.q_proj.bias"]).view(B, q_len, nh, hd).transpose(1, 2)
        k = F.linear(hn, sd[b + "self_attn.k_proj.weight"], sd[b + "self_attn.k_proj.bias"]).view(B, q_len, nkv, hd).transpose(1, 2)
        v = F.linear(hn, sd[b + "self_attn.v_proj.weight"], sd[b + "self_attn.v_proj.bias"]).view(B, q_len, nkv, hd).transpose(1, 2)
        q = (q * cos) + (rotate_half(q) * sin)
        k = (k * cos) + (rotate_half(k) * sin)
        k_all, v_all = cache.update(li, k, v)
        rep = nh // nkv
        if FAST_ATTENTION:
            a = F.scaled_dot_product_attention(q, k_all, v_all, attn_mask=mask[:, :, :, : k_all.shape[-2]], enable_gqa=True)
            a = a.transpose(1, 2).reshape(B, q_len, -1)
        else:
            kk = k_all[:, :, None].expand(B, nkv, rep, k_all.shape[-2], hd).reshape(B, nh, -1, hd)
            vv = v_all[:, :, None].expand(B, nkv, rep, v_all.shape[-2], hd).reshape(B, nh, -1, hd)
            w = torch.matmul(q, kk.transpose(2, 3)) * (hd ** -0.5)
            w = w + mask[:, :, :, : kk.shape[-2]]
            w = F.softmax(w, dim=-1, dtype=torch.float32).to(dt)
            a = torch.matmul(w, vv).transpose(1, 2).reshape(B, q_len, -1)
        x = x + F.linear(a, sd[b + "self_attn.o_proj.weight"])
        hn = rms_norm(x, sd[b + "post_attention_layernorm.weight"], d.rms_norm_eps)
        g = F.linear(hn, sd[b + "mlp.gate_proj.weight"])
        u = F.linear(hn, sd[b + "mlp.up_proj.weight"])
        x = x + F.linear(F.silu(g) * u, sd[b + "mlp.down_proj.weight"])
    return rms_norm(x, sd["decoder.norm.weight"], d.rms_norm_eps)


def model_forward(sd: SD, cfg, input_ids, attention_mask, position_ids, cache: OracleCache, tiles=None, grid_thw=None):
    """SuryaModel.forward with logits_to_keep=1 — surya/common/surya/__init__.py:274-338."""
    x = embed_inputs(sd, cfg, input_ids, tiles, grid_thw)
    h = decoder_forward(sd, cfg, x, attention_mask, position_ids, cache)[:, -1:, :].contiguous()
    bbox = torch.sigmoid(F.linear(h, sd["bbox_head.weight"], sd["bbox_head.bias"]))
    lm = F.linear(h, sd["lm_head.weight"], sd["lm_head.bias"])
    return lm, bbox


def process_outputs(lm_logits, bbox_logits, cfg):
    """RecognitionPredictor.process_outputs — surya/recognition/__init__.py:294-324."""
    nt = lm_logits[:, -1:, :].clone().float()
    nb = bbox_logits[:, -1:, :].clone().float()
    preds = torch.argmax(nt, dim=-1)
    done = ((preds == cfg.eos_token_id) | (preds == cfg.pad_token_id)).squeeze(-1)
    input_ids = torch.where(done.unsqueeze(1), torch.tensor(cfg.pad_token_id, device=preds.device), preds).to(torch.long)
    scores = torch.max(F.softmax(nt[:, -1], dim=-1), dim=-1).values
    scores = scores.masked_fill(done, 0).unsqueeze(1)
    boxes = (nb * cfg.bbox_size).to(torch.long)
    return input_ids, preds, boxes, done, scores


def detect_repeat_token(tokens: List[int], max_repeats: int = 40) -> bool:
    """surya/recognition/util.py:59-69."""
    if len(tokens) < max_repeats:
        return False
    last_n = tokens[-max_repeats:]
    u = len(set(last_n))
    if u > 5:
        return False
    return last_n[-u:] == last_n[-u * 2: -u]


# ------------------------------------------------------------------------------------------------ host-side processor
IMAGE_MEAN = np.array((0.485, 0.456, 0.406), dtype=np.float32)
IMAGE_STD = np.array((0.229, 0.224, 0.225), dtype=np.float32)


def scale_to_fit(img: np.ndarray, max_size=(1024, 256), min_size=(168, 168)) -> np.ndarray:
    """SuryaOCRProcessor.scale_to_fit — surya/common/surya/processor/__init__.py:140-178."""
    import cv2

    h, w = img.shape[:2]
    if w == 0 or h == 0:
        return img
    cur, mx, mn = w * h, max_size[0] * max_size[1], min_size[0] * min_size[1]
    if cur > mx:
        s = (mx / cur) ** 0.5
        nw, nh = math.floor(w * s), math.floor(h * s)
    elif cur < mn:
        s = (mn / cur) ** 0.5
        nw, nh = math.ceil(w * s), math.ceil(h * s)
    else:
        return img
    return cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LANCZOS4)


def process_and_tile(image: np.ndarray, patch: int = 14, merge: int = 2):
    """_process_and_tile — processor/__init__.py:185-230 (+ _image_processor :180-183)."""
    import cv2

    factor = patch * merge
    h, w = image.shape[:2]
    hb, wb = math.ceil(h / factor) * factor, math.ceil(w / factor) * factor
    if hb != h or wb != w:
        image = cv2.resize(image, (wb, hb), interpolation=cv2.INTER_CUBIC)
    image = image.astype(np.float64) * (1 / 255.0)
    image = (image.astype(np.float32) - IMAGE_MEAN) / IMAGE_STD
    h, w = image.shape[:2]
    t = torch.from_numpy(np.ascontiguousarray(image.transpose(2, 0, 1))).unsqueeze(0)
    gh, gw = h // patch, w // patch
    c = t.shape[1]
    t = t.reshape(1, 1, c, gh // merge, merge, patch, gw // merge, merge, patch).permute(0, 3, 6, 4, 7, 2, 1, 5, 8)
    return t.reshape(gh * gw, c * patch * patch), (1, gh, gw)


def build_batch(crops: List[np.ndarray], cfg, math_mode: bool = True):
    """prepare_input + SuryaOCRProcessor.__call__ for task ocr_with_boxes with empty input text —
    surya/recognition/__init__.py:259-292, processor/__init__.py:232-274, 288-329, 361-424 (left padding)."""
    all_ids, all_tiles, all_grid = [], [], []
    for crop in crops:
        img = scale_to_fit(np.asarray(crop, dtype=np.float32), (1024, 256))
        tiles, grid = process_and_tile(img, cfg.vision_encoder.patch_size, cfg.merge_size)
        n_tok = tiles.shape[0] // (cfg.merge_size ** 2)
        ids = [cfg.image_token_id] * n_tok + list(cfg.register_token_ids[: cfg.num_register_tokens])
        text_ids = [] if math_mode else [cfg.nomath_token_id]
        ids = ids + [cfg.ocr_with_boxes_bos_id] + text_ids + [cfg.eoi_token_id]
        all_ids.append(torch.tensor(ids, dtype=torch.long))
        all_tiles.append(tiles)
        all_grid.append(grid)
    S = max(len(i) for i in all_ids)
    input_ids = torch.full((len(all_ids), S), cfg.pad_token_id, dtype=torch.long)
    for i, ids in enumerate(all_ids):
        input_ids[i, S - len(ids):] = ids
    attention_mask = input_ids.ne(cfg.pad_token_id)
    position_ids = attention_mask.cumsum(-1) - 1
    position_ids[position_ids < 0] = 0
    position_ids = attention_mask.to(torch.long) * position_ids
    return {
        "input_ids": input_ids,
        "image_tiles": torch.cat(all_tiles, 0),
        "grid_thw": np.array(all_grid, dtype=np.int64),
        "attention_mask": attention_mask.to(torch.long),
        "position_ids": position_ids,
    }


# ------------------------------------------------------------------------------------------------ greedy loop
def greedy_decode(sd: SD, cfg, batch: dict, steps: int, dtype: torch.dtype, stop_rules: bool = False,
                  forced_tokens: Optional[torch.Tensor] = None, return_logits: bool = False):
    """One prefill + `steps-1` decode steps with a DynamicCache — the degenerate form of prediction_loop for a
    single full batch (surya/recognition/__init__.py:501-607, decode :326-352, prefill :354-471).
    Returns tokens [B, steps], scores [B, steps], boxes [B, steps, 6] (+ per-step logits when asked)."""
    cache = OracleCache()
    ids, mask, pos = batch["input_ids"], batch["attention_mask"], batch["position_ids"]
    tiles = batch["image_tiles"].to(dtype)
    toks, scores, boxes, logits_all = [], [], [], []
    B = ids.shape[0]
    active = [True] * B
    hist: List[List[int]] = [[] for _ in range(B)]
    with torch.inference_mode():
        lm, bb = model_forward(sd, cfg, ids, mask, pos, cache, tiles, batch["grid_thw"])
        for step in range(steps):
            nxt, preds, bx, done, sc = process_outputs(lm, bb, cfg)
            toks.append(preds[:, 0].clone())
            scores.append(sc[:, 0].clone())
            boxes.append(bx[:, 0].clone())
            if return_logits:
                logits_all.append(lm[:, 0].float().clone())
            if stop_rules:
                for b in range(B):
                    if not active[b]:
                        continue
                    hist[b].append(int(preds[b, 0]))
                    if step == 0:
                        if hist[b][-1] in (cfg.eos_token_id, cfg.no_output_token_id):
                            active[b] = False
                    elif hist[b][-1] in (cfg.eos_token_id, cfg.pad_token_id) or len(hist[b]) >= steps or \
                            detect_repeat_token(hist[b]):
                        active[b] = False
            if step == steps - 1:
                break
            if forced_tokens is not None:
                nxt = forced_tokens[:, step: step + 1].clone()
            mask = F.pad(mask, (0, 1), value=1)
            pos = pos[:, -1:] + 1
            lm, bb = model_forward(sd, cfg, nxt, mask, pos, cache)
    out = (torch.stack(toks, 1), torch.stack(scores, 1), torch.stack(boxes, 1))
    if stop_rules:
        out = out + (hist,)
    if return_logits:
        out = out + (torch.stack(logits_all, 1),)
    return out
