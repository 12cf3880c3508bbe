"""CPU ORACLE (test infrastructure, not product code) — layout path: Donut-Swin encoder + ADETR box decoder.

Functional PyTorch restatement of DonutSwinLayoutModel.forward (surya/layout/model/encoder.py:33-81 over
surya/common/donut/encoder.py) and SuryaLayoutDecoder.forward (surya/layout/model/decoder.py:95-126 over
surya/common/adetr/decoder.py), plus the predictor's greedy box loop (surya/layout/__init__.py:111-183, without the
host-side header/footer relabelling).  Pinned by tests/golden/layout_*.pt (oracle/make_golden.py).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------------------------ Swin encoder
def _ln(sd, name, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def window_partition(x, ws):
    """surya/common/donut/encoder.py window_partition: [B,H,W,C] -> [B*nW, ws, ws, C]."""
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def window_reverse(w, ws, H, W):
    C = w.shape[-1]
    x = w.view(-1, H // ws, W // ws, ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, H, W, C)


def relative_position_index(ws: int) -> torch.Tensor:
    """DonutSwinSelfAttention.__init__ — encoder.py:348-359."""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def shift_attn_mask(H, W, ws, shift, dtype):
    """DonutSwinLayer.get_attn_mask — encoder.py:562-590 (-100 between different regions)."""
    img = torch.zeros((1, H, W, 1), dtype=dtype)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = window_partition(img, ws).view(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, float(-100.0)).masked_fill(m == 0, float(0.0))


def sincos_2d(width, height, dim, temperature=10000.0):
    """DonutSwinStage.build_2d_sincos_position_embedding — encoder.py:736-761."""
    gw = torch.arange(int(width), dtype=torch.float32)
    gh = torch.arange(int(height), dtype=torch.float32)
    gw, gh = torch.meshgrid(gw, gh, indexing="ij")
    pos_dim = dim // 4
    omega = 1.0 / (temperature ** (torch.arange(pos_dim, dtype=torch.float32) / pos_dim))
    ow = gw.flatten()[..., None] @ omega[None]
    oh = gh.flatten()[..., None] @ omega[None]
    return torch.concat([ow.sin(), ow.cos(), oh.sin(), oh.cos()], dim=1)[None]


def swin_layer(sd: SD, cfg, p: str, x, H, W, nh, shift):
    """DonutSwinLayer.forward — encoder.py:598-685 (+ DonutSwinSelfAttention :383-442, SelfOutput, Intermediate, Output)."""
    B, _, C = x.shape
    ws = cfg.window_size
    if min(H, W) <= ws:       # set_shift_and_window_size :550-560
        shift, ws = 0, min(H, W)
    shortcut = x
    h = _ln(sd, p + "layernorm_before", x, cfg.layer_norm_eps).view(B, H, W, C)
    pad_r, pad_b = (ws - W % ws) % ws, (ws - H % ws) % ws
    h = F.pad(h, (0, 0, 0, pad_r, 0, pad_b))
    Hp, Wp = H + pad_b, W + pad_r
    if shift > 0:
        h = torch.roll(h, shifts=(-shift, -shift), dims=(1, 2))
    hw = window_partition(h, ws).view(-1, ws * ws, C)
    mask = shift_attn_mask(Hp, Wp, ws, shift, h.dtype) if shift > 0 else None
    nW = hw.shape[0]
    hd = C // nh
    a = p + "attention.self."
    q = F.linear(hw, sd[a + "query.weight"], sd[a + "query.bias"]).view(nW, -1, nh, hd).permute(0, 2, 1, 3)
    k = F.linear(hw, sd[a + "key.weight"], sd[a + "key.bias"]).view(nW, -1, nh, hd).permute(0, 2, 1, 3)
    v = F.linear(hw, sd[a + "value.weight"], sd[a + "value.bias"]).view(nW, -1, nh, hd).permute(0, 2, 1, 3)
    rpi = relative_position_index(cfg.window_size)
    bias = sd[a + "relative_position_bias_table"][rpi.view(-1)].view(cfg.window_size ** 2, cfg.window_size ** 2, -1)
    bias = bias.permute(2, 0, 1).contiguous().unsqueeze(0)
    am = bias.repeat(nW, 1, 1, 1) if mask is None else mask.repeat(nW // mask.shape[0], 1, 1).unsqueeze(1) + bias
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=am.to(q.dtype), dropout_p=0.0, scale=hd ** -0.5)
    o = o.transpose(1, 2).contiguous().view(nW, ws * ws, C)
    o = F.linear(o, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
    o = window_reverse(o.view(-1, ws, ws, C), ws, Hp, Wp)
    if shift > 0:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    if pad_r or pad_b:
        o = o[:, :H, :W, :].contiguous()
    x = shortcut + o.view(B, H * W, C)
    y = _ln(sd, p + "layernorm_after", x, cfg.layer_norm_eps)
    y = F.gelu(F.linear(y, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
    return x + F.linear(y, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])


def patch_merging(sd: SD, p: str, x, H, W):
    """DonutSwinPatchMerging.forward — encoder.py:289-318."""
    B, _, C = x.shape
    f = x.view(B, H, W, C)
    f = F.pad(f, (0, 0, 0, W % 2, 0, H % 2))
    f = torch.cat([f[:, 0::2, 0::2, :], f[:, 1::2, 0::2, :], f[:, 0::2, 1::2, :], f[:, 1::2, 1::2, :]], -1)
    f = f.view(B, -1, 4 * C)
    f = F.layer_norm(f, (4 * C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    return F.linear(f, sd[p + "reduction.weight"])


def swin_forward(sd: SD, cfg, pixel_values: torch.Tensor, prefix: str = "") -> torch.Tensor:
    """DonutSwinLayoutModel.forward — layout/model/encoder.py:33-81 -> [B, HW_last, hidden]."""
    p = prefix
    with torch.inference_mode():
        x = F.conv2d(pixel_values, sd[p + "embeddings.patch_embeddings.projection.weight"],
                     sd[p + "embeddings.patch_embeddings.projection.bias"], stride=cfg.patch_size)
        H, W = x.shape[2:]
        x = x.flatten(2).transpose(1, 2)
        x = _ln(sd, p + "embeddings.norm", x, 1e-5)
        gh, gw = cfg.grid
        for s, (depth, nh) in enumerate(zip(cfg.depths, cfg.num_heads)):
            C = cfg.embed_dim * 2 ** s
            # positional encoding sized from the config grid at construction (encoder.py:730-735, 773-776)
            x = x + sincos_2d(gw // 2 ** s, gh // 2 ** s, C).to(x.dtype)
            for b in range(depth):
                x = swin_layer(sd, cfg, f"{p}encoder.layers.{s}.blocks.{b}.", x, H, W, nh, 0 if b % 2 == 0 else cfg.window_size // 2)
            if s < len(cfg.depths) - 1:
                x = patch_merging(sd, f"{p}encoder.layers.{s}.downsample.", x, H, W)
                H, W = (H + 1) // 2, (W + 1) // 2
        return x + sd[p + "position_embeddings"][:, : x.shape[1], :]


# ------------------------------------------------------------------------------------------------ ADETR decoder
def adetr_rmsnorm(x, w, eps):
    """SuryaADETRDecoderRMSNorm — adetr/decoder.py:29-47."""
    xf = x.float()
    var = torch.clamp(xf.pow(2).mean(-1, keepdim=True), min=eps)
    out = xf * torch.rsqrt(var) * (1.0 + w.float())
    info = torch.finfo(x.dtype)
    out = out.clamp(min=info.min, max=info.max)
    out = torch.where(torch.isnan(out), torch.tensor(0.0), out)
    return out.type_as(x)


def bbox_embedding(sd: SD, cfg, boxes: torch.Tensor) -> torch.Tensor:
    """BboxEmbedding.forward — layout/model/decoder.py:36-57."""
    e = lambda n, i: sd[f"model.embed_tokens.{n}_embed.weight"][i]
    cx, cy, w, h, xs, ys, label = boxes.to(torch.long).unbind(dim=-1)
    xa = ((xs - cfg.bbox_size // 2) / 2).to(torch.long)
    ya = ((ys - cfg.bbox_size // 2) / 2).to(torch.long)
    cl = lambda t: t.clamp(0, cfg.bbox_size).to(torch.long)
    x1, y1 = cl(cx - w // 2 - xa), cl(cy - h // 2 - ya)
    x2, y2 = cl(cx + w // 2 - xa), cl(cy + h // 2 + ya)
    x3, y3 = cl(cx + w // 2 + xa), cl(cy + h // 2 + ya)
    x4, y4 = cl(cx - w // 2 + xa), cl(cy - h // 2 - ya)
    size = e("w", w) + e("h", h) + e("cx", cx) + e("cy", cy)
    skew = e("xskew", xs) + e("yskew", ys)
    corner = e("x1", x1) + e("y1", y1) + e("x2", x2) + e("y2", y2) + e("x3", x3) + e("y3", y3) + e("x4", x4) + e("y4", y4)
    return e("label", label) + size + skew + corner


def label_embedding(sd: SD, cfg, boxes: torch.Tensor) -> torch.Tensor:
    """LabelEmbedding.forward — table_rec/model/decoder.py:46-73 (columns: cx cy w h xskew yskew | category merges colspan
    is_header, shaper.py:54-80; is_header is not embedded)."""
    e = lambda n, i: sd[f"model.embed_tokens.{n}_embed.weight"][i]
    boxes = boxes.to(torch.long).clamp(0, cfg.vocab_size)
    cx, cy, w, h, xs, ys, cat, mrg, col, _hdr = boxes.unbind(dim=-1)
    xa = ((xs - cfg.bbox_size // 2) / 2).to(torch.long)
    ya = ((ys - cfg.bbox_size // 2) / 2).to(torch.long)
    cl = lambda t: t.clamp(0, cfg.bbox_size).to(torch.long)
    x1, y1 = cl(cx - w // 2 - xa), cl(cy - h // 2 - ya)
    x3, y3 = cl(cx + w // 2 + xa), cl(cy + h // 2 + ya)
    size = e("w", w) + e("h", h) + e("cx", cx) + e("cy", cy)
    skew = e("xskew", xs) + e("yskew", ys)
    corner = e("x1", x1) + e("y1", y1) + e("x3", x3) + e("y3", y3)
    box = size + skew + corner
    prop = e("category", cat) + e("merge", mrg) + e("colspan", col)
    return torch.cat([box, prop], dim=-1)


def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), -1)


class AdetrState:
    """Per-layer caches: cross K/V (computed once) and self K/V (concatenated per step)."""

    def __init__(self, n_layers):
        self.ck = [None] * n_layers
        self.cv = [None] * n_layers
        self.sk = [None] * n_layers
        self.sv = [None] * n_layers


def adetr_forward(sd: SD, cfg, boxes: torch.Tensor, enc: torch.Tensor, cache_position: torch.Tensor, st: AdetrState):
    """SuryaLayoutDecoder.forward — layout/model/decoder.py:95-126; layers: adetr/decoder.py:380-456 (double_res_forward
    for layout), cross attention :150-194, self attention :238-290 (dynamic cache), MLP :342-357, mask :620-637."""
    B, q, _ = boxes.shape
    nh, nkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    x = label_embedding(sd, cfg, boxes) if cfg.kind == "table" else bbox_embedding(sd, cfg, boxes)
    dt = x.dtype
    pos = cache_position.unsqueeze(0)
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
    freqs = (inv[None, :, None].float().expand(1, -1, 1) @ pos[:, None, :].float()).transpose(1, 2)
    emb = torch.cat((freqs, freqs), -1)
    cos, sin = emb.cos().to(dt).unsqueeze(1), emb.sin().to(dt).unsqueeze(1)
    rep = nh // nkv
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        raw = x
        # ---- cross attention (no mask, K/V of the encoder states cached at the first call)
        n = adetr_rmsnorm(x, sd[p + "cross_pre_norm.weight"], cfg.rms_norm_eps)
        cq = F.linear(n, sd[p + "cross_attn_block.q_proj.weight"]).view(B, q, nh, hd).transpose(1, 2)
        if st.ck[l] is None:
            st.ck[l] = F.linear(enc, sd[p + "cross_attn_block.k_proj.weight"]).view(B, -1, nkv, hd).transpose(1, 2)
            st.cv[l] = F.linear(enc, sd[p + "cross_attn_block.v_proj.weight"]).view(B, -1, nkv, hd).transpose(1, 2)
        ck = st.ck[l].repeat_interleave(rep, 1)
        cv = st.cv[l].repeat_interleave(rep, 1)
        a = F.scaled_dot_product_attention(cq, ck, cv, scale=hd ** -0.5).transpose(1, 2).reshape(B, q, nh * hd)
        cross = F.linear(a, sd[p + "cross_attn_block.o_proj.weight"], sd[p + "cross_attn_block.o_proj.bias"]) + raw
        # ---- causal self attention with RoPE
        n = adetr_rmsnorm(cross, sd[p + "temporal_pre_norm.weight"], cfg.rms_norm_eps)
        sq = F.linear(n, sd[p + "temporal_block.q_proj.weight"]).view(B, q, nh, hd).transpose(1, 2)
        sk = F.linear(n, sd[p + "temporal_block.k_proj.weight"]).view(B, q, nkv, hd).transpose(1, 2)
        sv = F.linear(n, sd[p + "temporal_block.v_proj.weight"]).view(B, q, nkv, hd).transpose(1, 2)
        sq = (sq * cos) + (_rot_half(sq) * sin)
        sk = (sk * cos) + (_rot_half(sk) * sin)
        st.sk[l] = sk if st.sk[l] is None else torch.cat([st.sk[l], sk], 2)
        st.sv[l] = sv if st.sv[l] is None else torch.cat([st.sv[l], sv], 2)
        kk, vv = st.sk[l].repeat_interleave(rep, 1), st.sv[l].repeat_interleave(rep, 1)
        S = kk.shape[2]
        mask = torch.zeros(q, S, dtype=dt)
        mask = mask.masked_fill(torch.arange(S)[None, :] > cache_position[:, None], torch.finfo(dt).min)
        a = F.scaled_dot_product_attention(sq, kk, vv, attn_mask=mask[None, None], scale=hd ** -0.5)
        a = a.transpose(1, 2).reshape(B, q, nh * hd)
        temporal = F.linear(a, sd[p + "temporal_block.o_proj.weight"], sd[p + "temporal_block.o_proj.bias"])
        res = (temporal + raw) if cfg.double_residual_flow else (temporal + cross)
        n = adetr_rmsnorm(res, sd[p + "channel_pre_norm.weight"], cfg.rms_norm_eps)
        m = F.gelu(F.linear(n, sd[p + "mlp_block.gate_proj.weight"]), approximate="tanh") * F.linear(n, sd[p + "mlp_block.up_proj.weight"])
        x = F.linear(m, sd[p + "mlp_block.down_proj.weight"]) + res
    x = adetr_rmsnorm(x, sd["model.final_norm.weight"], cfg.rms_norm_eps)
    h = F.layer_norm(x, (x.shape[-1],), sd["pre_output_norm.weight"], sd["pre_output_norm.bias"], cfg.layer_norm_eps)
    if cfg.kind == "table":      # SuryaTableRecDecoder.forward — table_rec/model/decoder.py:121-155
        out = {k: F.linear(h, sd[f"box_property_heads.{k}.weight"]) for k in ("bbox", "category", "merges", "colspan", "is_header")}
        out["bbox"] = torch.sigmoid(out["bbox"])
        return out
    return torch.sigmoid(F.linear(h, sd["bbox_head.weight"], sd["bbox_head.bias"])), F.linear(h, sd["lm_head.weight"])


def layout_greedy(sd_enc: SD, sd_dec: SD, cfg, pixel_values: torch.Tensor, steps: int, return_logits: bool = False):
    """Encoder once + `steps` greedy box-decoding steps (surya/layout/__init__.py:83-137, 183): returns
    tokens [B, steps, 7] (6 box coords + class id) and optionally the per-step (bbox, class) logits."""
    d = cfg.decoder
    with torch.inference_mode():
        enc = swin_forward(sd_enc, cfg.encoder, pixel_values)
        B = pixel_values.shape[0]
        boxes = torch.full((B, 1 + d.pause_token_count, 7), d.bos_token_id, dtype=torch.long)
        if d.pause_token_count:
            boxes[:, 1:] = d.pause_token_id
        cache_position = torch.arange(boxes.shape[1])
        st = AdetrState(d.num_hidden_layers)
        toks, bl, cl = [], [], []
        for _ in range(steps):
            bbox, cls = adetr_forward(sd_dec, d, boxes, enc, cache_position, st)
            cache_position = cache_position[-1:] + 1
            b, c = bbox[:, -1, :], cls[:, -1, :]
            preds = c.argmax(-1)
            boxes = torch.cat([(b * d.bbox_size).unsqueeze(1), preds.unsqueeze(1).unsqueeze(1)], dim=-1).to(torch.long)
            toks.append(boxes[:, 0].clone())
            bl.append(b.float().clone())
            cl.append(c.float().clone())
    out = (torch.stack(toks, 1), enc)
    if return_logits:
        out = out + (torch.stack(bl, 1), torch.stack(cl, 1))
    return out


def table_next_tokens(out: dict, cfg):
    """TableRecPredictor.inference_loop's per-step processing (table_rec/__init__.py:76-121) + LabelShaper.dict_to_labels
    (shaper.py:12-52) + torch.tensor(..., dtype=long): -> tokens [B, 10] int64, done [B] bool."""
    last = {k: v[:, -1, :] for k, v in out.items()}
    cat = last["category"].argmax(-1)
    done = (cat == cfg.eos_token_id) | (cat == cfg.pad_token_id)
    bbox = (last["bbox"] * cfg.bbox_size).float().clamp(0, cfg.bbox_size)
    colspan = torch.round(torch.clamp(last["colspan"], min=1))[:, 0]
    tok = torch.cat([bbox.to(torch.long), cat[:, None], last["merges"].argmax(-1)[:, None], colspan.to(torch.long)[:, None],
                     last["is_header"].argmax(-1)[:, None]], dim=1)
    return tok, done


def table_greedy(sd_enc: SD, sd_dec: SD, cfg, pixel_values: torch.Tensor, prompt: torch.Tensor, steps: int):
    """Encoder + row/column decoding pass of TableRecPredictor (table_rec/__init__.py:33-131, 180-190): multi-token prompt
    prefill, then `steps` greedy box tokens.  Returns tokens [B, steps, 10], done [B, steps], encoder states, and the per-step
    head outputs (dict of [B, steps, n])."""
    d = cfg.decoder
    with torch.inference_mode():
        enc = swin_forward(sd_enc, cfg.encoder, pixel_values)
        ids = prompt.clone()
        cache_position = torch.arange(ids.shape[1])
        st = AdetrState(d.num_hidden_layers)
        toks, dones, outs = [], [], []
        for _ in range(steps):
            out = adetr_forward(sd_dec, d, ids, enc, cache_position, st)
            cache_position = cache_position[-1:] + 1
            tok, done = table_next_tokens(out, d)
            ids = tok.unsqueeze(1)
            toks.append(tok)
            dones.append(done)
            outs.append({k: v[:, -1].float().clone() for k, v in out.items()})
    heads = {k: torch.stack([o[k] for o in outs], 1) for k in outs[0]}
    return torch.stack(toks, 1), torch.stack(dones, 1), enc, heads
