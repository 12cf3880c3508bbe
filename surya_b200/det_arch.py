"""Structure of the detection network (EfficientViT-L segmentation) derived from DetConfig.

Mirrors how the reference assembles its modules (surya/detection/model/encoderdecoder.py:425-481 build_local_block,
:484-511 Stem, :514-577 EfficientVitLargeStage, :580-630 EfficientVitLarge, :673-722 DecodeHead) as plain data:
a list of blocks, each a list of conv specs with the reference's state_dict prefix.  Used by the synthetic
weight generator, the weight packer (BN folding) and the oracle so that all three agree on names and shapes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

from .config import DetConfig


@dataclass
class ConvSpec:
    name: str            # state_dict prefix of the ConvNormAct (…conv.weight / …norm.*), or of a bare nn.Conv2d
    cin: int
    cout: int
    k: int = 1
    stride: int = 1
    groups: int = 1
    bias: bool = False
    norm: bool = False
    eps: float = 1e-6
    act: str = "none"    # "hardswish" | "relu" | "none"
    bare: bool = False   # nn.Conv2d without the ConvNormAct wrapper (names: name.weight / name.bias)

    @property
    def pad(self) -> int:
        # get_padding(kernel, stride, dilation=1) — encoderdecoder.py:48-50
        return ((self.stride - 1) + (self.k - 1)) // 2

    @property
    def wkey(self) -> str:
        return f"{self.name}.weight" if self.bare else f"{self.name}.conv.weight"

    @property
    def bkey(self) -> str:
        return f"{self.name}.bias" if self.bare else f"{self.name}.conv.bias"


@dataclass
class Block:
    kind: str                    # "conv" | "convblock" | "fused" | "mb" | "vit"
    convs: List[ConvSpec]
    residual: bool = False
    stage: int = -1              # index into the 4 feature stages (-1 = stem)
    mla: Optional[List[ConvSpec]] = None   # vit only: [qkv, aggreg dw5x5, aggreg grouped 1x1, proj]
    heads: int = 0               # vit only: heads per scale (total = 2 * heads)
    dim: int = 0


def det_blocks(cfg: DetConfig) -> List[Block]:
    eps, act = cfg.layer_norm_eps, "hardswish"
    blocks: List[Block] = []
    w0 = cfg.widths[0]
    s0 = cfg.strides[0]
    blocks.append(Block("conv", [ConvSpec("vit.stem.in_conv", cfg.num_channels, w0, s0 + 1, s0, norm=True, eps=eps, act=act)]))
    for i in range(cfg.depths[0]):
        p = f"vit.stem.res{i}.main"
        blocks.append(Block("convblock", [ConvSpec(f"{p}.conv1", w0, w0, 3, 1, norm=True, eps=eps, act=act),
                                          ConvSpec(f"{p}.conv2", w0, w0, 3, 1, norm=True, eps=eps)], residual=True))
    cin = w0
    for si, (w, d, s) in enumerate(zip(cfg.widths[1:], cfg.depths[1:], cfg.strides[1:])):
        vit = si >= 3
        fewer = si >= 2

        def local(prefix, ci, co, stride, k, expand, fewer_norm, fused):
            mid = round(ci * expand)
            if fused:
                return Block("fused", [ConvSpec(f"{prefix}.spatial_conv", ci, mid, k, stride, norm=True, eps=eps, act=act),
                                       ConvSpec(f"{prefix}.point_conv", mid, co, 1, 1, norm=True, eps=eps)])
            return Block("mb", [ConvSpec(f"{prefix}.inverted_conv", ci, mid, 1, 1, bias=fewer_norm, norm=not fewer_norm, eps=eps, act=act),
                                ConvSpec(f"{prefix}.depth_conv", mid, mid, k, stride, groups=mid, bias=fewer_norm, norm=not fewer_norm, eps=eps, act=act),
                                ConvSpec(f"{prefix}.point_conv", mid, co, 1, 1, norm=True, eps=eps)])

        b0 = local(f"vit.stages.{si}.blocks.0.main", cin, w, s, s + 1, 24 if vit else 16, vit or fewer, not fewer)
        b0.stage = si
        blocks.append(b0)
        cin = w
        for bi in range(1, d + 1):
            if vit:
                p = f"vit.stages.{si}.blocks.{bi}"
                heads = w // cfg.head_dim
                td = heads * cfg.head_dim
                mla = [ConvSpec(f"{p}.context_module.main.qkv", w, 3 * td, 1),
                       ConvSpec(f"{p}.context_module.main.aggreg.0.0", 3 * td, 3 * td, 5, 1, groups=3 * td, bare=True),
                       ConvSpec(f"{p}.context_module.main.aggreg.0.1", 3 * td, 3 * td, 1, 1, groups=3 * heads, bare=True),
                       ConvSpec(f"{p}.context_module.main.proj", 2 * td, w, 1, norm=True, eps=eps)]
                mb = local(f"{p}.local_module.main", w, w, 1, 3, 6, True, False)
                blocks.append(Block("vit", mb.convs, residual=True, stage=si, mla=mla, heads=heads, dim=cfg.head_dim))
            else:
                b = local(f"vit.stages.{si}.blocks.{bi}.main", w, w, 1, 3, 4, fewer, not fewer)
                b.residual, b.stage = True, si
                blocks.append(b)
    return blocks


def det_head_specs(cfg: DetConfig):
    """DecodeHead parameter shapes (encoderdecoder.py:673-697)."""
    dl, dh = cfg.decoder_layer_hidden_size, cfg.decoder_hidden_size
    n_stage = len(cfg.widths) - 1
    return {
        "linear_c": [(f"decode_head.linear_c.{i}.proj", w, dl) for i, w in enumerate(cfg.widths[1:])],
        "fuse": ("decode_head.linear_fuse", dl * n_stage, dh),
        "bn": "decode_head.batch_norm",
        "cls": ("decode_head.classifier", dh, cfg.num_labels),
    }
