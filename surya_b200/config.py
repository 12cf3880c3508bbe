"""Shape configs for the recognition path (mirror of the reference's config classes).

Field names follow surya/common/surya/{config.py:8-71, encoder/config.py:7-53, decoder/config.py:8-85}.
The shipped hyper-parameters live in an S3 checkpoint that is not available offline, so the
benchmark uses the *declared synthetic* config SYN_REC; kernels are shape-generic.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass, field
from typing import Tuple


@dataclass
class RecEncoderConfig:
    depth: int = 8
    hidden_size: int = 1280
    intermediate_size: int = 3420
    num_heads: int = 16
    in_channels: int = 3
    patch_size: int = 14
    spatial_merge_size: int = 2
    temporal_patch_size: int = 1
    window_size: int = 112
    out_hidden_size: int = 1280
    fullatt_block_indexes: Tuple[int, ...] = (3, 7)

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads

    @property
    def patch_dim(self) -> int:
        return self.in_channels * self.temporal_patch_size * self.patch_size * self.patch_size


@dataclass
class RecDecoderConfig:
    hidden_size: int = 1280
    intermediate_size: int = 3420
    num_hidden_layers: int = 12
    num_attention_heads: int = 16
    num_key_value_heads: int = 4
    rope_theta: float = 10000.0
    rms_norm_eps: float = 1e-6

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


@dataclass
class RecConfig:
    vocab_size: int = 65792
    bbox_size: int = 1025
    bos_token_id: int = 0
    eos_token_id: int = 1
    pad_token_id: int = 2
    image_token_id: int = 3
    register_token_ids: Tuple[int, ...] = (4, 5, 6, 7)
    num_register_tokens: int = 4
    # synthetic stand-ins for ids that the real tokenizer table defines (processor/__init__.py:73-96)
    ocr_with_boxes_bos_id: int = 8
    eoi_token_id: int = 9
    no_output_token_id: int = 10
    nomath_token_id: int = 11
    image_embed_encoding_size: int = 1024
    image_embed_encoding_multiplier: int = 256
    max_sequence_length: int = 1536
    vision_encoder: RecEncoderConfig = field(default_factory=RecEncoderConfig)
    decoder: RecDecoderConfig = field(default_factory=RecDecoderConfig)

    @property
    def hidden_size(self) -> int:
        return self.decoder.hidden_size

    @property
    def merge_size(self) -> int:
        return self.vision_encoder.spatial_merge_size

    def to_dict(self) -> dict:
        return asdict(self)


def syn_rec() -> RecConfig:
    """SYN-REC: the declared synthetic recognition config of SURVEY.md §8(d) (BASELINE config 2)."""
    return RecConfig()


def tiny_rec() -> RecConfig:
    """Small config for fast CPU oracle / golden tests (same code paths: GQA 4:1, head_dim 80, odd I)."""
    enc = RecEncoderConfig(depth=2, hidden_size=160, intermediate_size=210, num_heads=2, out_hidden_size=320,
                           fullatt_block_indexes=(1,))
    dec = RecDecoderConfig(hidden_size=320, intermediate_size=428, num_hidden_layers=2, num_attention_heads=4,
                           num_key_value_heads=1)
    return RecConfig(vocab_size=1000, vision_encoder=enc, decoder=dec, image_embed_encoding_size=1024)


def align(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ------------------------------------------------------------------------------------------------ detection
@dataclass
class DetConfig:
    """Mirror of EfficientViTConfig (surya/detection/model/config.py:12-31)."""
    num_channels: int = 3
    widths: Tuple[int, ...] = (32, 64, 128, 256, 512)
    depths: Tuple[int, ...] = (1, 1, 1, 6, 6)
    strides: Tuple[int, ...] = (2, 2, 2, 2, 2)
    head_dim: int = 32
    layer_norm_eps: float = 1e-6
    decoder_layer_hidden_size: int = 128
    decoder_hidden_size: int = 512
    num_labels: int = 2
    head_bn_eps: float = 1e-5   # nn.BatchNorm2d default in DecodeHead (encoderdecoder.py:691)
    mla_eps: float = 1e-5       # LiteMLA eps (encoderdecoder.py:287)

    def to_dict(self) -> dict:
        return asdict(self)


def det_default() -> DetConfig:
    return DetConfig()


def det_tiny() -> DetConfig:
    """Reduced depth (same block types and channel widths) for fast CPU oracle tests."""
    return DetConfig(depths=(1, 1, 1, 2, 2))


# ------------------------------------------------------------------------------------------------ layout (Donut-Swin + ADETR)
@dataclass
class SwinConfig:
    """Mirror of DonutSwinLayoutConfig / DonutSwinTableRecConfig (surya/layout/model/config.py:85-106,
    surya/table_rec/model/config.py:89-111)."""
    image_size: Tuple[int, int] = (768, 768)
    patch_size: int = 4
    num_channels: int = 3
    embed_dim: int = 128
    depths: Tuple[int, ...] = (2, 2, 16, 2)
    num_heads: Tuple[int, ...] = (4, 8, 16, 32)
    window_size: int = 8
    mlp_ratio: float = 4.0
    layer_norm_eps: float = 1e-5
    encoder_length: int = 768

    @property
    def hidden_size(self) -> int:
        return self.embed_dim * 2 ** (len(self.depths) - 1)

    @property
    def grid(self) -> Tuple[int, int]:
        return (self.image_size[0] // self.patch_size, self.image_size[1] // self.patch_size)


@dataclass
class AdetrConfig:
    """Mirror of SuryaLayoutDecoderConfig (surya/layout/model/config.py:138-179)."""
    num_hidden_layers: int = 8
    vocab_size: int = 1025
    bbox_size: int = 1024
    label_count: int = 20
    skew_scaler: int = 512
    special_token_count: int = 3
    hidden_size: int = 1024
    intermediate_size: int = 4096
    encoder_hidden_size: int = 1024
    num_attention_heads: int = 16
    num_key_value_heads: int = 4
    rms_norm_eps: float = 1e-6
    layer_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    pad_token_id: int = 0
    eos_token_id: int = 1
    bos_token_id: int = 1
    pause_token_id: int = 2
    pause_token_count: int = 0
    double_residual_flow: bool = True
    max_boxes: int = 100
    # table_rec variant (SuryaTableRecDecoderConfig, surya/table_rec/model/config.py:142-226): LabelEmbedding widths, property
    # head sizes (classification counts include the 5 special tokens), 10-column tokens
    kind: str = "layout"
    box_embed_size: int = 0
    property_embed_size: int = 0
    category_count: int = 0
    merge_count: int = 0
    header_count: int = 0
    query_end_token_id: int = 4

    @property
    def token_width(self) -> int:
        return 7 if self.kind == "layout" else 10

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


@dataclass
class LayoutConfig:
    encoder: SwinConfig = field(default_factory=SwinConfig)
    decoder: AdetrConfig = field(default_factory=AdetrConfig)

    def to_dict(self) -> dict:
        return asdict(self)


def layout_default() -> LayoutConfig:
    return LayoutConfig()


def table_decoder(num_hidden_layers: int = 6) -> AdetrConfig:
    return AdetrConfig(num_hidden_layers=num_hidden_layers, hidden_size=512, intermediate_size=2048, num_attention_heads=8,
                       num_key_value_heads=4, special_token_count=5, double_residual_flow=False, max_boxes=150, kind="table",
                       box_embed_size=448, property_embed_size=64, category_count=10, merge_count=9, header_count=7, label_count=0)


def table_default() -> LayoutConfig:
    """DonutSwinTableRecConfig + SuryaTableRecDecoderConfig defaults (surya/table_rec/model/config.py:89-226)."""
    return LayoutConfig(encoder=SwinConfig(depths=(2, 2, 12, 2), encoder_length=1024), decoder=table_decoder())


def table_tiny() -> LayoutConfig:
    return LayoutConfig(encoder=SwinConfig(image_size=(256, 256), depths=(2, 2, 2, 2), encoder_length=64), decoder=table_decoder(2))


def layout_tiny() -> LayoutConfig:
    """Shallow variant (same widths, windows and head dims) for fast CPU oracle tests; 256x256 input."""
    enc = SwinConfig(image_size=(256, 256), depths=(2, 2, 2, 2), encoder_length=64)
    dec = AdetrConfig(num_hidden_layers=2)
    return LayoutConfig(encoder=enc, decoder=dec)


# ------------------------------------------------------------------------------------------------ ocr_error (DistilBERT)
@dataclass
class OcrErrorConfig:
    """Mirror of DistilBertConfig (surya/ocr_error/model/config.py:13-52) for DistilBertForSequenceClassification."""
    vocab_size: int = 30522
    max_position_embeddings: int = 512
    n_layers: int = 6
    n_heads: int = 12
    dim: int = 768
    hidden_dim: int = 3072
    num_labels: int = 2
    pad_token_id: int = 0
    layer_norm_eps: float = 1e-12   # nn.LayerNorm(eps=1e-12) in Embeddings / TransformerBlock (encoder.py:54, 417, 420)

    def to_dict(self) -> dict:
        return asdict(self)


def ocr_error_default() -> OcrErrorConfig:
    return OcrErrorConfig()


def ocr_error_tiny() -> OcrErrorConfig:
    """Two layers, 4 heads of 64, small vocabulary: fast CPU oracle tests."""
    return OcrErrorConfig(vocab_size=1000, max_position_embeddings=128, n_layers=2, n_heads=4, dim=256, hidden_dim=512)
