"""ORACLE TOOLING — generate golden vectors from the reference's own nn.Modules (run in the build container).

    python -m oracle.make_golden            # writes tests/golden/rec_*.pt

The reference (imported unmodified from /root/reference through oracle/ref_shim.py) is instantiated for a
declared config, loaded with surya_b200.synth's seeded weights, and driven exactly like
RecognitionPredictor.prefill/decode drive it (surya/recognition/__init__.py:326-352, 398-409): one prefill
with a fresh cache, then greedy decode steps feeding process_outputs' input_ids back.  Outputs are stored in
fp32; inputs are regenerated from seeds by the tests, so the fixtures stay small.
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import rec_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402
from surya_b200.config import syn_rec, tiny_rec  # noqa: E402
from surya_b200.synth import rec_state_dict, rec_synthetic_crops  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"


def golden_crops(kind: str):
    """The seeded inputs the goldens are defined on (tests regenerate them with this same function)."""
    if kind == "tiny":
        crops = list(rec_synthetic_crops(3, 48, 512, seed=1234))
        crops.append(rec_synthetic_crops(1, 40, 300, seed=5)[0])    # shorter prompt -> left padding
        crops.append(rec_synthetic_crops(1, 64, 900, seed=6)[0])    # wider crop -> more windows
        return crops
    if kind == "synrec":
        return list(rec_synthetic_crops(2, 48, 512, seed=1234))
    raise ValueError(kind)


def run_reference(cfg, sd, batch, steps: int, attn: str = "sdpa"):
    from transformers import DynamicCache

    model = ref_shim.build_reference_rec_model(cfg, sd, attn=attn)
    ids, mask, pos = batch["input_ids"], batch["attention_mask"], batch["position_ids"]
    lms, bbs = [], []
    with torch.inference_mode():
        cache = DynamicCache()
        out = model(input_ids=ids, image_tiles=batch["image_tiles"], grid_thw=torch.from_numpy(batch["grid_thw"]),
                    attention_mask=mask, position_ids=pos, inputs_embeds=None, past_key_values=cache, use_cache=True,
                    logits_to_keep=1, encoder_chunk_size=4096)
        for step in range(steps):
            lm, bb = out["lm_logits"], out["bbox_logits"]
            lms.append(lm[:, 0].float().clone())
            bbs.append(bb[:, 0].float().clone())
            if step == steps - 1:
                break
            nxt, *_ = O.process_outputs(lm, bb, cfg)
            mask = F.pad(mask, (0, 1), value=1)
            pos = pos[:, -1:] + 1
            out = model(input_ids=nxt, attention_mask=mask, position_ids=pos, use_cache=True, past_key_values=cache,
                        logits_to_keep=1)
    return torch.stack(lms, 1), torch.stack(bbs, 1)


def summarise(lm: torch.Tensor, bb: torch.Tensor, cfg, full_logits: bool):
    tok = lm.argmax(-1)
    top2 = lm.topk(2, dim=-1).values
    g = {
        "tokens": tok,
        "margin": (top2[..., 0] - top2[..., 1]),
        "score": lm.softmax(-1).max(-1).values,
        "logsumexp": lm.logsumexp(-1),
        "bbox": bb,
        "boxes": (bb * cfg.bbox_size).to(torch.long),
        "logit_idx": torch.arange(0, lm.shape[-1], 97),
    }
    g["logit_sample"] = lm[..., g["logit_idx"]].clone()
    g["logit_max"] = lm.max(-1).values
    if full_logits:
        g["logits"] = lm
    return g


def make_layout_golden():
    """Reference Swin encoder + ADETR decoder driven like LayoutPredictor.batch_layout_detection
    (surya/layout/__init__.py:83-137): encoder once, prefill call, greedy box steps."""
    from surya_b200.config import layout_tiny
    from surya_b200.synth import adetr_layout_state_dict, layout_synthetic_pages, swin_state_dict

    cfg = layout_tiny()
    e, d = cfg.encoder, cfg.decoder
    sde, sdd = swin_state_dict(e, 0), adetr_layout_state_dict(d, 0)
    enc, dec = ref_shim.build_reference_layout_models(cfg, sde, sdd)
    x = layout_synthetic_pages(2, e.image_size, seed=7)
    steps = 8
    with torch.inference_mode():
        ref_enc = enc(pixel_values=x)[0]
        dec.model._setup_cache(dec.config, 2, "cpu", torch.float32)
        boxes = torch.full((2, 1, 7), d.bos_token_id, dtype=torch.long)
        pos = torch.arange(1)
        toks, bbs, cls = [], [], []
        for s in range(steps):
            out = dec(input_boxes=boxes, encoder_hidden_states=ref_enc, cache_position=pos, use_cache=True, prefill=(s == 0))
            pos = pos[-1:] + 1
            b, c = out["bbox_logits"][:, -1, :], out["class_logits"][:, -1, :]
            boxes = torch.cat([(b * d.bbox_size).unsqueeze(1), c.argmax(-1).unsqueeze(1).unsqueeze(1)], dim=-1).to(torch.long)
            toks.append(boxes[:, 0].clone())
            bbs.append(b.float().clone())
            cls.append(c.float().clone())
    g = {"encoder": ref_enc.float().clone(), "tokens": torch.stack(toks, 1), "bbox": torch.stack(bbs, 1),
         "class_logits": torch.stack(cls, 1), "input_checksum": x.double().sum(),
         "meta": {"kind": "layout_tiny", "steps": steps, "seed": 0, "page_seed": 7, "torch": str(torch.__version__),
                  "reference": "VikParuchuri/surya@80e9a7e (v0.14.6), fp32 CPU, eager attention"}}
    torch.save(g, GOLDEN / "layout_tiny.pt")
    top2 = g["class_logits"].topk(2, -1).values
    print(f"[golden] layout_tiny: tokens[0,:2]={g['tokens'][0, :2].tolist()} min class margin={(top2[..., 0] - top2[..., 1]).min():.4f}")


def make_table_golden():
    """Reference table_rec encoder + decoder driven like TableRecPredictor.inference_loop (table_rec/__init__.py:33-131):
    3-token query prompt prefill, then greedy steps with the predictor's own token formation."""
    from oracle import layout_oracle as L
    from surya_b200.config import table_tiny
    from surya_b200.synth import adetr_table_state_dict, layout_synthetic_pages, swin_state_dict, table_query_tokens

    cfg = table_tiny()
    e, d = cfg.encoder, cfg.decoder
    sde, sdd = swin_state_dict(e, 1), adetr_table_state_dict(d, 1)
    enc, dec = ref_shim.build_reference_table_models(cfg, sde, sdd)
    x = layout_synthetic_pages(2, e.image_size, seed=9)
    ids = table_query_tokens(d, 2)
    steps = 8
    with torch.inference_mode():
        ref_enc = enc(pixel_values=x).last_hidden_state
        dec.model._setup_cache(dec.config, 2, "cpu", torch.float32)
        pos = torch.ones_like(ids[0, :, 0], dtype=torch.int64).cumsum(0) - 1
        toks, heads = [], []
        for s in range(steps):
            out = dec(input_ids=ids, encoder_hidden_states=ref_enc, cache_position=pos, use_cache=True, prefill=(s == 0))
            pos = pos[-1:] + 1
            logits = out["box_property_logits"]
            tok, done = L.table_next_tokens(logits, d)
            ids = tok.unsqueeze(1)
            toks.append(tok)
            heads.append({k: v[:, -1].float().clone() for k, v in logits.items()})
    g = {"encoder": ref_enc.float().clone(), "tokens": torch.stack(toks, 1),
         "heads": {k: torch.stack([h[k] for h in heads], 1) for k in heads[0]}, "input_checksum": x.double().sum(),
         "meta": {"kind": "table_tiny", "steps": steps, "seed": 1, "page_seed": 9, "torch": str(torch.__version__),
                  "reference": "VikParuchuri/surya@80e9a7e (v0.14.6), fp32 CPU, eager attention"}}
    torch.save(g, GOLDEN / "table_tiny.pt")
    print(f"[golden] table_tiny: tokens[0,:2]={g['tokens'][0, :2].tolist()}")


def make_layout_variants_golden():
    """Extra pins for the oracle only (CPU test): a NON-SQUARE Swin input (256x512: exercises the (W, H) order of the
    sin-cos tables, window partition of unequal sides and the shift masks) and a table_rec CELL-pass prompt (query + 4 column
    boxes, q_len = 7 prefill; table_rec/__init__.py:206-222, processor.py:78-82)."""
    from oracle import layout_oracle as L
    from surya_b200.config import AdetrConfig, LayoutConfig, SwinConfig, table_decoder
    from surya_b200.synth import adetr_table_state_dict, layout_synthetic_pages, swin_state_dict, table_query_tokens

    enc_cfg = SwinConfig(image_size=(256, 512), depths=(2, 2, 2, 2), encoder_length=128)
    cfg = LayoutConfig(encoder=enc_cfg, decoder=table_decoder(2))
    sde, sdd = swin_state_dict(enc_cfg, 3), adetr_table_state_dict(cfg.decoder, 3)
    enc, dec = ref_shim.build_reference_table_models(cfg, sde, sdd)
    x = layout_synthetic_pages(2, enc_cfg.image_size, seed=21)
    d = cfg.decoder
    q = table_query_tokens(d, 2)
    rng = np.random.default_rng(5)
    cols = []
    for _ in range(4):
        b = rng.integers(0, 1025, size=6).tolist()
        cols.append(b + [2 + d.special_token_count, d.special_token_count, 0, d.special_token_count])
    ids = torch.cat([q, torch.tensor(cols, dtype=torch.long).unsqueeze(0).repeat(2, 1, 1)], dim=1)     # [2, 7, 10]
    steps = 5
    with torch.inference_mode():
        ref_enc = enc(pixel_values=x).last_hidden_state
        dec.model._setup_cache(dec.config, 2, "cpu", torch.float32)
        pos = torch.ones_like(ids[0, :, 0], dtype=torch.int64).cumsum(0) - 1
        cur, toks, heads = ids, [], []
        for s in range(steps):
            out = dec(input_ids=cur, encoder_hidden_states=ref_enc, cache_position=pos, use_cache=True, prefill=(s == 0))
            pos = pos[-1:] + 1
            logits = out["box_property_logits"]
            tok, _ = L.table_next_tokens(logits, d)
            cur = tok.unsqueeze(1)
            toks.append(tok)
            heads.append({k: v[:, -1].float().clone() for k, v in logits.items()})
    g = {"encoder": ref_enc.float().clone(), "prompt": ids, "tokens": torch.stack(toks, 1),
         "heads": {k: torch.stack([h[k] for h in heads], 1) for k in heads[0]},
         "meta": {"kind": "table_nonsquare_cellpass", "steps": steps, "seed": 3, "page_seed": 21, "image_size": [256, 512],
                  "torch": str(torch.__version__), "reference": "VikParuchuri/surya@80e9a7e (v0.14.6), fp32 CPU"}}
    torch.save(g, GOLDEN / "table_nonsquare_cellpass.pt")
    print(f"[golden] table_nonsquare_cellpass: enc {tuple(ref_enc.shape)} tokens[0,0]={g['tokens'][0, 0].tolist()}")


def make_det_golden():
    """Reference EfficientViT segmentation logits for one seeded 512x512 page (surya/detection/__init__.py:94-104)."""
    from surya_b200.config import det_default
    from surya_b200.synth import det_normalize, det_state_dict, det_synthetic_pages

    cfg = det_default()
    sd = det_state_dict(cfg, seed=0)
    m = ref_shim.build_reference_det_model(cfg, sd)
    x = det_normalize(det_synthetic_pages(1, 512, seed=11, text_like=True))
    with torch.inference_mode():
        logits = m(pixel_values=x).logits.float()
        logits16 = m.half()(pixel_values=x.half()).logits.float()    # the reference's own fp16 path (settings MODEL_DTYPE on GPU)
    g = {"logits": logits, "logits_fp16_path": logits16, "input_checksum": x.double().sum(),
         "meta": {"reference": "VikParuchuri/surya@80e9a7e EfficientViTForSemanticSegmentation, CPU; fp32 and model.half() runs",
                  "size": 512, "seed": 11}}
    torch.save(g, GOLDEN / "det_default.pt")
    print(f"[golden] det_default: logits {tuple(logits.shape)} max={logits.max():.4f}")


def trace_crops():
    """Seeded crops of mixed widths for the predictor-trace golden: prompts of different lengths (left padding, merge offsets of
    both signs) through 3 batch rows."""
    widths = [512, 300, 700, 256, 900, 380, 620]
    return [rec_synthetic_crops(1, 48, w, seed=200 + i)[0] for i, w in enumerate(widths)]


def make_predictor_trace_golden():
    """The reference's UNMODIFIED RecognitionPredictor.prediction_loop (prefill / decode / merge / maybe_trim_cache_padding,
    surya/recognition/__init__.py:326-607) run over B200SuryaModel + SlotCache with the CPU oracle as the engine
    (oracle/ref_predictors.py); every model call and cache operation is logged so the GPU test can replay the exact sequence
    against the CUDA engine.  fp32, tiny config, 7 crops through 3 rows, max_tokens 10, trim threshold lowered to 2."""
    from oracle import ref_predictors as RP

    cfg = tiny_rec()
    sd = rec_state_dict(cfg, seed=0)
    t0 = time.time()
    events, tokens, bboxes, scores, eng = RP.record_rec_trace(cfg, sd, trace_crops(), batch_size=3, max_tokens=10,
                                                              min_trim_length=2)
    assert len(eng.free_slots) == eng.max_slots, "the predictor run leaked KV slots"
    g = {"events": events, "tokens": tokens, "scores": scores, "bboxes": bboxes,
         "meta": {"kind": "rec_predictor_trace", "batch_size": 3, "max_tokens": 10, "min_trim_length": 2, "seed": 0,
                  "torch": str(torch.__version__), "reference": "VikParuchuri/surya@80e9a7e RecognitionPredictor.prediction_loop, fp32 CPU"}}
    torch.save(g, GOLDEN / "rec_predictor_trace.pt")
    kinds = [e["kind"] for e in events]
    print(f"[golden] rec_predictor_trace: {time.time() - t0:.1f}s {len(events)} events "
          f"({kinds.count('prefill')} prefill, {kinds.count('decode')} decode, merges "
          f"{[(e['idxs'], e['offset']) for e in events if e['kind'] == 'merge']}, trims {[e['n'] for e in events if e['kind'] == 'trim']}) "
          f"tokens[0]={tokens[0]}")


def make_swin_window_padding_golden():
    """Oracle pin for DonutSwinLayer.maybe_pad (donut/encoder.py:591-596, crop :657-659): a 288x352 input gives token grids
    72x88 / 36x44 / 18x22 / 9x11 -- every merge sees even sides (the reference's floor-sized stage position tables demand it) but
    stages 1-3 are not multiples of the 8x8 window, so their layers run on zero-padded grids with the shift mask of the padded
    size.  Encoder output only (1 page)."""
    from surya_b200.config import LayoutConfig, SwinConfig, table_decoder
    from surya_b200.synth import adetr_table_state_dict, layout_synthetic_pages, swin_state_dict

    enc_cfg = SwinConfig(image_size=(288, 352), depths=(2, 2, 2, 2), encoder_length=99)
    cfg = LayoutConfig(encoder=enc_cfg, decoder=table_decoder(2))
    sde, sdd = swin_state_dict(enc_cfg, 5), adetr_table_state_dict(cfg.decoder, 5)
    enc, _ = ref_shim.build_reference_table_models(cfg, sde, sdd)
    x = layout_synthetic_pages(1, enc_cfg.image_size, seed=23)
    with torch.inference_mode():
        ref_enc = enc(pixel_values=x).last_hidden_state
    g = {"encoder": ref_enc.float().clone(), "input_checksum": x.double().sum(),
         "meta": {"kind": "swin_window_padding", "seed": 5, "page_seed": 23, "image_size": [288, 352], "torch": str(torch.__version__),
                  "reference": "VikParuchuri/surya@80e9a7e (v0.14.6), fp32 CPU"}}
    torch.save(g, GOLDEN / "swin_window_padding.pt")
    print(f"[golden] swin_window_padding: enc {tuple(ref_enc.shape)} absmax {ref_enc.abs().max():.3f}")


def make_ocr_error_golden():
    """Reference DistilBertForSequenceClassification (surya/ocr_error/model/encoder.py:697-766, eager attention, fp32 CPU) on seeded
    right-padded batches: logits, [CLS] hidden state and the predictor's argmax labels (surya/ocr_error/__init__.py:55-57)."""
    from surya_b200.config import ocr_error_default, ocr_error_tiny
    from surya_b200.synth import ocr_error_state_dict, ocr_error_synthetic_batch

    for kind, cfg, n, max_len, seed in (("tiny", ocr_error_tiny(), 12, 40, 3), ("default", ocr_error_default(), 16, 96, 3)):
        sd = ocr_error_state_dict(cfg, seed=0)
        m = ref_shim.build_reference_ocr_error_model(cfg, sd)
        ids, mask = ocr_error_synthetic_batch(cfg, n, max_len, seed=seed)
        with torch.inference_mode():
            logits = m(ids, attention_mask=mask).logits.float()
            hidden = m.distilbert(ids, attention_mask=mask)[0].float()
        g = {"input_ids": ids, "attention_mask": mask, "logits": logits, "cls_hidden": hidden[:, 0].clone(), "labels": logits.argmax(1),
             "meta": {"reference": "VikParuchuri/surya@80e9a7e DistilBertForSequenceClassification, eager attention, fp32 CPU",
                      "n": n, "max_len": max_len, "seed": seed, "kind": kind}}
        torch.save(g, GOLDEN / f"ocr_error_{kind}.pt")
        print(f"[golden] ocr_error_{kind}: logits {tuple(logits.shape)} labels={g['labels'].tolist()}")


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    which = set(sys.argv[1:]) or {"rec", "det", "layout", "table", "trace", "ocr_error"}
    if "layout" in which:
        make_layout_golden()
    if "table" in which:
        make_table_golden()
        make_layout_variants_golden()
        make_swin_window_padding_golden()
    if "det" in which:
        make_det_golden()
    if "trace" in which:
        make_predictor_trace_golden()
    if "ocr_error" in which:
        make_ocr_error_golden()
    if "rec" not in which:
        return
    for kind, cfg, steps in (("tiny", tiny_rec(), 32), ("synrec", syn_rec(), 40)):
        t0 = time.time()
        sd = rec_state_dict(cfg, seed=0)
        batch = O.build_batch(golden_crops(kind), cfg)
        lm, bb = run_reference(cfg, sd, batch, steps)
        g = summarise(lm, bb, cfg, full_logits=(kind == "tiny"))
        g["meta"] = {"kind": kind, "steps": steps, "seed": 0, "torch": str(torch.__version__),
                     "reference": "VikParuchuri/surya@80e9a7e (v0.14.6), fp32 CPU, attn=sdpa",
                     "input_ids_shape": list(batch["input_ids"].shape), "n_tiles": int(batch["image_tiles"].shape[0])}
        g["input_ids"] = batch["input_ids"]
        g["grid_thw"] = torch.from_numpy(batch["grid_thw"])
        g["tiles_checksum"] = batch["image_tiles"].double().sum()
        torch.save(g, GOLDEN / f"rec_{kind}.pt")
        print(f"[golden] {kind}: {time.time() - t0:.1f}s tokens={g['tokens'].tolist()} min margin={g['margin'].min():.4f}")


if __name__ == "__main__":
    main()
