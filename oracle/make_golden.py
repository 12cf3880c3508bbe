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


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    for kind, cfg, steps in (("tiny", tiny_rec(), 12), ("synrec", syn_rec(), 3)):
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
