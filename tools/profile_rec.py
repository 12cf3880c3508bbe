"""Profiling driver (run under ncu on the GPU box): one recognition prefill + a few eager decode steps at BASELINE
config-2 sizes (256 crops, SYN-REC, bf16).  `--det` profiles one detection forward (B=4, 1024x1024, fp16) instead."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def rec(n_decode: int = 3, batch: int = 256):
    from surya_b200.config import syn_rec
    from surya_b200.recognition import RecEngine, RecognitionRunner, build_prefill_plan
    from surya_b200.synth import rec_state_dict, rec_synthetic_crops

    cfg = syn_rec()
    eng = RecEngine(cfg, rec_state_dict(cfg, 0), dtype=torch.bfloat16, max_slots=batch + 4, s_max=256,
                    max_patches=batch * 160, max_tokens=batch * 46)
    runner = RecognitionRunner(eng, batch_size=batch, max_tokens=128)
    tiles, grids, seqs = runner.preprocess(list(rec_synthetic_crops(batch, 48, 512, seed=1234)))
    slots = eng.alloc_slots(batch)
    plan = build_prefill_plan(cfg, np.array(grids), seqs, slots)
    tiles_dev = torch.from_numpy(np.concatenate(tiles, 0)).cuda()
    slot_t = torch.tensor(slots, dtype=torch.int32, device="cuda")
    pos = torch.tensor([len(s) for s in seqs], dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("prefill")
    out = eng.prefill(tiles_dev, plan)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    ids = out["next_ids"]
    for _ in range(n_decode):
        o = eng.decode(ids, slot_t, pos)
        ids, pos = o["next_ids"], pos + 1
    torch.cuda.synchronize()


def det(batch: int = 4, size: int = 1024):
    from surya_b200.config import det_default
    from surya_b200.detection import DetEngine
    from surya_b200.synth import det_normalize, det_state_dict, det_synthetic_pages

    cfg = det_default()
    eng = DetEngine(cfg, det_state_dict(cfg, 0), torch.float16, max_batch=batch, max_hw=(size, size))
    x = det_normalize(det_synthetic_pages(batch, size, seed=3, text_like=True)).cuda().half()
    eng.forward(x)
    torch.cuda.synchronize()
    eng.forward(x)
    torch.cuda.synchronize()


def layout(batch: int = 16):
    from surya_b200.config import layout_default
    from surya_b200.layout import LayoutEngine, layout_greedy
    from surya_b200.synth import adetr_layout_state_dict, layout_synthetic_pages, swin_state_dict

    cfg = layout_default()
    eng = LayoutEngine(cfg, swin_state_dict(cfg.encoder, 0), adetr_layout_state_dict(cfg.decoder, 0), dtype=torch.float16,
                       max_batch=batch)
    x = layout_synthetic_pages(batch, cfg.encoder.image_size, seed=3).half().cuda()
    layout_greedy(eng, x, 3, use_graph=False)      # encoder + cross K/V + 3 eager decode steps
    torch.cuda.synchronize()


def new_paths():
    """Round-2 additions: ocr_error forward (64 texts right-padded to 512, default DistilBERT, fp16) and the device preprocessing of
    256 crops of 48 x 512 (sb_rec_preprocess), each run twice (the second run is the one to read)."""
    from surya_b200.config import ocr_error_default, tiny_rec
    from surya_b200.ocr_error import B200DistilBert, build_pack_plan
    from surya_b200.recognition import RecEngine, RecognitionRunner
    from surya_b200.synth import ocr_error_state_dict, ocr_error_synthetic_batch, rec_state_dict, rec_synthetic_crops

    cfg = ocr_error_default()
    model = B200DistilBert(cfg, ocr_error_state_dict(cfg, 0), dtype=torch.float16)
    ids, mask = ocr_error_synthetic_batch(cfg, 64, 512, seed=21, min_len=16)
    plan = build_pack_plan(ids.numpy(), mask.numpy(), cfg)
    print("ocr_error tokens", plan["n_tok"])
    for _ in range(2):
        model.forward_packed(plan)
        torch.cuda.synchronize()
    tcfg = tiny_rec()
    eng = RecEngine(tcfg, rec_state_dict(tcfg, 0), dtype=torch.float16, max_slots=8, s_max=256, max_patches=2048, max_tokens=512)
    runner = RecognitionRunner(eng, batch_size=4, max_tokens=4)
    crops = list(rec_synthetic_crops(256, 48, 512, seed=1234))
    for _ in range(2):
        runner.preprocess_device(crops)
        torch.cuda.synchronize()


if __name__ == "__main__":
    if "--new" in sys.argv:
        new_paths()
        sys.exit(0)
    if "--layout" in sys.argv:
        layout()
        sys.exit(0)
    if "--det" in sys.argv:
        b = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else 4
        det(batch=b)
    else:
        rec()
