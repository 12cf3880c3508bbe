"""Recognition crop preprocessing on the device (GPU): sb_rec_preprocess (Lanczos4 scale_to_fit, cubic resize to x28,
normalise, tile) against the numpy restatement (oracle/preproc_oracle.py, pinned to cv2 in test_preproc_cpu.py) and against the
product's cv2 host path; then the recognition runner fed by either path.

Tolerance: floating point.  5e-4 on the 0..255 pixel scale for one resize (float32 summation-order noise of a 64-tap sum; the kernel
keeps the restatement's tap order, so it usually sits near 1e-5), twice that through both stages; in normalised units that is
x 1 / (255 * 0.224).  bf16 / fp16 tiles round at 4e-3 / 5e-4 relative, three orders above this."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import preproc_oracle as P

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
TOL_255 = 5e-4
NORM = 1.0 / 255 / 0.224


def _report(name, payload):
    OUT.mkdir(exist_ok=True)
    path = OUT / "preproc_parity.json"
    data = json.loads(path.read_text()) if path.exists() else {}
    data[name] = payload
    path.write_text(json.dumps(data, indent=1))


def _crops():
    rng = np.random.default_rng(7)
    sizes = [(48, 512), (40, 300), (64, 900), (300, 2000), (20, 60), (168, 168), (56, 560), (250, 1100), (33, 47), (28, 1008)]
    crops = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    yy, xx = np.mgrid[0:60, 0:700]
    smooth = (127 + 100 * np.sin(xx / 9.0) * np.cos(yy / 5.0))[..., None].repeat(3, 2)
    crops.append(smooth.astype(np.uint8))                                  # text-like low-frequency content
    return crops


def _tiny_engine(dtype=torch.float16):
    from surya_b200.config import tiny_rec
    from surya_b200.recognition import RecEngine
    from surya_b200.synth import rec_state_dict

    cfg = tiny_rec()
    return cfg, RecEngine(cfg, rec_state_dict(cfg, seed=0), dtype=dtype, max_slots=16, s_max=512, max_patches=16384, max_tokens=4096)


def test_device_preprocess_matches_restatement_and_cv2(built_lib):
    from surya_b200.recognition import RecognitionRunner, scale_to_fit, tile_image

    cfg, eng = _tiny_engine()
    runner = RecognitionRunner(eng, batch_size=4, max_tokens=4)
    crops = _crops()
    tiles, grids, seqs = runner.preprocess_device(crops)
    torch.cuda.synchronize()
    tiles = tiles.cpu().numpy()
    h_tiles, h_grids, h_seqs = runner.preprocess(crops)                    # the cv2 host path on the same crops
    assert [tuple(g) for g in grids] == [tuple(g) for g in h_grids]
    assert all(np.array_equal(a, b) for a, b in zip(seqs, h_seqs))
    row, worst_o, worst_c = 0, 0.0, 0.0
    for i, crop in enumerate(crops):
        o_tiles, o_grid = P.process_crop(crop, cfg.vision_encoder.patch_size, cfg.merge_size)
        n = o_tiles.shape[0]
        assert tuple(o_grid) == tuple(grids[i])
        mine = tiles[row:row + n]
        eo, ec = np.abs(mine - o_tiles).max(), np.abs(mine - h_tiles[i]).max()
        worst_o, worst_c = max(worst_o, eo), max(worst_c, ec)
        assert eo <= 2 * TOL_255 * NORM, f"crop {i} {crop.shape}: {eo / NORM:.2e} (0..255 scale) off the restatement"
        assert ec <= 2 * TOL_255 * NORM, f"crop {i} {crop.shape}: {ec / NORM:.2e} (0..255 scale) off the cv2 host path"
        row += n
    assert row == tiles.shape[0]
    _report("tiles", {"max_err_vs_restatement_255": float(worst_o / NORM), "max_err_vs_cv2_255": float(worst_c / NORM), "crops": len(crops)})
    eng.close()


def test_runner_on_device_preprocessing_equals_host_preprocessing(built_lib):
    """Same crops through preprocess='device' and preprocess='host': identical grids / prompts, prefill logits equal to 16-bit rounding
    noise, tokens equal except on near-ties."""
    from surya_b200.recognition import RecognitionRunner
    from surya_b200.synth import rec_synthetic_crops

    cfg, eng = _tiny_engine()
    crops = [rec_synthetic_crops(1, 48, 256 + 41 * i, seed=300 + i)[0] for i in range(9)] + _crops()[3:6]
    runner = RecognitionRunner(eng, batch_size=6, max_tokens=12, poll=4)
    td, sd, bd = runner.run(crops, preprocess="device")
    th, sh, bh = runner.run(crops, preprocess="host")
    same = sum(int(a == b) for a, b in zip(td, th))
    first = sum(int(a[0] == b[0]) for a, b in zip(td, th))
    _report("runner", {"crops": len(crops), "identical_token_lists": same, "identical_first_tokens": first})
    assert first >= len(crops) - 1 and same >= len(crops) - 2, (td, th)
    with pytest.raises(Exception):
        runner.run([np.zeros((48, 100, 3), np.float32)], preprocess="device")      # float crops belong to the host path
    assert runner.run([], preprocess="device")[0] == []
    eng.close()


def test_device_preprocess_at_bench_size(built_lib):
    """BASELINE config-2 input (256 crops of 48 x 512): every crop goes through both resizes; spot-check 3 crops against the
    restatement and all tiles for finiteness / range."""
    from surya_b200.recognition import RecognitionRunner
    from surya_b200.synth import rec_synthetic_crops

    cfg, eng = _tiny_engine()
    runner = RecognitionRunner(eng, batch_size=4, max_tokens=4)
    crops = list(rec_synthetic_crops(256, 48, 512, seed=1234))
    tiles, grids, seqs = runner.preprocess_device(crops)
    torch.cuda.synchronize()
    assert tiles.shape == (256 * 160, 588) and all(tuple(g) == (1, 4, 40) for g in grids)
    t = tiles.cpu().numpy()
    assert np.isfinite(t).all() and t.min() > -6.0 and t.max() < 6.0       # Lanczos / cubic overshoot on noise is not clipped (float images)
    for i in (0, 100, 255):
        o_tiles, _ = P.process_crop(crops[i])
        assert np.abs(t[i * 160:(i + 1) * 160] - o_tiles).max() <= 2 * TOL_255 * NORM
    eng.close()
