"""RecognitionRunner's host logic on the CPU: the scheduler (prefill when > 20 % of the rows are free, slot bookkeeping, polling,
stop rules on the "device" state or in the per-token host loop) runs over oracle.ref_predictors.OracleRecEngine — the same surface
as the CUDA engine, computed by the CPU oracle — and must return, per crop, what the oracle's single-batch greedy loop returns.
(The CUDA engine under the same runner is tests/test_rec_gpu.py::test_runner_continuous_batching_matches_oracle.)"""
import numpy as np
import pytest
import torch

from oracle import rec_oracle as O
from oracle.ref_predictors import OracleRecEngine, stop_rules_step
from surya_b200.config import tiny_rec
from surya_b200.recognition import RecognitionRunner, detect_repeat_token
from surya_b200.synth import rec_state_dict, rec_synthetic_crops


@pytest.fixture(scope="module")
def setup():
    torch.set_num_threads(4)
    cfg = tiny_rec()
    sd = rec_state_dict(cfg, seed=0)
    crops = [rec_synthetic_crops(1, 48, 256 + 61 * i, seed=100 + i)[0] for i in range(7)]
    steps = 10
    ref = []
    for c in crops:
        otok, osc, obox, hist = O.greedy_decode(sd, cfg, O.build_batch([c], cfg), steps, torch.float32, stop_rules=True)
        ref.append((hist[0], osc[0].numpy(), obox[0].numpy()))
    return cfg, sd, crops, steps, ref


@pytest.mark.parametrize("mode,batch,poll", [("device", 3, 4), ("host", 3, 4), ("device", 4, 1), ("device", 8, 16)])
def test_runner_over_the_oracle_engine_matches_the_greedy_loop(setup, mode, batch, poll):
    cfg, sd, crops, steps, ref = setup
    eng = OracleRecEngine(cfg, sd, dtype=torch.float32, max_slots=batch + 1, s_max=256)
    runner = RecognitionRunner(eng, batch_size=batch, max_tokens=steps, poll=poll, stop_rules=mode)
    tokens, scores, bboxes = runner.run(crops)
    for i, (rtok, rsc, rbox) in enumerate(ref):
        assert tokens[i] == rtok, (mode, i)
        assert np.allclose(scores[i], rsc[: len(rtok)], atol=1e-5)
        assert np.array_equal(bboxes[i, : len(rtok)], rbox[: len(rtok)])
    assert len(eng.free_slots) == batch + 1 and not eng.caches                   # every slot came back
    assert getattr(eng, "sched", None) is None                                   # the runner never leaves its state bound
    n_prefills = sum(c.startswith("prefill") for c in eng.calls)
    assert n_prefills >= -(-len(crops) // batch)


def test_runner_repeat_rule_and_fixed_steps(setup):
    """A repeat window of 2 makes the repeat rule fire as soon as a token doubles; both rule implementations agree with the
    per-token host rules applied to the fixed-length run."""
    cfg, sd, crops, steps, ref = setup
    eng = OracleRecEngine(cfg, sd, dtype=torch.float32, max_slots=5, s_max=256)
    fixed = RecognitionRunner(eng, batch_size=4, max_tokens=steps, poll=3).run(crops[:4], fixed_steps=True)[0]
    assert all(len(t) == steps for t in fixed)
    expect = []
    for t in fixed:
        k = next((j + 1 for j in range(len(t)) if j > 0 and (t[j] in (cfg.eos_token_id, cfg.pad_token_id) or detect_repeat_token(t[: j + 1], 2))),
                 steps)
        expect.append(t[:k])
    for mode in ("device", "host"):
        r = RecognitionRunner(eng, batch_size=4, max_tokens=steps, poll=3, stop_rules=mode)
        r.MAX_REPEATS = 2
        got = r.run(crops[:4])[0]
        for g, e in zip(got, expect):
            if e[0] in (cfg.eos_token_id, cfg.no_output_token_id):
                assert g == e[:1]
            else:
                assert g == e, mode


def test_stop_rules_mirror_equals_host_rules():
    """The Python mirror of stop_rules_kernel used by OracleRecEngine against detect_repeat_token on crafted streams (the CUDA kernel
    is checked against the same host rules in tests/test_rec_gpu.py::test_stop_rules_kernel_matches_python_rules)."""
    rng = np.random.default_rng(3)
    for R, max_tokens in ((40, 100), (6, 1000), (2, 50)):
        T, B, EOS = 120, 5, 1
        toks = np.stack([rng.permutation(5000)[:T] + 10, np.full(T, 77), np.tile([5, 6, 7], T)[:T], np.tile([1, 2, 3, 4, 5, 6], T)[:T] + 20,
                         rng.integers(10, 14, T)], 1).astype(np.int64)
        toks[33, 0] = EOS
        first = np.array([9, 77, 7, 26, 11])
        gen = torch.ones(B, dtype=torch.int32)
        ring = torch.zeros((B, R), dtype=torch.int64)
        ring[:, 0] = torch.from_numpy(first)
        done, valid, active = torch.zeros(B, dtype=torch.uint8), torch.zeros(B, dtype=torch.int32), torch.zeros(1, dtype=torch.int32)
        th, dh = torch.from_numpy(toks), torch.from_numpy((toks == EOS).astype(np.uint8))
        hist = [[int(f)] for f in first]
        alive = [True] * B
        for s in range(T):
            stop_rules_step(th, dh, s, gen, ring, done, valid, active, max_tokens, R)
            for r in range(B):
                if alive[r]:
                    hist[r].append(int(toks[s, r]))
                    if hist[r][-1] == EOS or len(hist[r]) >= max_tokens or detect_repeat_token(hist[r], R):
                        alive[r] = False
            assert done.tolist() == [0 if a else 1 for a in alive], (R, s)
            assert gen.tolist() == [len(h) for h in hist]
            assert int(active) == sum(alive)
