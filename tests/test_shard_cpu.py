"""World-size-2 gloo test of the sharding host logic (deal, weight broadcast, result all-gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_run(widths, max_tokens):
    def run(indices):
        toks, scs = [], []
        boxes = np.zeros((len(indices), max_tokens, 6), np.int64)
        for j, i in enumerate(indices):
            L = 1 + (i * 7) % max_tokens
            toks.append([(i * 31 + k) % 1000 for k in range(L)])
            scs.append([0.5 + 0.001 * k for k in range(L)])
            boxes[j, :L] = i
        return toks, scs, boxes
    return run


def _worker(rank, world, port, widths, max_tokens, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from surya_b200.shard import broadcast_tensors, sharded_recognition

    w = [torch.arange(12, dtype=torch.float32).reshape(3, 4), torch.ones(5, dtype=torch.bfloat16) * 3] if rank == 0 else None
    w = broadcast_tensors(w)
    ok_w = torch.equal(w[0], torch.arange(12, dtype=torch.float32).reshape(3, 4)) and w[1].dtype == torch.bfloat16 and float(w[1].sum()) == 15.0
    out = sharded_recognition(_fake_run(widths, max_tokens), widths, max_tokens)
    q.put((rank, ok_w, out[0], out[1], out[2].sum()))
    dist.destroy_process_group()


def test_deal_round_robin_balances_width_sorted_units():
    from surya_b200.shard import deal_round_robin

    widths = [100, 500, 300, 500, 50, 700, 20]
    shares = deal_round_robin(widths, 2)
    assert sorted(shares[0] + shares[1]) == list(range(7))
    assert shares[0][0] == 5 and shares[1][0] == 1            # widest first, ties by index
    assert abs(len(shares[0]) - len(shares[1])) <= 1


def test_sharded_gather_world2_matches_single_process():
    widths = [int(w) for w in np.random.default_rng(0).integers(20, 900, size=13)]
    max_tokens = 16
    from surya_b200.shard import sharded_recognition

    ref = sharded_recognition(_fake_run(widths, max_tokens), widths, max_tokens)     # world size 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, widths, max_tokens, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok_w, toks, scs, boxsum in res:
        assert ok_w, "weight broadcast mismatch"
        assert toks == ref[0]
        assert all(np.allclose(a, b) for a, b in zip(scs, ref[1]))
        assert boxsum == ref[2].sum()


def _pages_worker(rank, world, port, n_pages, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from surya_b200.shard import sharded_pages

    def run(lo, hi):
        return torch.arange(lo, hi, dtype=torch.float32)[:, None, None] * torch.ones(1, 2, 3)

    out = sharded_pages(run, n_pages)
    q.put((rank, out))
    dist.destroy_process_group()


def test_page_slices_and_sharded_pages_world2():
    from surya_b200.shard import page_slices, sharded_pages

    assert page_slices(32, 8) == [(4 * r, 4 * r + 4) for r in range(8)]
    assert page_slices(5, 2) == [(0, 3), (3, 5)] and page_slices(1, 2) == [(0, 1), (1, 1)]
    ref = sharded_pages(lambda lo, hi: torch.arange(lo, hi, dtype=torch.float32)[:, None, None] * torch.ones(1, 2, 3), 5)
    for n_pages in (5, 1):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_pages_worker, args=(r, 2, port, n_pages, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in range(2)]
        for p in procs:
            p.join(timeout=60)
        for _, out in res:
            assert out.shape == (n_pages, 2, 3)
            assert torch.equal(out, ref[:n_pages])


def _ocr_worker(rank, world, port, n_pages, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from surya_b200.pipeline import Line, sharded_ocr

    class FakePipe:
        """Stands in for OcrPipeline: page p (global id carried in the pixel data) has p % 3 lines with deterministic content."""

        def run(self, pages, fixed_steps=False):
            out = []
            for i, pg in enumerate(pages):
                gid = int(pg[0, 0, 0])
                out.append([Line(page=i, polygon=[[gid, j], [gid + 5, j], [gid + 5, j + 2], [gid, j + 2]], confidence=0.5 + 0.1 * j,
                                 tokens=[gid, j, 7][: 1 + (gid + j) % 3], scores=[0.25, 0.5, 0.75][: 1 + (gid + j) % 3])
                            for j in range(gid % 3)])
            return out, {"fake": 0.0}

    pages = np.zeros((n_pages, 4, 4, 3), dtype=np.uint8)
    pages[:, 0, 0, 0] = np.arange(n_pages)
    res, _ = sharded_ocr(FakePipe(), pages, max_tokens=4)
    q.put((rank, [[(ln.page, ln.polygon, round(ln.confidence, 4), ln.tokens, [round(s, 4) for s in ln.scores]) for ln in pg] for pg in res]))
    dist.destroy_process_group()


def test_sharded_ocr_world2():
    """BASELINE config 5 sharding: every rank runs the pipeline on its page share, lines of ALL pages come back on every rank in
    page order (gloo, world size 2; 7 pages so the shares are uneven and one page has no line)."""
    n_pages = 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ocr_worker, args=(r, 2, port, n_pages, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] == res[1][1]
    pages = res[0][1]
    assert len(pages) == n_pages
    for gid, pg in enumerate(pages):
        assert len(pg) == gid % 3
        for j, (page, poly, conf, toks, scs) in enumerate(pg):
            assert page == gid and poly[0] == [gid, j] and conf == round(0.5 + 0.1 * j, 4)
            assert toks == [gid, j, 7][: 1 + (gid + j) % 3] and len(scs) == len(toks)


def _gather_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from surya_b200.shard import gather_step_results

    T, B = 5, 4
    tok = torch.arange(T * B, dtype=torch.int64).reshape(T, B) + 1000 * rank
    score = torch.full((T, B), 0.25 * (rank + 1))
    bbox = (torch.arange(T * B * 6, dtype=torch.int64).reshape(T, B, 6) % 1025) + rank
    toks, scores, boxes = gather_step_results(tok, score, bbox)
    ok = all(torch.equal(toks[r].long(), torch.arange(T * B).reshape(T, B) + 1000 * r) for r in range(world))
    ok &= all(torch.equal(boxes[r].long(), (torch.arange(T * B * 6).reshape(T, B, 6) % 1025) + r) for r in range(world))
    ok &= all(float(scores[r][0, 0]) == 0.25 * (r + 1) for r in range(world)) and boxes[0].dtype == torch.int16
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gather_step_results_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res)


class _FakeErrModel:
    """Stands in for B200DistilBert on CPU: label = parity of the first real token after [CLS] (deterministic, text-local)."""
    from surya_b200.config import ocr_error_tiny as _cfg
    cfg = _cfg()

    def forward_packed(self, plan):
        first = torch.tensor([int(plan["ids"][s + 1]) % 2 if n > 1 else 0 for s, n in zip(plan["seq_start"], plan["seq_len"])])
        return torch.nn.functional.one_hot(first, 2).float()


def _err_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from surya_b200.ocr_error import sharded_error_labels
    from surya_b200.synth import ocr_error_synthetic_batch

    ids, mask = ocr_error_synthetic_batch(_FakeErrModel.cfg, 11, 24, seed=9)
    q.put((rank, sharded_error_labels(_FakeErrModel(), ids, mask, batch_size=4)))
    dist.destroy_process_group()


def test_sharded_error_labels_world2_matches_single_process():
    """ocr_error over 2 ranks (gloo): 11 texts split 6 / 5, labels all-gathered in text order == the single-process loop."""
    from surya_b200.ocr_error import detect_errors, sharded_error_labels
    from surya_b200.synth import ocr_error_synthetic_batch

    ids, mask = ocr_error_synthetic_batch(_FakeErrModel.cfg, 11, 24, seed=9)
    ref = detect_errors(_FakeErrModel(), ids, mask, batch_size=4)
    assert sharded_error_labels(_FakeErrModel(), ids, mask, batch_size=4) == ref and len(set(ref)) == 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_err_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, labels in res:
        assert labels == ref, f"rank {rank}"
