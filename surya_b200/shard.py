"""Multi-GPU sharding of independent units (line crops / pages): one process per GPU, identical replicas.

The reference has no distributed code; units are independent, so the only exchange steps are the
one-off weight broadcast and the per-call all-gather of fixed-shape result tensors.  Works with any
torch.distributed backend: NCCL on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def deal_round_robin(widths: Sequence[int], n_ranks: int) -> List[List[int]]:
    """Sort units by width (descending, stable — RecognitionPredictor sorts crops by width before batching,
    surya/recognition/__init__.py:848-854) and deal them round-robin so every rank gets a similar length mix."""
    order = sorted(range(len(widths)), key=lambda i: (-widths[i], i))
    return [order[r::n_ranks] for r in range(n_ranks)]


def _collective_device(device=None) -> torch.device:
    """Device the collectives of the current process group run on: the caller's choice, else the current CUDA device
    under NCCL and the CPU under gloo (a rank must never fall back to 'cpu' just because its own share is empty)."""
    if device is not None:
        return torch.device(device)
    if dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_tensors(tensors: List[torch.Tensor] | None, meta_src: int = 0, device=None) -> List[torch.Tensor]:
    """Rank `meta_src` owns the packed weights; every other rank receives shapes then data (one broadcast per tensor)."""
    rank, n = world()
    if n == 1:
        return tensors
    meta = [[(tuple(t.shape), str(t.dtype).split(".")[-1]) for t in tensors]] if rank == meta_src else [None]
    dist.broadcast_object_list(meta, src=meta_src)
    if rank != meta_src:
        tensors = [torch.empty(s, dtype=getattr(torch, d), device=_collective_device(device)) for s, d in meta[0]]
    for t in tensors:
        dist.broadcast(t, src=meta_src)
    return tensors


def sharded_recognition(run_local: Callable[[List[int]], Tuple[List[List[int]], List[List[float]], np.ndarray]],
                        widths: Sequence[int], max_tokens: int, device=None):
    """Run `run_local(indices)` on this rank's share and all-gather (tokens, scores, bboxes) for ALL units.

    Results are exchanged as fixed-shape tensors: tokens int32 [n, max_tokens] (-1 padded), lengths int32 [n],
    scores fp32 [n, max_tokens], boxes int32 [n, max_tokens, 6]; shares are padded to the largest share."""
    rank, n_ranks = world()
    shares = deal_round_robin(widths, n_ranks)
    mine = shares[rank]
    toks, scs, boxes = run_local(mine) if mine else ([], [], np.zeros((0, max_tokens, 6), np.int64))
    cap = max(len(s) for s in shares)
    t_tok = torch.full((cap, max_tokens), -1, dtype=torch.int32)
    t_len = torch.zeros(cap, dtype=torch.int32)
    t_sc = torch.zeros((cap, max_tokens), dtype=torch.float32)
    t_box = torch.zeros((cap, max_tokens, 6), dtype=torch.int32)
    for j, (tk, sc) in enumerate(zip(toks, scs)):
        L = min(len(tk), max_tokens)
        t_len[j] = L
        t_tok[j, :L] = torch.tensor(tk[:L], dtype=torch.int32)
        t_sc[j, :L] = torch.tensor(sc[:L], dtype=torch.float32)
    if len(mine):
        t_box[: len(mine)] = torch.from_numpy(np.asarray(boxes)[:, :max_tokens].astype(np.int32))
    parts = [t_tok, t_len, t_sc, t_box]
    if device is not None:
        parts = [p.to(device) for p in parts]
    gathered = []
    for p in parts:
        if n_ranks > 1:
            outs = [torch.empty_like(p) for _ in range(n_ranks)]
            dist.all_gather(outs, p)
        else:
            outs = [p]
        gathered.append([o.cpu() for o in outs])
    n = len(widths)
    tokens: List[List[int]] = [[] for _ in range(n)]
    scores: List[List[float]] = [[] for _ in range(n)]
    bboxes = np.zeros((n, max_tokens, 6), dtype=np.int64)
    for r, share in enumerate(shares):
        for j, idx in enumerate(share):
            L = int(gathered[1][r][j])
            tokens[idx] = gathered[0][r][j, :L].tolist()
            scores[idx] = gathered[2][r][j, :L].tolist()
            bboxes[idx] = gathered[3][r][j].numpy()
    return tokens, scores, bboxes


def page_slices(n_pages: int, n_ranks: int) -> List[Tuple[int, int]]:
    """Contiguous page ranges per rank (detection / layout / table_rec: pages are equal-cost units),
    sizes differing by at most one."""
    base, extra = divmod(n_pages, n_ranks)
    out, lo = [], 0
    for r in range(n_ranks):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def sharded_pages(run_local: Callable[[int, int], torch.Tensor], n_pages: int, device=None, result_meta=None) -> torch.Tensor:
    """Run `run_local(lo, hi)` -> tensor [hi - lo, ...] on this rank's page range and all-gather the per-page results of ALL
    ranks in page order (detection heatmaps [n, 2, H/4, W/4], layout / table token histories [n, steps, cols], ...).
    Shares are padded to the largest share so the collective has a fixed shape.  result_meta = (per-page shape, dtype) skips
    the (host-side, pickled) metadata exchange when the caller knows the result layout — the steady-state serving path."""
    rank, n_ranks = world()
    if n_pages <= 0:
        return torch.zeros((0,), device=_collective_device(device))
    slices = page_slices(n_pages, n_ranks)
    lo, hi = slices[rank]
    mine = run_local(lo, hi) if hi > lo else None
    if n_ranks == 1:
        return mine
    # every rank needs the trailing shape / dtype even if its share is empty
    if result_meta is not None:
        shape, dtype = tuple(result_meta[0]), result_meta[1]
    else:
        meta = [None] * n_ranks
        dist.all_gather_object(meta, None if mine is None else (tuple(mine.shape[1:]), str(mine.dtype).split(".")[-1]))
        shape, dname = next(m for m in meta if m is not None)
        dtype = getattr(torch, dname)
    cap = max(h - l for l, h in slices)
    # every rank must hand the collective a tensor on the backend's device type, also the ranks whose share is empty
    dev = _collective_device(device)
    buf = torch.zeros((cap,) + tuple(shape), dtype=dtype, device=dev)
    if mine is not None:
        buf[: hi - lo] = mine.to(dev)
    outs = [torch.empty_like(buf) for _ in range(n_ranks)]
    dist.all_gather(outs, buf)
    return torch.cat([o[: h - l] for o, (l, h) in zip(outs, slices)], 0)


def gather_step_results(tok: torch.Tensor, score: torch.Tensor, bbox: torch.Tensor):
    """All-gather of one recognition step's fixed-shape DEVICE results over the replicas: token ids as int32
    [T, B], scores fp32 [T, B], boxes int16 [T, B, 6] (boxes are < 1025).  Returns per-rank lists in rank order; a no-op list of
    the local tensors for world size 1.  ~2.6 KB per crop on the wire."""
    rank, n = world()
    box16 = bbox.to(torch.int16).contiguous()
    if n == 1:
        return [tok.to(torch.int32)], [score.to(torch.float32)], [box16]
    # NCCL has no 16-bit integer type: the six int16 coordinates of a box travel as three int32 words
    parts = (tok.to(torch.int32), score.to(torch.float32), box16.view(torch.int32))
    out = []
    for p in parts:
        p = p.contiguous()
        bufs = [torch.empty_like(p) for _ in range(n)]
        dist.all_gather(bufs, p)
        out.append(bufs)
    out[2] = [b.view(torch.int16) for b in out[2]]
    return tuple(out)
