"""Micro-benchmark (GPU box): the small kernels of one recognition decode step, each as 48 back-to-back launches in a graph."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from surya_b200 import ops  # noqa: E402


def graph_time(fn, n=48, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1e3


def main():
    dt = torch.bfloat16
    B, H, V = 256, 1280, 65792
    x = torch.randn(B, H, device="cuda").to(dt)
    w = torch.randn(H, device="cuda").to(dt)
    y = torch.empty_like(x)
    print(f"rmsnorm 256x1280: {graph_time(lambda: ops.rmsnorm(x, w, 1e-6, out=y)):.2f} us")
    logits = torch.randn(B, V, device="cuda").to(dt)
    print(f"argmax_score 256x65792: {graph_time(lambda: ops.argmax_score(logits, 1, 0)):.2f} us")
    emb = torch.randn(V, H, device="cuda").to(dt)
    ids = torch.randint(0, V, (B,), device="cuda")
    print(f"embed_rows: {graph_time(lambda: ops.embed_rows(ids, emb)):.2f} us")
    hw = torch.randn(6, H, device="cuda").to(dt)
    hb = torch.randn(6, device="cuda").to(dt)
    print(f"small_head: {graph_time(lambda: ops.small_head(x, hw, hb, sigmoid=True, box_scale=1025.0)):.2f} us")


if __name__ == "__main__":
    main()
