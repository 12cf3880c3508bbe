"""Micro-benchmark (GPU box): decode-sized GEMMs, plain tile configs vs split-K cluster configs; back-to-back launches of the
same shape over rotating weight copies (so weights come from HBM like in a real decode step), CUDA-event timed."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from surya_b200 import ops  # noqa: E402


def timeit(fn, n=48, reps=5):
    """n back-to-back launches captured in ONE CUDA graph (no host launch cost), replayed `reps` times."""
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
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
    shapes = [("qkv", 256, 1920, 1280, False), ("o", 256, 1280, 1280, False), ("gate_up", 256, 6848, 1280, True),
              ("down", 256, 1280, 3424, False), ("lm_head", 256, 65792, 1280, False), ("layout_q", 16, 1024, 1024, False),
              ("layout_gu", 16, 8192, 1024, True)]
    if len(sys.argv) > 1:
        shapes = [s for s in shapes if s[0] in sys.argv[1:]]
    for name, M, N, K, swi in shapes:
        copies = max(3, int(300e6 // (N * K * 2)))          # > L2
        ws = [torch.randn(N, K, device="cuda").to(dt) * 0.05 for _ in range(min(copies, 24))]
        a = torch.randn(M, K, device="cuda").to(dt)
        out = torch.empty(M, N // 2 if swi else N, device="cuda", dtype=dt)
        res = []
        cfgs = [("bn32", 32), ("bn64", 64), ("bn96", 96), ("bn128", 128), ("bn256", 256)]
        for pk in (1, 2, 3, 4):
            for bn in (64, 128):
                cfgs.append((f"sk{pk}x{bn}", 1000 * pk + bn))
        for label, force in cfgs:
            try:
                us = timeit(lambda i: ops.gemm(a, ws[i % len(ws)], act="silu" if swi else "none", swiglu=swi, out=out, force_bn=force))
                res.append((us, label))
            except Exception as e:      # config not launchable for this shape
                res.append((float("inf"), f"{label}:ERR"))
        res.sort()
        print(f"{name:10s} M={M} N={N} K={K}: " + "  ".join(f"{l}={u:.1f}" for u, l in res[:8]), flush=True)
        print(f"{'':10s} worst: " + "  ".join(f"{l}={u:.1f}" for u, l in res[-5:]), flush=True)


if __name__ == "__main__":
    main()
