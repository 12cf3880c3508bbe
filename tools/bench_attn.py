"""Micro-benchmark (GPU box): decode attention at the BASELINE config-2 shape (B = 256 rows, 16/4 heads, d = 80) for several
cache lengths; 12 launches over 12 different per-layer caches captured in one CUDA graph (KV streams from HBM)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from surya_b200 import ops  # noqa: E402


def main():
    dt = torch.bfloat16
    B, nh, nkv, hd, s_max, L = 256, 16, 4, 80, 256, 12
    kc = [torch.randn(B, nkv, s_max, hd, device="cuda").to(dt) for _ in range(L)]
    vc = [torch.randn(B, nkv, s_max, hd, device="cuda").to(dt) for _ in range(L)]
    qkv = torch.randn(B, (nh + 2 * nkv) * hd, device="cuda").to(dt)
    slot = torch.arange(B, dtype=torch.int32, device="cuda")
    inv = (1.0 / (10000.0 ** (torch.arange(0, hd, 2).float() / hd))).cuda()
    out = torch.empty(B, nh * hd, device="cuda", dtype=dt)
    for S in (46, 80, 110, 140, 173):
        pos = torch.full((B,), S, dtype=torch.int32, device="cuda")
        for l in range(L):
            ops.decode_attn(qkv, kc[l], vc[l], slot, pos, inv, nh, nkv, hd, hd ** -0.5, out=out)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for l in range(L):
                ops.decode_attn(qkv, kc[l], vc[l], slot, pos, inv, nh, nkv, hd, hd ** -0.5, out=out)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / (10 * L) * 1e3
        mb = B * nkv * (S + 1) * hd * 2 * 2 / 1e6
        print(f"S={S:4d}: {us:6.1f} us/launch, KV {mb:5.1f} MB -> {mb / us * 1e-3 * 1e3:6.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
