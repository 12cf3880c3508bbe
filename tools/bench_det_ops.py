"""Micro-benchmark (GPU box): the CUDA-core kernels of the detection path at BASELINE config-3 sizes (32 pages of 1024x1024),
CUDA-event timed over rotating buffers (> L2), with the bytes each MUST move and the resulting GB/s."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from surya_b200 import ops  # noqa: E402

DT = torch.float16


def timeit(fn, reps=10):
    for _ in range(2):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    B = 32
    rows = []

    def rec(name, us, nbytes):
        print(f"{name:44s} {us:9.1f} us   {nbytes / 1e6:8.1f} MB   {nbytes / us / 1e3:7.0f} GB/s", flush=True)
        rows.append((name, us, nbytes))

    for name, H, C, ks in (("dw3x3 s1  stage 2 (64x64x1024)", 64, 1024, 3), ("dw3x3 s1  stage 3 (32x32x3072)", 32, 3072, 3),
                           ("dw5x5 s1  LiteMLA (32x32x1536)", 32, 1536, 5)):
        xs = [torch.randn(B, H, H, C, device="cuda").to(DT) for _ in range(3)]
        w = (torch.randn(ks * ks, C, device="cuda") / ks).to(DT)
        b = torch.randn(C, device="cuda") * 0.1
        us = timeit(lambda i: ops.dwconv_nhwc(xs[i % 3], w, b if ks == 3 else None, ks, 1, ks // 2, "hardswish" if ks == 3 else "none"))
        rec(name, us, 2 * xs[0].numel() * 2)
    for name, H, C in (("dw3x3 s2  stage 2 (128x128x2048)", 128, 2048), ("dw3x3 s2  stage 3 (64x64x6144)", 64, 6144)):
        xs = [torch.randn(B, H, H, C, device="cuda").to(DT) for _ in range(2)]
        w = (torch.randn(9, C, device="cuda") / 3).to(DT)
        b = torch.randn(C, device="cuda") * 0.1
        us = timeit(lambda i: ops.dwconv_nhwc(xs[i % 2], w, b, 3, 2, 1, "hardswish"), reps=5)
        rec(name, us, int(xs[0].numel() * 2 * 1.25))
    for name, H, cin, cout, stride, res in (("conv3x3 32->32 s1 stem block (512x512)", 512, 32, 32, 1, True),
                                          ("conv3x3 32->512 s2 stage 0 (512->256)", 512, 32, 512, 2, False),
                                          ("conv3x3 64->256 s1 stage 0 (256x256)", 256, 64, 256, 1, False)):
        xs = [torch.randn(B, H, H, cin, device="cuda").to(DT) for _ in range(2)]
        w = (torch.randn(cout, 9 * cin, device="cuda") * (9 * cin) ** -0.5).to(DT)
        b = torch.randn(cout, device="cuda") * 0.1
        Ho = H // stride
        r = torch.randn(B, Ho, Ho, cout, device="cuda").to(DT) if res else None
        us = timeit(lambda i: ops.conv2d_nhwc(xs[i % 2], w, b, r, 3, stride, 1, "none" if res else "hardswish"), reps=5)
        rec(name, us, (xs[0].numel() + B * Ho * Ho * cout * (2 if res else 1)) * 2)
    heads, dim, HW = 16, 32, 1024
    qa = [torch.randn(B * HW, 3 * heads * dim, device="cuda").to(DT) for _ in range(3)]
    qb = [torch.randn(B * HW, 3 * heads * dim, device="cuda").to(DT) for _ in range(3)]
    us = timeit(lambda i: ops.lite_mla(qa[i % 3], qb[i % 3], B, HW, heads, dim, 1e-5))
    rec("lite_mla (32x32, 2x16 heads x 32)", us, (2 * qa[0].numel() + B * HW * 2 * heads * dim) * 2)
    print("\n| kernel | us | MB that must move | GB/s |\n|---|---:|---:|---:|")
    for n, u, nb in rows:
        print(f"| {n} | {u:.1f} | {nb / 1e6:.1f} | {nb / u / 1e3:.0f} |")


if __name__ == "__main__":
    main()
