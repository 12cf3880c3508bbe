"""GPU box: in-kernel timeline of gemm_chain (CTA 0) for one decoder layer's chain at B = 256, inside a warm sequence, plus the
in-graph time of the chain vs the four separate launches."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from surya_b200 import ops  # noqa: E402

DT = torch.bfloat16
B, D, Q, IP = 256, 1280, 1920, 3424
EPS = 1e-6


def timeit(fn, n=24, reps=5):
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
    rn = lambda *s, sc=1.0: (torch.randn(*s, device="cuda") * sc).to(DT)
    NW = 8
    wo = [rn(D, D, sc=0.03) for _ in range(NW)]
    wg = [rn(2 * IP, D, sc=0.03) for _ in range(NW)]
    wd = [rn(D, IP, sc=0.02) for _ in range(NW)]
    wq = [rn(Q, D, sc=0.03) for _ in range(NW)]
    bq = torch.randn(Q, device="cuda") * 0.1
    ao, x = rn(B, D), rn(B, D, sc=0.1)
    act = torch.empty(B, IP, device="cuda", dtype=DT)
    qkv = torch.empty(B, Q, device="cuda", dtype=DT)
    bar = torch.zeros(4, dtype=torch.int32, device="cuda")

    def phases(i):
        return [dict(a=ao, w=wo[i % NW], out=x, residual=x), dict(a=x, w=wg[i % NW], out=act, act="silu", swiglu=True, rms_eps=EPS),
                dict(a=act, w=wd[i % NW], out=x, residual=x), dict(a=x, w=wq[i % NW], out=qkv, bias=bq, rms_eps=EPS)]

    def chain(i):
        ops.gemm_chain(phases(i), bar)

    def separate(i):
        ops.gemm(ao, wo[i % NW], residual=x, out=x)
        ops.gemm(x, wg[i % NW], act="silu", swiglu=True, out=act, rms_eps=EPS)
        ops.gemm(act, wd[i % NW], residual=x, out=x, splitk=True)
        ops.gemm(x, wq[i % NW], bias=bq, out=qkv, rms_eps=EPS)

    print(f"chain (4 phases, one launch): {timeit(chain):.2f} us   separate launches: {timeit(separate):.2f} us", flush=True)
    for n in (1, 2, 3):
        print(f"chain of the first {n} phase(s): {timeit(lambda i: ops.gemm_chain(phases(i)[:n], bar)):.2f} us", flush=True)
    tl = torch.zeros(64, dtype=torch.int64, device="cuda")
    for i in range(6):
        tl.zero_()
        ops.gemm_chain(phases(i), bar, timeline=tl)
    torch.cuda.synchronize()
    t = tl.cpu().tolist()
    t0 = t[0]
    names = ["start", "first k-block", "last MMA", "acc ready", "epilogue done", "barrier passed"]
    print("timeline of CTA 0 (us since phase 0 start):")
    for pi, nm in enumerate(["o_proj bn%d" % ops.gemm_chain_bn(B, D), "gate_up bn%d" % ops.gemm_chain_bn(B, 2 * IP, True),
                             "down bn%d" % ops.gemm_chain_bn(B, D), "qkv bn%d" % ops.gemm_chain_bn(B, Q)]):
        row = [(t[8 * pi + k] - t0) / 1e3 if t[8 * pi + k] else float("nan") for k in range(6)]
        print(f"  {nm:16s} " + "  ".join(f"{n}={v:7.2f}" for n, v in zip(names, row)))


if __name__ == "__main__":
    main()
