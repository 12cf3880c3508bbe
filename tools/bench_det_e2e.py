"""Detection host-to-host variants (GPU box): detect_text_front_host with uint8 / fp16 pages and different chunk sizes, and the
resident forward at the chunk's batch size (the pipeline cannot beat sum(forward(chunk)))."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main(B=32, S=1024):
    from surya_b200.config import det_default
    from surya_b200.detection import DetEngine, detect_text_front_host
    from surya_b200.synth import det_normalize, det_state_dict, det_synthetic_pages

    cfg = det_default()
    eng = DetEngine(cfg, det_state_dict(cfg, 0), torch.float16, max_batch=B, max_hw=(S, S))
    pages = det_synthetic_pages(B, S, seed=1234)
    u8 = torch.from_numpy(pages).contiguous().pin_memory()
    x16 = det_normalize(pages).half().pin_memory()
    xd = x16.cuda()

    def timeit(fn, reps=4):
        fn(); fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    for c in (4, 8, 16, 32):
        ms = timeit(lambda: [eng.forward(xd[i:i + c]) for i in range(0, B, c)])
        print(f"resident forward in chunks of {c:2d}: {ms:7.2f} ms  {B / ms * 1e3:7.1f} pages/s")
    for name, src in (("uint8", u8), ("fp16", x16)):
        for c in (4, 8, 16, 32):
            ms = timeit(lambda: detect_text_front_host(eng, src, chunk=c))
            print(f"front half host->host, {name} pages, chunk {c:2d}: {ms:7.2f} ms  {B / ms * 1e3:7.1f} pages/s")


if __name__ == "__main__":
    main()
