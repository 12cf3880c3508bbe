timeout 300 python tools/bench_det_e2e.py 2>&1 | tail -14 > gpurun_out/r02y_det_e2e.txt; cat gpurun_out/r02y_det_e2e.txt
