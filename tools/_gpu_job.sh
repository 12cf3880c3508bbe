timeout 600 python -m pytest tests/test_det_gpu.py -x -q -s -k "fused_kernels or engine_vs or golden or batch32" 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r02u_det_tests.log
(timeout 300 python tools/bench_det_forward.py; SB_FMB_FUSED=2 timeout 300 python tools/bench_det_forward.py; SB_DET_FUSED_HEAD=0 timeout 300 python tools/bench_det_forward.py) > gpurun_out/r02u_det_forward.txt 2>&1
cat gpurun_out/r02u_det_tests.log gpurun_out/r02u_det_forward.txt
