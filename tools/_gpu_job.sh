timeout 900 python -m pytest tests/test_det_gpu.py -x -q -s 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02s_det_tests.log
(SB_FMB_FUSED=1 timeout 120 python tools/bench_det_forward.py; SB_FMB_FUSED=0 timeout 300 python tools/bench_det_forward.py; SB_FMB_FUSED=0 SB_DET_FUSED_HEAD=0 timeout 300 python tools/bench_det_forward.py; SB_FMB_FUSED=0 SB_DET_FUSED_HEAD=0 SB_CONV_HALO=0 timeout 300 python tools/bench_det_forward.py) > gpurun_out/r02s_det_forward.txt 2>&1
timeout 200 python tools/bench_det_ops.py 2>&1 | tail -12 > gpurun_out/r02s_det_ops.txt
cat gpurun_out/r02s_det_tests.log gpurun_out/r02s_det_forward.txt gpurun_out/r02s_det_ops.txt
