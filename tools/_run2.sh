python -m pytest tests/test_gpu_multi.py tests/test_gpu_launcher.py -x -q 2>&1 | tail -8
B200RT_REFERENCE_ROOT=$PWD/.ab/reference python -m pytest tests/test_reference_main_e2e.py -x -q 2>&1 | tail -8
mkdir -p gpurun_out/e2e && python tools/run_reference_main.py .ab/reference gpurun_out/e2e "Cornell box" 512 100 2>&1 | tail -12 > gpurun_out/r2_reference_main.log; tail -3 gpurun_out/r2_reference_main.log
