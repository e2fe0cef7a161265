timeout 600 python -m pytest tests/test_gpu_mono_rivers.py tests/test_gpu_water.py -m gpu -x -q > gpurun_out/t_mono.log 2>&1; tail -n 25 gpurun_out/t_mono.log
