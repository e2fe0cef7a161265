timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bands.py -m gpu -x -q > gpurun_out/t_par.log 2>&1; tail -n 3 gpurun_out/t_par.log
HC_NO_ACC_OVERLAP=1 timeout 300 python profiles/tools/chain_profile.py 10801 2 > gpurun_out/prof_c2_noov.json 2> gpurun_out/prof_c2.err; cat gpurun_out/prof_c2_noov.json | cut -c1-800; tail -n 3 gpurun_out/prof_c2.err
HC_NO_ACC_OVERLAP=1 timeout 300 python profiles/tools/chain_profile.py 10801 2 c4 > gpurun_out/prof_c4_noov.json 2> gpurun_out/prof_c4.err; cat gpurun_out/prof_c4_noov.json | cut -c1-800; tail -n 3 gpurun_out/prof_c4.err
