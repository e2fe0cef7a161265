python -m pytest tests/test_gpu_parity.py tests/test_gpu_bands.py tests/test_gpu_host_stream.py -m gpu -x -q > gpurun_out/t_par.log 2>&1; tail -n 5 gpurun_out/t_par.log
HC_TRACE=1 python profiles/tools/chain_profile.py 10801 2 > gpurun_out/prof_c2_trace.json 2> gpurun_out/prof_c2_trace.err; tail -n 4 gpurun_out/prof_c2_trace.err
python profiles/tools/chain_profile.py 10801 5 > gpurun_out/prof_c2.json 2> gpurun_out/prof_c2.err; cat gpurun_out/prof_c2.json; tail -n 3 gpurun_out/prof_c2.err
HC_PITCHED=0 python profiles/tools/chain_profile.py 10801 5 > gpurun_out/prof_c2_dense.json 2> gpurun_out/prof_c2_dense.err; cat gpurun_out/prof_c2_dense.json; tail -n 3 gpurun_out/prof_c2_dense.err
