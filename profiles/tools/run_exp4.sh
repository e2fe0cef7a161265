timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bands.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/t_par.log 2>&1; tail -n 3 gpurun_out/t_par.log
timeout 60 python profiles/tools/chain_profile.py 10801 5 > gpurun_out/prof_a.json 2> gpurun_out/prof_a.err; tail -n 3 gpurun_out/prof_a.err
timeout 60 python profiles/tools/chain_profile.py 10801 3 c4 > gpurun_out/prof_c.json 2> gpurun_out/prof_c.err
python - <<'PY'
import json
for t in 'ac':
    try:
        d=json.loads(open(f'gpurun_out/prof_{t}.json').read().strip().splitlines()[-1])
        print(t, round(d['ms_per_step'],3), round(d['kernel_ms_sum'],3), {k:v['ms'] for k,v in list(d['kernels'].items())[:8]})
        print({k:d['debug'][k] for k in ('slow_sweeps','slow_visits','light_visits','light_cycles','light_cycles_load','light_cycles_fold','light_cycles_sweeps','lake_cycles')})
    except Exception as e: print(t, 'failed', e)
PY
