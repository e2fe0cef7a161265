set -x
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bands.py -m gpu -x -q > gpurun_out/t_par.log 2>&1; tail -n 3 gpurun_out/t_par.log
python profiles/tools/chain_profile.py 10801 5 > gpurun_out/prof_a.json 2> gpurun_out/prof_a.err; tail -n 3 gpurun_out/prof_a.err
HC_NO_ACC_OVERLAP=1 python profiles/tools/chain_profile.py 10801 5 > gpurun_out/prof_b.json 2> gpurun_out/prof_b.err
HC_FLAT_CTAS_PER_SM=2 python profiles/tools/chain_profile.py 10801 5 > gpurun_out/prof_c.json 2> gpurun_out/prof_c.err
HC_FLAT_CTAS_PER_SM=1 python profiles/tools/chain_profile.py 10801 5 > gpurun_out/prof_d.json 2> gpurun_out/prof_d.err
python - <<'PY'
import json
for t in 'abcd':
    try:
        d=json.loads(open(f'gpurun_out/prof_{t}.json').read().strip().splitlines()[-1])
        print(t, round(d['ms_per_step'],3), round(d['kernel_ms_sum'],3), {k:v['ms'] for k,v in list(d['kernels'].items())[:9]})
        if t=='a': print(d['debug'])
    except Exception as e: print(t, 'failed', e)
PY
