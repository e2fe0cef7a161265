timeout 600 python -m pytest tests/test_gpu_stencils.py tests/test_gpu_veg.py tests/test_gpu_water.py -m gpu -x -q > gpurun_out/t_st.log 2>&1; tail -n 5 gpurun_out/t_st.log
timeout 600 python bench.py --workload c5 > gpurun_out/bench_c5.json 2> gpurun_out/c5.err; tail -c 400 gpurun_out/c5.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_c5.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d.get("kernels"))
PY
