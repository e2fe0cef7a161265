timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bands.py -m gpu -x -q > gpurun_out/t_par.log 2>&1; tail -n 3 gpurun_out/t_par.log
HC_TRACE=1 timeout 120 python bench.py --workload bands --local-bands 2 --steps 2 --warmup 1 --no-e2e 2> gpurun_out/bands_trace.err > gpurun_out/bands_trace.json; grep "flats" gpurun_out/bands_trace.err | cut -c1-300 | tail -4; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bands_trace.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], json.dumps(d["config"].get("phase_ms_rank0")), d["config"].get("checks"))
PY
