timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_full.log 2>&1; tail -n 5 gpurun_out/t_full.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["passes"], d["roofline"]["kernel"], d["roofline"]["frac"])
PY
