timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_tests.log 2>&1; tail -n 3 gpurun_out/r02f_tests.log
timeout 200 python bench.py > gpurun_out/r02f_bench_c2.json 2> gpurun_out/r02f_bench_c2.err; tail -c 300 gpurun_out/r02f_bench_c2.err
timeout 120 python bench.py --workload c4 --steps 5 --no-e2e --cpu-sample 0 > gpurun_out/r02f_bench_c4.json 2> gpurun_out/r02f_bench_c4.err
python - <<'PY'
import json
for t in ("c2", "c4"):
    try:
        d = json.loads(open(f"gpurun_out/r02f_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, round(d["ms_per_step"], 3), round(d["value"], 1), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"))
        print({k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
    except Exception as e:
        print(t, "failed", e)
PY
