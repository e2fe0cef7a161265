#!/bin/bash
# round-2 closing evidence on one B200: GPU suite, bench lines (C2 default, C5, label, C1), ncu launch list and --set full
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_tests.log 2>&1; tail -n 3 gpurun_out/r02f_tests.log
timeout 200 python bench.py > gpurun_out/r02f_bench_c2.json 2> gpurun_out/r02f_bench_c2.err; tail -c 300 gpurun_out/r02f_bench_c2.err
timeout 120 python bench.py --workload c5 --steps 5 > gpurun_out/r02f_bench_c5.json 2> gpurun_out/r02f_bench_c5.err
timeout 120 python bench.py --workload label --steps 5 > gpurun_out/r02f_bench_label.json 2> gpurun_out/r02f_bench_label.err
timeout 120 python bench.py --workload c4 --steps 5 --no-e2e --cpu-sample 0 > gpurun_out/r02f_bench_c4.json 2> gpurun_out/r02f_bench_c4.err
timeout 120 python bench.py --workload c1 --steps 2 --warmup 1 > gpurun_out/r02f_bench_c1.json 2> gpurun_out/r02f_bench_c1.err
python - <<'PY'
import json
for t in ("c2", "c5", "label", "c4", "c1"):
    try:
        d = json.loads(open(f"gpurun_out/r02f_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, round(d["ms_per_step"], 3), round(d["value"], 1), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(t, "failed", e)
PY
python profiles/tools/one_chain.py 10801 2 > gpurun_out/plain_one_chain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -c 400 --csv \
    --log-file gpurun_out/r02f_launches.csv python profiles/tools/one_chain.py 10801 2 > gpurun_out/ncu_launches.log 2>&1
tail -n 2 gpurun_out/ncu_launches.log
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_ptr|k_edges|k_apply_d8|k_flat_fast_sparse|k_acc_tile_local|k_acc_tile_final|k_flat_init_lists|k_bor_persistent|k_flat_dirs_lists|k_flat_relax|k_perim|k_acc_link_walk" \
    -c 24 -f -o gpurun_out/r02f_full python profiles/tools/one_chain.py 10801 2 > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log
ls -la gpurun_out/r02f_full.ncu-rep
