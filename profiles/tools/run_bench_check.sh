python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','cpu_baseline','gpu_launches','clocks')})"
python bench.py --workload bands --local-bands 2 --size 2000 --steps 3 > gpurun_out/bench_bands_emul.json 2> gpurun_out/bench_bands_emul.err; tail -c 600 gpurun_out/bench_bands_emul.err; python -c "
import json; d=json.load(open('gpurun_out/bench_bands_emul.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','roofline')}, d['config']['checks'])"
python bench.py --workload c1 --steps 1 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; tail -c 800 gpurun_out/bench_c1.err; cat gpurun_out/bench_c1.json
