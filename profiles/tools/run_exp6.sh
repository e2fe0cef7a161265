timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 60 python profiles/tools/chain_profile.py 10801 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B3', round(d['ms_per_step'],3), d['kernels']['k_flat_relax'], d['debug']['slow_sweeps'])"
sed -i 's/#define FL_BATCH 3 /#define FL_BATCH 2 /' pydrodem_b200/csrc/flats.cu
make -C pydrodem_b200/csrc > /dev/null 2>&1
timeout 60 python profiles/tools/chain_profile.py 10801 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B2', round(d['ms_per_step'],3), d['kernels']['k_flat_relax'], d['debug']['slow_sweeps'])"
timeout 60 python profiles/tools/chain_profile.py 10801 3 c4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B2 c4', round(d['ms_per_step'],3), d['kernels']['k_flat_relax'])"
timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bands.py -m gpu -x -q 2>&1 | tail -2
