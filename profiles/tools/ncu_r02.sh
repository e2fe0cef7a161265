#!/bin/bash
# round-2 ncu evidence: (1) launch list with per-launch time and DRAM bytes of two chains (the second one is warm),
# (2) --set full of the full-grid kernels of the second chain.  Plain run first, as the recipe asks.
python profiles/tools/one_chain.py 10801 2 > gpurun_out/plain_one_chain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -c 400 --csv \
    --log-file gpurun_out/r02_launches.csv python profiles/tools/one_chain.py 10801 2 > gpurun_out/ncu_launches.log 2>&1
tail -n 2 gpurun_out/ncu_launches.log
python profiles/tools/one_chain.py 10801 2 > gpurun_out/plain_one_chain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:k_ptr|k_edges|k_apply_d8|k_flat_fast_sparse|k_flat_list|k_acc_tile_local|k_acc_tile_final|k_flat_init|k_bor_persistent|k_flat_dirs" \
    -f -o gpurun_out/r02_full python profiles/tools/one_chain.py 10801 2 > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log
