python profiles/tools/chain_profile.py 10801 3 > gpurun_out/prof_c2.json 2> gpurun_out/prof_c2.err; python -c "
import json; d=json.load(open('gpurun_out/prof_c2.json')); print(d['ms_per_step'], d['debug'])"; tail -n 3 gpurun_out/prof_c2.err
python profiles/tools/chain_profile.py 10801 3 c4 > gpurun_out/prof_c4.json 2> gpurun_out/prof_c4.err; cat gpurun_out/prof_c4.json; tail -n 3 gpurun_out/prof_c4.err
bash profiles/tools/ncu_chain.sh "k_flat_fast_sparse|k_flat_init|k_flat_relax|k_apply_d8|k_flat_dirs" r02_flats
