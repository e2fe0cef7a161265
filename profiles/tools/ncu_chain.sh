#!/bin/bash
# usage: bash profiles/tools/ncu_chain.sh <kernel-regex> <out-name> [size]   (plain run first, as the recipe asks)
K=$1; O=$2; N=${3:-10801}
python profiles/tools/chain_profile.py $N 1 > gpurun_out/plain_$O.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 2 -f -o gpurun_out/$O python profiles/tools/chain_profile.py $N 1 > gpurun_out/ncu_$O.log 2>&1
tail -n 3 gpurun_out/ncu_$O.log
