#!/bin/bash
# builds and runs the TMA probe against the in-tree libhydrocore.so (hc_tmap_2d lives there)
cd "$(dirname "$0")/../.."
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/tma_probe tests/probes/tma_probe.cu -Lpydrodem_b200 -lhydrocore -Xlinker -rpath=$PWD/pydrodem_b200 || exit 1
for args in "64 32 0 0" "64 64 0 0" "68 66 0 0" "68 66 8 8" "64 32 -1 -1" "68 66 -1 -1" "68 66 100 70" "80 80 -8 -8" "32 8 4 4"; do /tmp/tma_probe $args; done
