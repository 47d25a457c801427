#!/bin/bash
# Developer tool: ncu (full set) of the streaming edge kernels at the proteins shape -> gpurun_out/stream_raw.csv
set -e
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_edge_proj|k_edge_mlp' -c 8 -f -o gpurun_out/prof_stream \
  python tools/stream_probe.py > gpurun_out/ncu_stream.log 2>&1
ncu -i gpurun_out/prof_stream.ncu-rep --page raw --csv > gpurun_out/stream_raw.csv
