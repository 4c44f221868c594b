#!/bin/bash
# ncu full capture of one workload's kernels through tools/ab_probe.py (direct launches in timing mode are what ncu sees best)
#     gpurun --timeout 600 -- 'bash tools/ncu_probe.sh r2_vis truck_4k_dof "k_fragments|k_spans|k_clear"'
tag=${1:-probe}; wl=${2:-truck_4k_dof}; kern=${3:-"k_fragments|k_spans|k_clear|k_dof"}
out=gpurun_out; mkdir -p $out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$kern" -s ${5:-30} -c ${6:-10} -f -o $out/${tag}_prof \
    python tools/ab_probe.py --workloads $wl --frames 6 --modes ${4:-fast} > $out/${tag}_ncu.log 2>&1
ls -la $out/${tag}_prof.ncu-rep
