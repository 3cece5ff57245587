#!/bin/bash
# launch list + full ncu capture of the frame-tiled K1 at the bench's launch shape
set -u
TAG=${1:-r2a}
export SSB_K1=${SSB_K1:-ft}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_${TAG}.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:gmm_scan_ft -s 1 -c 1 -f -o gpurun_out/prof_gmm_scan_ft_${TAG} \
    python bench.py --steps 1 --warmup 1 --utts 4096 --no-cpu-baseline > gpurun_out/prof_gmm_scan_ft_${TAG}.stdout 2>&1
echo "ncu rc=$?"
ls -la gpurun_out | grep ${TAG}
