#!/bin/bash
# Run on the GPU box (via gpurun): launch list of one bench step + full ncu captures of the
# two scoring kernels.  Outputs land in gpurun_out/ (scratch); summaries are copied into
# profiles/ by tools/summarise_profiles.py in the build container.
set -u
TAG=${1:-r1}
UTTS=${2:-1024}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 64 --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_${TAG}.stdout 2>&1
echo "launch list rc=$?"
for K in gmm_topn senone_mix_active chain_viterbi; do
  ncu --set full --clock-control none --import-source on -k regex:${K} -s 1 -c 1 -f \
      -o gpurun_out/prof_${K}_${TAG} \
      python bench.py --steps 1 --warmup 1 --utts ${UTTS} --no-cpu-baseline > gpurun_out/prof_${K}_${TAG}.stdout 2>&1
  echo "ncu ${K} rc=$?"
done
ls -la gpurun_out
# K4 in the reference's default mode (config #3)
ncu --set full --clock-control none --import-source on -k regex:fsg_search_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_fsg_search_active_${TAG} \
    python tools/bench_fsg.py --active --steps 1 --utts ${UTTS} > gpurun_out/prof_fsg_search_active_${TAG}.stdout 2>&1
echo "ncu fsg_search rc=$?"
ls -la gpurun_out
