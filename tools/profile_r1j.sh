#!/bin/bash
# Final measurement pass of round 1 (run through gpurun): bench lines of every config, the
# launch list of one bench step, full ncu captures of the kernels at the bench's launch shape.
set -u
TAG=r1j
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 300 gpurun_out/bench_${TAG}.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>/dev/null
python tools/bench_fsg.py > gpurun_out/bench_fsg_dense_${TAG}.json 2>/dev/null
python tools/bench_fsg.py --active > gpurun_out/bench_fsg_active_${TAG}.json 2>/dev/null
python tools/bench_longform.py 2>/dev/null | tail -1 > gpurun_out/bench_longform_${TAG}.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 64 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_${TAG}.stdout 2>&1
for K in gmm_topn senone_mix_active chain_viterbi backtrace; do
  ncu --set full --clock-control none --import-source on -k regex:${K} -s 1 -c 1 -f -o gpurun_out/prof_${K}_${TAG} \
      python bench.py --steps 1 --warmup 1 --utts 4096 --no-cpu-baseline > gpurun_out/prof_${K}_${TAG}.stdout 2>&1
  echo "ncu ${K} rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:fsg_search_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_fsg_search_active_${TAG} \
    python tools/bench_fsg.py --active --steps 1 > gpurun_out/prof_fsg_search_active_${TAG}.stdout 2>&1
echo "ncu fsg_search rc=$?"
ls gpurun_out | grep ${TAG}
