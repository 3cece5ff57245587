#!/bin/bash
# Measurement pass of round 2 (run through gpurun): the driver's bench line, the reference arm, the
# launch list of one bench step, full ncu captures of the kernels at the bench's launch shape, and
# of the two kernels that carry the long-form case (frame-tiled K1, cut-chain K3).
set -u
TAG=${1:-r2j}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 300 gpurun_out/bench_${TAG}.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 64 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/launches_${TAG}.stdout 2>&1
for K in gmm_topn senone_mix_active chain_viterbi; do
  ncu --set full --clock-control none --import-source on -k regex:${K} -s 1 -c 1 -f -o gpurun_out/prof_${K}_${TAG} \
      python bench.py --steps 1 --warmup 1 --utts 4096 --no-cpu-baseline --no-extra > gpurun_out/prof_${K}_${TAG}.stdout 2>&1
  echo "ncu ${K} rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:gmm_scan_ft -s 1 -c 1 -f -o gpurun_out/prof_gmm_scan_ft_longform_${TAG} \
    python tools/bench_longform.py > gpurun_out/prof_gmm_scan_ft_longform_${TAG}.stdout 2>&1
echo "ncu ft longform rc=$?"
ncu --set full --clock-control none --import-source on -k regex:chain_viterbi -s 1 -c 1 -f -o gpurun_out/prof_chain_viterbi_cut_${TAG} \
    python tools/bench_longform.py > gpurun_out/prof_chain_viterbi_cut_${TAG}.stdout 2>&1
echo "ncu K3 cut rc=$?"
ls -la gpurun_out | grep ${TAG}
