set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err; tail -c 400 gpurun_out/bench_r1i.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1i.json 2>/dev/null
python tools/bench_fsg.py > gpurun_out/bench_fsg_dense_r1i.json 2>/dev/null
python tools/bench_fsg.py --active > gpurun_out/bench_fsg_active_r1i.json 2>/dev/null
python tools/bench_longform.py 2>/dev/null | tail -1 > gpurun_out/bench_longform_r1i.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 64 --csv --log-file gpurun_out/launches_r1i.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_r1i.stdout 2>&1
for K in senone_mix_active backtrace; do
  ncu --set full --clock-control none --import-source on -k regex:${K} -s 1 -c 1 -f -o gpurun_out/prof_${K}_r1i python bench.py --steps 1 --warmup 1 --utts 4096 --no-cpu-baseline > gpurun_out/prof_${K}_r1i.stdout 2>&1
  echo "ncu ${K} rc=$?"
done
ls gpurun_out | grep r1i
