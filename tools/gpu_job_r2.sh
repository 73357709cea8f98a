# round-2 measurement job (one B200): parity tests, bench lines of every config, ncu captures.
# Usage: gpurun -- 'bash tools/gpu_job_r2.sh [tag]'
tag=${1:-f}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_gputest_$tag.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_C3_$tag.json 2> gpurun_out/r2_bench_C3_$tag.err
for c in C1 C2 C3b C4 C5; do
  python bench.py --config $c --steps 5 --warmup 3 --e2e-steps 3 > gpurun_out/r2_bench_${c}_$tag.json 2> gpurun_out/r2_bench_${c}_$tag.err
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_$tag.json 2> gpurun_out/r2_bench_reference_$tag.err
python tools/prof_step.py 128 8 2 > gpurun_out/r2_prof_step_128_$tag.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k[23]_(kernel|cell3)" --launch-skip 8 --launch-count 8 -o gpurun_out/r2_k2k3_128_$tag -f python tools/prof_step.py 128 8 2 > gpurun_out/r2_ncu_k2k3_$tag.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_bench_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/r2_ncu_launch_$tag.log 2>&1
tail -c 300 gpurun_out/r2_gputest_$tag.log
