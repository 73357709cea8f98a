set -x
python -m pytest tests/test_gpu_round2.py -x -q -k "coeffs" 2>&1 | tail -8 > gpurun_out/r2_gputest_d.log
for c in C1 C2 C3b C4 C5; do
  python bench.py --config $c --steps 5 --warmup 3 --e2e-steps 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err
done
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_C3.json 2> gpurun_out/r2_bench_C3.err
python tools/prof_step.py 128 8 2 > gpurun_out/r2_prof_step_128.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k[23]_(kernel|cell3)" --launch-skip 8 --launch-count 8 -o gpurun_out/r2_k2k3_128 -f python tools/prof_step.py 128 8 2 > gpurun_out/r2_ncu_k2k3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/r2_ncu_launch.log 2>&1
tail -c 300 gpurun_out/r2_gputest_d.log
ls -la gpurun_out | tail -20
