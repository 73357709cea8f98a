# final measurement job of round 2 (one B200): [parity tests,] bench lines of every config, launch list, reparameterization
# Usage: gpurun -- 'bash tools/gpu_job_r2_final.sh [tag] [notests]'
tag=${1:-z}
if [ "$2" != "notests" ]; then python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_gputest_$tag.log; fi
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_C3_$tag.json 2> gpurun_out/r2_bench_C3_$tag.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_$tag.json 2> gpurun_out/r2_bench_reference_$tag.err
for c in C1 C2 C3b C4 C5; do
  python bench.py --config $c --steps 5 --warmup 3 --e2e-steps 3 > gpurun_out/r2_bench_${c}_$tag.json 2> gpurun_out/r2_bench_${c}_$tag.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_bench_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/r2_ncu_launch_$tag.log 2>&1
if [ "$2" != "notests" ]; then
python tools/bench_reparam.py --n 256 > gpurun_out/r2_reparam_256_$tag.jsonl 2> gpurun_out/r2_reparam_256_$tag.err
tail -c 300 gpurun_out/r2_gputest_$tag.log
fi
for c in C1 C2 C3 C3b C4 C5; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_${c}_$tag.json").read().strip().splitlines()[-1])
    print("$c", d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("$c", "failed", e)
PY
done
