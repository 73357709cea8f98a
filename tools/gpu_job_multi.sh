# round-2 multi-GPU job: bash tools/gpu_job_multi.sh N   (under gpurun --gpus N)
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@"; }
run --steps 10 --warmup 3 --e2e-steps 2 > gpurun_out/r2_bench_C3_weak_n$N.json 2> gpurun_out/r2_bench_C3_weak_n$N.err
run --steps 10 --warmup 3 --e2e-steps 2 --scaling strong > gpurun_out/r2_bench_C3_strong_n$N.json 2> gpurun_out/r2_bench_C3_strong_n$N.err
if [ "$N" = "8" ]; then
  run --config C5 --steps 5 --warmup 3 --e2e-steps 0 --scaling strong > gpurun_out/r2_bench_C5_strong_n$N.json 2> gpurun_out/r2_bench_C5_strong_n$N.err
fi
tail -c 300 gpurun_out/r2_bench_C3_strong_n$N.err
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*_n$N.json')):
    try:
        d=json.load(open(f)); print(f, round(d['ms_per_step'],2), round(d['value']), d['e2e']['ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
