# ncu captures of one step on the 128^3 gyroid (K2 / K3 with --set full) -> distilled into profiles/r2_kernel_metrics.json
tag=${1:-x}
python tools/prof_step.py 128 8 2 > gpurun_out/r2_prof_step_128_$tag.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k[23]_(kernel|cell3)" --launch-skip 8 --launch-count 8 -o gpurun_out/r2_k2k3_128_$tag -f python tools/prof_step.py 128 8 2 > gpurun_out/r2_ncu_k2k3_$tag.log 2>&1
tail -2 gpurun_out/r2_prof_step_128_$tag.txt
