python tools/prof_slab.py 256 64 32 > gpurun_out/r2_prof_slab.txt 2>&1
QG_TRACE=1 python tools/prof_slab.py 32 >> gpurun_out/r2_prof_slab.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_slab32.csv python tools/prof_slab.py 32 > /dev/null 2>&1
