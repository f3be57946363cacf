#!/bin/bash
# One gpurun call: GPU tests, the default bench (both arms), the ncu launch list of a reduced bench
# command, a --set full capture of the hot kernels inside a pool run, and a memcheck of the small paths.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --impl reference > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 300 gpurun_out/bench_final.err
SMALL="--steps 2 --warmup 1 --batch 131072 --max-time 1e6 --no-cpu --exact-step 0 --relaxed-step 0 --e2e-steps 1 --warmup-e2e 0"
python bench.py $SMALL > gpurun_out/bench_small_final.json 2> /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py $SMALL > gpurun_out/ncu_list_final.log 2>&1
# full-occupancy capture of the hot kernels inside a pool run; the report stays on the box (tens of MB), its raw page comes back
N=2000000 CAP=40000 ncu --set full --clock-control none -k regex:'k_integrate_thread|k_free_flight|k_pool_resolve' \
    -s 17 -c 18 -o /tmp/prof_final_pool2 python scripts/pool_once.py > gpurun_out/ncu_full_final2.log 2>&1
ncu -i /tmp/prof_final_pool2.ncu-rep --page raw --csv > gpurun_out/prof_final_pool2_raw.csv
tail -2 gpurun_out/ncu_full_final2.log
