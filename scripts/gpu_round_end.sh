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
N=1000000 python scripts/trace_pool.py
N=1000000 ncu --set full --clock-control none --import-source on -k regex:'k_integrate_thread|k_free_flight|k_pool_resolve' \
    -s 40 -c 16 -o gpurun_out/prof_final_pool python scripts/trace_pool.py > gpurun_out/ncu_full_final.log 2>&1
tail -2 gpurun_out/ncu_full_final.log
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_traced.py tests/test_gpu_sweep.py -m gpu -x -q -k "golden or argument or cli" > gpurun_out/memcheck_final.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck_final.log
