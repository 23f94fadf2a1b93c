set -x
FUZZ_LIB=abm timeout 900 python tools/fuzz.py engine 200000 400 2>&1 | tail -3 | tee gpurun_out/r2_gpu_fuzz.log
FUZZ_LIB=abm FUZZ_SCHED=1 ABM_TILE_ROUNDS=2 timeout 600 python tools/fuzz.py engine 210000 200 2>&1 | tail -3 | tee -a gpurun_out/r2_gpu_fuzz.log
FUZZ_LIB=abm timeout 600 python tools/fuzz.py astar 220000 300 2>&1 | tail -3 | tee -a gpurun_out/r2_gpu_fuzz.log
