set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "astar or alternated or golden or tick_by_tick or host_api" > gpurun_out/r2_astar2_gputest.log 2>&1; tail -4 gpurun_out/r2_astar2_gputest.log
FUZZ_LIB=abm timeout 300 python tools/fuzz.py astar 230000 600 2>&1 | tail -2 | tee gpurun_out/r2_astar2_fuzz.log
python bench.py --workload C4 --replans 16384 --ref-replans 8 > gpurun_out/r2_astar2_C4_16384.json 2> gpurun_out/r2_astar2_C4.err; tail -2 gpurun_out/r2_astar2_C4.err; cut -c1-700 gpurun_out/r2_astar2_C4_16384.json
