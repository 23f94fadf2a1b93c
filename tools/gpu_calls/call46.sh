set -x
nvidia-smi -L
timeout 400 python -m pytest tests -m gpu -x -q -k "astar or alternated or route or host_api or golden" > gpurun_out/r2_astar_v4_gputest.log 2>&1; tail -4 gpurun_out/r2_astar_v4_gputest.log
FUZZ_LIB=abm timeout 300 python tools/fuzz.py astar 240000 600 2>&1 | tail -2 | tee gpurun_out/r2_astar_v4_fuzz.log
timeout 200 python bench.py --workload C4 --replans 32768 --ref-replans 4 > gpurun_out/r2_astar_v4_C4_32768replans.json 2> gpurun_out/r2_astar_v4_C4.err; tail -2 gpurun_out/r2_astar_v4_C4.err; cut -c1-560 gpurun_out/r2_astar_v4_C4_32768replans.json
timeout 200 python bench.py --workload C4 --replans 65536 --ref-replans 4 > gpurun_out/r2_astar_v4_C4_65536replans.json 2>> gpurun_out/r2_astar_v4_C4.err; cut -c1-560 gpurun_out/r2_astar_v4_C4_65536replans.json
timeout 200 python bench.py --workload V6 --steps 5 --warmup 2 --no-sub --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2_astar_v4_bench_V6.json 2> gpurun_out/r2_astar_v4_bench_V6.err; tail -2 gpurun_out/r2_astar_v4_bench_V6.err; cut -c1-300 gpurun_out/r2_astar_v4_bench_V6.json
