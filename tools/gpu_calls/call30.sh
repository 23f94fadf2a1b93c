set -x
timeout 1700 python bench.py --workload C4 --replans 1048576 --ref-replans 32 > gpurun_out/r2_astar_C4_1048576replans.json 2> gpurun_out/r2_astar_C4.err; tail -3 gpurun_out/r2_astar_C4.err; cut -c1-900 gpurun_out/r2_astar_C4_1048576replans.json
