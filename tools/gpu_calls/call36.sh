set -x
ncu --set full --import-source on --clock-control none -k k_cta --launch-skip 1 --launch-count 1 -o gpurun_out/r2_astar_full python bench.py --workload C4 --replans 2048 --ref-replans 2 > gpurun_out/r2_astar_full.log 2>&1
tail -2 gpurun_out/r2_astar_full.log | cut -c1-300
