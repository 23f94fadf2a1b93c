# NOT to be repeated: `ncu --set full` on the A* kernel (seconds per pass, ~40 instrumented passes) ran into the call's time limit,
# wedged the box and cost the round's last GPU minutes; no report came back.  The metrics pass of call48.sh is the A* evidence.
set -x
ncu --set full --import-source on --clock-control none -k k_cta --launch-skip 1 --launch-count 1 -o gpurun_out/r2_astar_v5_full python bench.py --workload C4 --replans 4736 --ref-replans 2 > gpurun_out/r2_astar_v5_full.log 2>&1
tail -2 gpurun_out/r2_astar_v5_full.log | cut -c1-300
ls -la gpurun_out/r2_astar_v5_full.ncu-rep
