set -x
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/r2_s4_C2.json 2> gpurun_out/r2_s4_C2.err; tail -c 700 gpurun_out/r2_s4_C2.json; tail -5 gpurun_out/r2_s4_C2.err
ncu --set full --import-source on --clock-control none -k regex:k_cta_lb --launch-skip 3 --launch-count 1 -o gpurun_out/r2_s4_C3_stream python bench.py --workload C3 --steps 3 --warmup 2 --no-cpu-baseline --e2e-steps 1 --no-sub > gpurun_out/ncu_s4.log 2>&1
tail -3 gpurun_out/ncu_s4.log
timeout 600 python -m pytest tests -m gpu -x -q -k "golden or tile_path or stream or packed" > gpurun_out/r2_gputest_4.log 2>&1; tail -3 gpurun_out/r2_gputest_4.log
