set -x
python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/r2_s2_C3.json 2> gpurun_out/r2_s2_C3.err; tail -c 900 gpurun_out/r2_s2_C3.json; tail -3 gpurun_out/r2_s2_C3.err
ncu --set full --import-source on --clock-control none -k regex:k_cta_lb --launch-skip 3 --launch-count 1 -o gpurun_out/r2_s2_C3_stream python bench.py --workload C3 --steps 3 --warmup 2 --no-cpu-baseline --e2e-steps 1 --no-sub > gpurun_out/ncu_s2.log 2>&1
tail -3 gpurun_out/ncu_s2.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_2.log 2>&1; tail -5 gpurun_out/r2_gputest_2.log
