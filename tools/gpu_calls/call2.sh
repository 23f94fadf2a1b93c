set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_1.log 2>&1; tail -5 gpurun_out/r2_gputest_1.log
python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_s1_C3.json 2> gpurun_out/r2_s1_C3.err; tail -c 1500 gpurun_out/r2_s1_C3.json; tail -3 gpurun_out/r2_s1_C3.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_s1_C2.json 2> gpurun_out/r2_s1_C2.err; tail -c 600 gpurun_out/r2_s1_C2.json; tail -3 gpurun_out/r2_s1_C2.err
