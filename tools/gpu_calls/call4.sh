set -x
python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/r2_s3_C3.json 2> gpurun_out/r2_s3_C3.err; tail -c 700 gpurun_out/r2_s3_C3.json; tail -3 gpurun_out/r2_s3_C3.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_s3_C2.json 2> gpurun_out/r2_s3_C2.err; tail -c 300 gpurun_out/r2_s3_C2.json; tail -5 gpurun_out/r2_s3_C2.err
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_3.log 2>&1; tail -5 gpurun_out/r2_gputest_3.log
ncu --set full --import-source on --clock-control none -k regex:k_cta --launch-skip 12 --launch-count 3 -o gpurun_out/r2_s3_C2_full python bench.py --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 1 --no-sub > gpurun_out/ncu_s3.log 2>&1
tail -3 gpurun_out/ncu_s3.log
