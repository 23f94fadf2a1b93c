set -x
ncu --set full --import-source on --clock-control none -k regex:k_cta --launch-skip 12 --launch-count 4 -o gpurun_out/r2_base_C2_full python bench.py --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_c2.log 2>&1
ncu --set full --import-source on --clock-control none -k 'regex:k_cta|k_foreach' --launch-skip 6 --launch-count 4 -o gpurun_out/r2_base_C3_full python bench.py --workload C3 --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_c3.log 2>&1
ls -la gpurun_out/*.ncu-rep
