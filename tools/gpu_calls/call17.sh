set -x
python tools/variant_probe.py --workload C2 --size 4096 animal-movement-abm_b200/libabm_base.so animal-movement-abm_b200/libabm.so 2>&1 | tee gpurun_out/r2_probe_v4.log
python tools/phase_probe.py animal-movement-abm_b200/libabm_phase.so 2>&1 | tee gpurun_out/r2_phase_v4.log
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or tile_path or baseline_configs or tick_by_tick" > gpurun_out/r2_gputest_v4.log 2>&1; tail -3 gpurun_out/r2_gputest_v4.log
ncu --set full --import-source on --clock-control none -k regex:k_cta --launch-skip 12 --launch-count 1 -o gpurun_out/r2_v4_C2_resolve python bench.py --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 0 --no-sub > gpurun_out/ncu_v4.log 2>&1
