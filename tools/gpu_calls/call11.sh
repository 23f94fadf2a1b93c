set -x
python tools/variant_probe.py --workload C2 --size 4096 animal-movement-abm_b200/libabm.so animal-movement-abm_b200/libabm_h6u768.so animal-movement-abm_b200/libabm_h7u768.so animal-movement-abm_b200/libabm_a512.so animal-movement-abm_b200/libabm_a384.so 2>&1 | tee gpurun_out/r2_probe_tail.log
ABM_TILE_ROUNDS=12 python tools/variant_probe.py --workload C2 --size 4096 animal-movement-abm_b200/libabm.so 2>&1 | tee -a gpurun_out/r2_probe_tail.log
timeout 600 python -m pytest tests -m gpu -x -q -k "golden or tile_path or baseline_configs" > gpurun_out/r2_gputest_8.log 2>&1; tail -3 gpurun_out/r2_gputest_8.log
