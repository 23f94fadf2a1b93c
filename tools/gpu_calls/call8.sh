set -x
python tools/variant_probe.py --workload C3 --size 8192 animal-movement-abm_b200/libabm.so animal-movement-abm_b200/libabm_s128_12.so animal-movement-abm_b200/libabm_s128_16.so animal-movement-abm_b200/libabm_s256_6.so animal-movement-abm_b200/libabm_s256_8.so animal-movement-abm_b200/libabm_s64_24.so 2>&1 | tee gpurun_out/r2_probe_stream.log
timeout 600 python -m pytest tests -m gpu -x -q -k "group or golden" > gpurun_out/r2_gputest_6.log 2>&1; tail -3 gpurun_out/r2_gputest_6.log
