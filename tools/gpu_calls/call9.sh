set -x
python tools/variant_probe.py --workload C3 --size 8192 animal-movement-abm_b200/libabm.so animal-movement-abm_b200/libabm_s64_32.so animal-movement-abm_b200/libabm_s96_21.so animal-movement-abm_b200/libabm_s32_32.so 2>&1 | tee gpurun_out/r2_probe_stream2.log
python tools/variant_probe.py --workload C2 --size 4096 animal-movement-abm_b200/libabm.so animal-movement-abm_b200/libabm_f6.so animal-movement-abm_b200/libabm_f8.so 2>&1 | tee gpurun_out/r2_probe_foreach.log
timeout 600 python -m pytest tests -m gpu -x -q -k "golden or stream" > gpurun_out/r2_gputest_7.log 2>&1; tail -3 gpurun_out/r2_gputest_7.log
