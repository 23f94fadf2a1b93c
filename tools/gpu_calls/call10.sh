set -x
python tools/variant_probe.py --workload C2 --size 4096 animal-movement-abm_b200/libabm.so animal-movement-abm_b200/libabm_h8u768.so animal-movement-abm_b200/libabm_h6u768.so animal-movement-abm_b200/libabm_h5u640.so animal-movement-abm_b200/libabm_h4u640.so animal-movement-abm_b200/libabm_h4u1024.so animal-movement-abm_b200/libabm_c1.so 2>&1 | tee gpurun_out/r2_probe_halo.log
