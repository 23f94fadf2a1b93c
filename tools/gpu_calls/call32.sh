D=animal-movement-abm_b200
python tools/variant_probe.py --workload C2 --size 4096 --ticks 30 $D/libabm_prev.so $D/libabm.so $D/libabm_prev.so $D/libabm.so 2>&1 | tee gpurun_out/r2_probe_pairadd.log
