D=animal-movement-abm_b200
python tools/variant_probe.py --workload C3 --size 8192 --ticks 20 $D/libabm_base.so $D/libabm.so $D/libabm_base.so $D/libabm.so 2>&1 | tee gpurun_out/r2_probe_c3_sched.log
