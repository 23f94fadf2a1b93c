D=animal-movement-abm_b200
python tools/variant_probe.py --workload C2 --size 4096 --ticks 20 $D/libabm_base.so $D/libabm_v4agg.so $D/libabm_v4noagg.so $D/libabm_v4flat.so $D/libabm_v4flatagg.so 2>&1 | tee gpurun_out/r2_probe_v4variants.log
ABM_TILE_STAMPS=0 python tools/variant_probe.py --workload C2 --size 4096 --ticks 20 $D/libabm_v4noagg.so $D/libabm_v4flat.so 2>&1 | tee -a gpurun_out/r2_probe_v4variants.log
