for r in 2 3 4 5 6 8 12; do
echo "ABM_TILE_ROUNDS=$r"
ABM_TILE_ROUNDS=$r python tools/variant_probe.py --workload C2 --size 4096 --ticks 30 animal-movement-abm_b200/libabm_base.so animal-movement-abm_b200/libabm.so 2>&1 | tee -a gpurun_out/r2_probe_rounds.log
done
