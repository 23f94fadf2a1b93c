for m in 4 5 6 7 8; do
echo "ABM_TILE_MARGIN=$m"
ABM_TILE_MARGIN=$m python tools/variant_probe.py --workload C2 --size 4096 --ticks 30 animal-movement-abm_b200/libabm.so 2>&1 | tee -a gpurun_out/r2_probe_margin.log
done
