set -x
python tools/parity_fullsize.py animal-movement-abm_b200/libabm.so animal-movement-abm_b200/libabm_r1.so 2>&1 | tee gpurun_out/r2_parity_fullsize.log
