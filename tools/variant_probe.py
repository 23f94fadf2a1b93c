import sys, json
sys.path.insert(0, '.')
import numpy as np
import abm_b200
from abm_b200 import capi
import importlib.util
spec = importlib.util.spec_from_file_location("bench", "bench.py"); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
cfg, w = b.make_world("C2", 4096, 4096 * 4096 // 4)
for path in ("animal-movement-abm_b200/libabm.so", "animal-movement-abm_b200/libabm_onemaybe.so"):
    lib = capi.CLib(path, "abm_")
    e = b.load(lib, cfg, w, device=0)
    e.set_profiling(True)
    e.step(3)
    tot = {}
    for _ in range(10):
        e.step(1)
        for k, v in e.kernel_stats().items():
            tot[k] = tot.get(k, 0) + v["ms"] / 10
    print(path.split("/")[-1], {k: round(v, 4) for k, v in tot.items()}, flush=True)
    e.close()
