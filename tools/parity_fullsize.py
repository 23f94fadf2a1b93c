"""Full-size C2 world: each library given on the command line against the CPU oracle, tick by tick; prints where they
differ (agent id, positions, energies) instead of asserting.  Round-1 builds lack the newer entry points: only the wide
ABI is used."""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import abm_b200
from abm_b200 import build, capi, worlds

class OldLib(capi.CLib):
    def __init__(self, path, prefix):
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path)
        self.fn = {}
        for name, (res, args) in capi.SIGNATURES.items():
            if hasattr(self.dll, prefix + name):
                f = getattr(self.dll, prefix + name); f.restype, f.argtypes = res, args; self.fn[name] = f

size = int(os.environ.get("SIZE", "4096")); ticks = int(os.environ.get("TICKS", "2")); seed = int(os.environ.get("SEED", "5"))
w = worlds.make_world(size, size, size * size // 4, seed=seed, periodic=True)
cfg = dict(nx=size, ny=size, periodic=1, walk_type=2, seed=77, visual_distance=3, prob_random_walk=0.9, reproduction_prob=0.001)
orc = capi.CLib(build.build_oracle(), "oracle_")
def load(lib):
    e = lib.create(lib.default_config(**cfg)); e.set_grid(w["fully_grown"], w["countdown"]); e.set_agents(w["ids"], w["x"], w["y"], w["energy"]); return e
a = load(orc)
engines = [(p, load(OldLib(p, "abm_"))) for p in sys.argv[1:]]
prev = a.snapshot()
for t in range(ticks):
    a.step(1); sa = a.snapshot()
    for p, e in engines:
        e.step(1); sb = e.snapshot()
        for k in ("ids", "x", "y", "energy", "fully_grown", "countdown"):
            if sa[k].shape != sb[k].shape: print(os.path.basename(p), "tick", t, k, "shape differs", sa[k].shape, sb[k].shape); continue
            bad = np.flatnonzero(np.asarray(sa[k]).reshape(-1) != np.asarray(sb[k]).reshape(-1))
            if bad.size:
                print(os.path.basename(p), "tick", t, k, "differs at", bad[:6], bad.size, "places")
                if k in ("x", "y"):
                    for i in bad[:4]:
                        pi = np.searchsorted(prev["ids"], sa["ids"][i])
                        print("   id", sa["ids"][i], "was", (prev["x"][pi], prev["y"][pi]), "oracle ->", (sa["x"][i], sa["y"][i]), "engine ->", (sb["x"][i], sb["y"][i]), "energy", sa["energy"][i], sb["energy"][i])
        else:
            pass
        print(os.path.basename(p), "tick", t, "checked", flush=True)
    prev = sa
