"""A* generation-stamp wrap check on the GPU: thousands of searches through a handful of slots (every slot's 8-bit stamp wraps
more than once), paths and costs compared with the CPU oracle (the expansion counts are printed only: the engine answers
unreachable goals from the walkmap components and counts re-run requests twice, the oracle explores).  `ABM_ASTAR_SLOTS=4 python tools/astar_wrap_check.py`"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from conftest import build, capi, load_case, make_case  # noqa: E402

os.environ.setdefault("ABM_ASTAR_SLOTS", "4")
n, nx = int(sys.argv[1]) if len(sys.argv) > 1 else 3000, 192
cfgkw, w = make_case(nx, nx, n, 3, 1, terrain=True, seed=77, astar_metric=1, astar_penalty=1, counter=9)
ok = np.argwhere(w["walkmap"] == 1)
goals = ok[np.random.default_rng(5).integers(0, len(ok), n)]
res = []
for lib in (capi.CLib(build.build_oracle(), "oracle_"), capi.CLib(build.LIBABM, "abm_")):
    e = load_case(lib, cfgkw, w)
    e.plan_routes(w["ids"], goals[:, 1] + 1, goals[:, 0] + 1)
    off, px, py, cost = e.get_paths(w["ids"])
    res.append((off, px, py, cost, e.timing()["astar_expansions"]))
    e.close()
a, b = res
same = all(np.array_equal(a[i], b[i]) for i in range(4))
print("paths/costs identical:", same, "| expansions oracle", int(a[4]), "engine", int(b[4]), "| searches per slot", n // int(os.environ["ABM_ASTAR_SLOTS"]))
sys.exit(0 if same else 1)
