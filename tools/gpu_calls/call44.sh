ABM_ASTAR_SLOTS=4 python tools/astar_wrap_check.py 3000 2>&1 | tail -2
ABM_ASTAR_SLOTS=64 python tools/astar_wrap_check.py 40000 2>&1 | tail -2
