"""abm_b200 — B200-native stepping engine for the GridSpace sheep model of tatakof/Animal-Movement-ABM.

The package directory is named ``animal-movement-abm_b200`` (not importable as written); import it through the
``abm_b200`` shim at the repository root.  What lives here is only what the hot path needs: ``csrc/`` (CUDA
kernels and the C ABI of ``include/abm.h``), the ctypes binding, and the host-side mirror of the reference's
operator interface (``herbivore_dynamics`` / ``grass_growth_dynamics`` / ``step`` / ``run``).
"""
from . import capi, model, strips, worlds  # noqa: F401
from .capi import (ALTERNATED_WALK, RANDOM_WALK, RANDOM_WALK_ATTRACTOR, RANDOM_WALK_GREGARIOUS,  # noqa: F401
                   RANDOM_WALKMAP, AbmConfig, AbmError, CLib, Engine)
from .model import (EngineAgentStep, EngineModelStep, SheepModel, WalkType, Where, count, count_grass,  # noqa: F401
                    grass_growth_dynamics, grasscolor, herbivore_dynamics, initialize_model, penaltymap_heat, plot_data, run, sheep, step)
