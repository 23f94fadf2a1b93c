"""Shared fixtures.  `-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic, ABI surface, and
the engine sources compiled for the host (tests/emu, see csrc/abm_platform.h).  `-m gpu` runs on a B200 and calls
the product library libabm.so through its C ABI."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import abm_b200  # noqa: E402
from abm_b200 import build, capi, worlds  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle():
    return capi.CLib(build.build_oracle(), "oracle_")


@pytest.fixture(scope="session")
def emu():
    return capi.CLib(build.build_emu(), "emu_")


@pytest.fixture(scope="session")
def libabm():
    """The product library.  On the GPU box it was built in-tree before the snapshot; rebuilt if stale."""
    return capi.CLib(build.build_libabm(), "abm_")


def make_case(nx, ny, n, walk_type, periodic, seed=321, terrain=False, **kw):
    """(config kwargs, world) of a small synthetic case shared by every parity test."""
    wm = el = None
    if terrain:
        el = worlds.make_terrain(nx, ny, seed=seed + 1, base=16)
        wm = worlds.walkmap_from_elevation(el)
    cfgkw = dict(nx=nx, ny=ny, walk_type=walk_type, periodic=periodic, attractor_x=nx / 2, attractor_y=ny / 2,
                 seed=seed)
    cfgkw.update(kw)
    w = worlds.make_world(nx, ny, n, seed=seed, regrowth_time=cfgkw.get("regrowth_time", 30), walkmap=wm, periodic=bool(periodic))
    w["walkmap"], w["elevation"] = wm, el
    return cfgkw, w


def load_case(lib, cfgkw, w, lut=None):
    e = lib.create(lib.default_config(**cfgkw))
    e.set_grid(w["fully_grown"], w["countdown"], w.get("walkmap"), w.get("elevation"))
    e.set_agents(w["ids"], w["x"], w["y"], w["energy"])
    if lut is not None:
        e.set_weight_lut(lut)
    return e


def assert_same_state(a, b, ctx=""):
    """Bit-exact on every integer field; energy: exact expected, 1e-6 relative is the stated tolerance."""
    for k in ("tick", "behav", "behav_counter", "next_id"):
        assert a[k] == b[k], f"{ctx}: {k} differs: {a[k]} vs {b[k]}"
    assert a["ids"].shape == b["ids"].shape, f"{ctx}: population differs: {a['ids'].size} vs {b['ids'].size}"
    for k in ("ids", "x", "y", "fully_grown", "countdown"):
        if not np.array_equal(a[k], b[k]):
            bad = np.flatnonzero(np.asarray(a[k]).reshape(-1) != np.asarray(b[k]).reshape(-1))
            raise AssertionError(f"{ctx}: {k} differs at {bad[:8]} ({bad.size} places)")
    np.testing.assert_allclose(a["energy"], b["energy"], rtol=1e-6, atol=0, err_msg=f"{ctx}: energy")
