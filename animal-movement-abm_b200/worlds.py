"""Synthetic worlds: the inputs `initialize_model` builds in the reference scripts, generated with numpy.

Reference: scripts/v0.1.jl:66-79 (sheep energies `rand(1:Δenergy*5) - 1`, uniform positions; grass
`fully_grown ~ Bernoulli(1/2)`, `countdown = grown ? regrowth_time : rand(1:regrowth_time) - 1`),
scripts/v0.5.jl:51-68 (elevation 1..40 from a heightmap, walkable iff water_level < e < mountain_level) and
scripts/v0.5.jl:115-122 (grass only on walkable cells).  The heightmap the scripts download is not available
offline, so terrain here is seeded value noise quantised to the same 1..40 range.

Grids are (ny, nx) C-ordered numpy arrays — the memory layout of Julia's (nx, ny) column-major arrays.
"""
from __future__ import annotations

import numpy as np


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=seed))


def make_terrain(nx: int, ny: int, seed: int = 7, octaves: int = 4, base: int = 64):
    """Periodic value-noise elevation quantised to 1..40 (reference scale, scripts/v0.5.jl:56)."""
    rng = _rng(seed)
    h = np.zeros((ny, nx), np.float64)
    amp, total = 1.0, 0.0
    for o in range(octaves):
        cell = max(2, base >> o)
        gx, gy = max(1, -(-nx // cell)), max(1, -(-ny // cell))
        lat = rng.random((gy, gx))
        ys = (np.arange(ny) / cell) % gy
        xs = (np.arange(nx) / cell) % gx
        y0 = np.floor(ys).astype(np.int64); x0 = np.floor(xs).astype(np.int64)
        fy = ys - y0; fx = xs - x0
        fy = fy * fy * (3 - 2 * fy); fx = fx * fx * (3 - 2 * fx)
        y1 = (y0 + 1) % gy; x1 = (x0 + 1) % gx
        a = lat[np.ix_(y0, x0)]; b = lat[np.ix_(y0, x1)]; c = lat[np.ix_(y1, x0)]; d = lat[np.ix_(y1, x1)]
        top = a + (b - a) * fx[None, :]
        bot = c + (d - c) * fx[None, :]
        h += amp * (top + (bot - top) * fy[:, None])
        total += amp
        amp *= 0.5
    h /= total
    # stretch so that the 1 < e < 18 land band of the scripts covers a sizeable, connected share of the map
    lo, hi = np.quantile(h, 0.02), np.quantile(h, 0.98)
    h = np.clip((h - lo) / max(hi - lo, 1e-12), 0.0, 1.0)
    return (np.floor(h * 39 * 0.62) + 1).astype(np.int64)


def _noise_rows(nx: int, ny: int, lattices, rows, cols):
    """Normalised value noise (same construction as make_terrain) at the given global row / column indices."""
    h = np.zeros((len(rows), len(cols)), np.float64)
    amp, total = 1.0, 0.0
    for cell, lat in lattices:
        gy, gx = lat.shape
        ys = (np.asarray(rows) / cell) % gy
        xs = (np.asarray(cols) / cell) % gx
        y0 = np.floor(ys).astype(np.int64); x0 = np.floor(xs).astype(np.int64)
        fy = ys - y0; fx = xs - x0
        fy = fy * fy * (3 - 2 * fy); fx = fx * fx * (3 - 2 * fx)
        y1 = (y0 + 1) % gy; x1 = (x0 + 1) % gx
        top = lat[np.ix_(y0, x0)]
        top += (lat[np.ix_(y0, x1)] - top) * fx[None, :]
        bot = lat[np.ix_(y1, x0)]
        bot += (lat[np.ix_(y1, x1)] - bot) * fx[None, :]
        top += (bot - top) * fy[:, None]
        h += amp * top
        total += amp
        amp *= 0.5
    return h / total


def make_terrain_rows(nx: int, ny: int, y0: int, rows: int, seed: int = 7, octaves: int = 4, base: int = 64, chunk: int = 256):
    """Rows y0 .. y0+rows-1 (wrapped) of a periodic nx x ny elevation map, without building the whole map: what one
    strip of a multi-GPU world needs (its own rows plus ghost rows).  Same value noise as make_terrain; the stretch to
    1..40 uses quantiles of a coarse global sample, so every rank derives the same map from the seed alone."""
    rng = _rng(seed)
    lattices = []
    for o in range(octaves):
        cell = max(2, base >> o)
        gx, gy = max(1, -(-nx // cell)), max(1, -(-ny // cell))
        lattices.append((cell, rng.random((gy, gx))))
    sy, sx = max(1, ny // 1024), max(1, nx // 1024)
    coarse = _noise_rows(nx, ny, lattices, np.arange(0, ny, sy), np.arange(0, nx, sx))
    lo, hi = np.quantile(coarse, 0.02), np.quantile(coarse, 0.98)
    out = np.empty((rows, nx), np.int64)
    cols = np.arange(nx)
    for r0 in range(0, rows, chunk):
        r1 = min(rows, r0 + chunk)
        h = _noise_rows(nx, ny, lattices, (y0 + np.arange(r0, r1)) % ny, cols)
        h = np.clip((h - lo) / max(hi - lo, 1e-12), 0.0, 1.0)
        out[r0:r1] = np.floor(h * 39 * 0.62) + 1
    return out


def walkmap_from_elevation(elevation, water_level: int = 1, mountain_level: int = 18):
    """land_walkmap[i, j] = water_level < elevation[i, j] < mountain_level (scripts/v0.5.jl:64-68)."""
    return ((elevation > water_level) & (elevation < mountain_level)).astype(np.uint8)


def has_walkable_neighbour(walkmap, periodic: bool = True):
    """Cells with at least one walkable Moore neighbour (random_walkmap throws on the others, rand(1:0))."""
    w = np.asarray(walkmap).astype(bool)
    out = np.zeros_like(w)
    src = w if periodic else np.pad(w, 1)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            if periodic:
                out |= np.roll(np.roll(src, dy, 0), dx, 1)
            else:
                out |= src[1 + dy: 1 + dy + w.shape[0], 1 + dx: 1 + dx + w.shape[1]]
    return out


def make_world(nx: int, ny: int, n_sheep: int, seed: int = 321, regrowth_time: int = 30, delta_energy: int = 4,
               walkmap=None, distinct_cells: bool = False, periodic: bool = True, placeable=None):
    """Agents (ids 1..n, 1-based positions, energies 0..5Δ-1) and grass arrays.  `placeable` (optional mask) replaces
    the walkable-with-a-walkable-neighbour rule, e.g. when the neighbours live in another strip's rows."""
    rng = _rng(seed)
    energy = rng.integers(0, delta_energy * 5, n_sheep).astype(np.float64)
    if walkmap is None:
        if distinct_cells:
            lin = rng.permutation(nx * ny)[:n_sheep]
        else:
            lin = rng.integers(0, nx * ny, n_sheep)
    else:
        # random_walkable: uniform over walkable cells; cells whose 8 neighbours are all blocked are left out
        # because the reference's random_walkmap throws there
        mask = placeable if placeable is not None else (np.asarray(walkmap).astype(bool) & has_walkable_neighbour(walkmap, periodic))
        ok = np.flatnonzero(np.asarray(mask).reshape(-1))
        lin = ok[rng.integers(0, ok.size, n_sheep)]  # random_walkable: uniform over walkable cells
    x = (lin % nx).astype(np.int64) + 1
    y = (lin // nx).astype(np.int64) + 1
    grown = rng.integers(0, 2, (ny, nx)).astype(np.uint8)
    cd = np.where(grown == 1, regrowth_time, rng.integers(0, max(regrowth_time, 1), (ny, nx))).astype(np.int64)
    if walkmap is not None:  # scripts/v0.5.jl:115-122: non-walkable cells keep falses/zeros
        w = np.asarray(walkmap).astype(bool)
        grown = np.where(w, grown, 0).astype(np.uint8)
        cd = np.where(w, cd, 0).astype(np.int64)
    ids = np.arange(1, n_sheep + 1, dtype=np.int64)
    return dict(ids=ids, x=x, y=y, energy=energy, fully_grown=grown, countdown=cd)


def weight_lut(visual_distance: int):
    """w[k] = exp(-sqrt(k)/3) for k = 0..2(r+1)^2 (f(x) = exp(-x / 3), src/agent_actions.jl:319)."""
    k = np.arange(2 * (visual_distance + 1) ** 2 + 1, dtype=np.float64)
    return np.exp(-np.sqrt(k) / 3.0)
