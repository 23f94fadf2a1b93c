"""Bounded runs of tools/fuzz.py: random configurations, engine kernels (emulated launcher) vs the CPU oracle, and the
oracle vs the independent Python restatement.  Longer runs: `python tools/fuzz.py engine 0 2000`."""
import os
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import fuzz  # noqa: E402


@pytest.mark.parametrize("start", [100000, 100040, 100080])
def test_engine_kernels_match_oracle_on_random_configurations(start):
    results = {s: fuzz.engine_vs_oracle(s) for s in range(start, start + 40)}
    bad = {s: r for s, r in results.items() if r.split()[0] in ("DIFF", "ERROR")}
    assert not bad, bad
    assert sum(r == "ok" for r in results.values()) >= 25


def test_engine_general_path_and_single_round_windows_on_random_configurations(monkeypatch):
    monkeypatch.setenv("FORCE_GENERAL", "1")
    monkeypatch.setenv("ABM_TILE_ROUNDS", "1")
    results = {s: fuzz.engine_vs_oracle(s) for s in range(200000, 200040)}
    bad = {s: r for s, r in results.items() if r.split()[0] in ("DIFF", "ERROR")}
    assert not bad, bad


def test_oracle_matches_second_restatement_on_random_configurations():
    results = {s: fuzz.oracle_vs_pyref(s) for s in range(300000, 300300)}
    bad = {s: r for s, r in results.items() if r.split()[0] in ("DIFF", "ERROR")}
    assert not bad, bad
    assert sum(r == "ok" for r in results.values()) >= 250


def test_astar_api_matches_oracle_on_random_terrain(monkeypatch):
    monkeypatch.setenv("ABM_ASTAR_SLOTS", "1")       # the fuzzer overwrites both per seed; monkeypatch restores them
    monkeypatch.setenv("ABM_ROUTE_POOL_GUESS", "5")
    results = {s: fuzz.astar_vs_oracle(s) for s in range(400000, 400150)}
    bad = {s: r for s, r in results.items() if r.split()[0] == "DIFF"}
    assert not bad, bad
