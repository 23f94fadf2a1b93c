"""examples/scripts_v0x.py — the reference's six grid scripts through the host mirror — on the emulated engine, step for step
against the oracle (same functions, other library object).  On a B200 `python examples/scripts_v0x.py` runs them on libabm.so;
the GPU parity of the same rules and API calls is tests/test_host_api.py and tests/test_parity.py under -m gpu."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT

spec = importlib.util.spec_from_file_location("scripts_v0x", os.path.join(ROOT, "examples", "scripts_v0x.py"))
S = importlib.util.module_from_spec(spec)
spec.loader.exec_module(S)


def _same(a, b):
    ia, xa, ya, ea = a.agents()
    ib, xb, yb, eb = b.agents()
    oa, ob = np.argsort(ia), np.argsort(ib)
    assert np.array_equal(ia[oa], ib[ob]) and np.array_equal(xa[oa], xb[ob]) and np.array_equal(ya[oa], yb[ob])
    assert np.array_equal(ea[oa], eb[ob])
    assert np.array_equal(a.fully_grown, b.fully_grown) and np.array_equal(a.countdown, b.countdown)


def _run_both(name, a_lib, b_lib, frames):
    (ma, oa), (mb, ob) = S.SCRIPTS[name](lib=a_lib, frames=frames), S.SCRIPTS[name](lib=b_lib, frames=frames)
    _same(ma, mb)
    if isinstance(oa, dict):
        assert oa["agent_df"].equals(ob["agent_df"]) and oa["model_df"].equals(ob["model_df"]) and oa["heatarray_shape"] == ob["heatarray_shape"]
        assert list(oa["agent_df"].columns) == ["step", "count_sheep"] and len(oa["agent_df"]) == frames + 1
    else:
        assert oa == ob and oa[-1]["frame"] == frames


@pytest.mark.parametrize("name", sorted(S.SCRIPTS))
def test_scripts_on_the_emulated_engine_match_the_oracle(emu, oracle, name):
    _run_both(name, emu, oracle, 30)
