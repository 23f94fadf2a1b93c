"""Generates tests/golden/*.npz from the CPU oracle (oracle/abm_oracle.cpp).

There is no runnable reference in this container (Julia absent; Agents.jl not vendored), and the reference's tests
hold no vectors, so these are ORACLE-GENERATED regression fixtures, labelled as such: they pin the oracle (and
through it the engine) against silent change; they are not outputs of the reference.
Run:  python tests/golden/make_golden.py [case ...]   (no argument: every case of tests/cases.py)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import build, capi, load_case  # noqa: E402
from cases import CASES, build_case  # noqa: E402


def main():
    lib = capi.CLib(build.build_oracle(), "oracle_")
    for name in (sys.argv[1:] or CASES):
        cfgkw, w, state, ticks = build_case(name)
        e = load_case(lib, cfgkw, w)
        if state:
            e.set_global_state(*state)
        out = {"meta": np.frombuffer(json.dumps(dict(case=name, generator="oracle/abm_oracle.cpp", ticks=list(ticks))).encode(), np.uint8)}
        done = 0
        for t in ticks:
            e.step(t - done)
            done = t
            s = e.snapshot()
            for k, v in s.items():
                a = np.asarray(v)
                if k in ("x", "y"):
                    a = a.astype(np.int32)
                elif k == "countdown":
                    a = a.astype(np.int8)
                out[f"t{t}_{k}"] = a
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "agents at end:", s["ids"].size)


if __name__ == "__main__":
    main()
