"""Freezes the A0/A1 factor tables the UNMODIFIED reference builds (chsimpy/experiment.py:148-190) into
tests/golden/factor_tables.npz.  The reference builds `rand_values` inline in experiment.main(); this
script runs that main() with a stand-in process pool that does no solves and only captures the table.
Run in the build container (needs /root/reference):  python tests/golden/make_factor_tables.py"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_shim  # noqa: E402

CASES = {
    "uniform_r7": ["-R", "7", "--A-source", "uniform", "--A-seed", "85972"],
    "uniform_r7_indep": ["-R", "7", "--A-source", "uniform", "--A-seed", "85972", "--independent"],
    "sobol_r5": ["-R", "5", "--A-source", "sobol", "--A-seed", "85972"],
    "sobol_r6_indep": ["-R", "6", "--A-source", "sobol", "--A-seed", "11", "--independent"],
    "grid_r10": ["-R", "10", "--A-source", "grid"],
    "grid_r17_indep": ["-R", "17", "--A-source", "grid", "--independent"],
}


def main():
    ref_shim.import_reference()
    import chsimpy.experiment as rex
    captured = {}

    class FakePool:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def imap_unordered(self, fn, items):
            items = list(items)
            captured["rand_values"] = np.array(rex.rand_values, copy=True)
            captured["n_items"] = len(items)
            return iter([(0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0, 0.0, 0, i, 1.0, 1.0) for i in items])

    rex.mp.Pool = FakePool
    out = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for name, argv in CASES.items():
                sys.argv = ["chsimpy-experiment", "-N", "32", "-n", "3", "--file-id", "ft"] + argv
                rex.main()
                out[name] = captured["rand_values"]
                out[name + "_n"] = np.int64(captured["n_items"])
        finally:
            os.chdir(cwd)
    np.savez(os.path.join(HERE, "factor_tables.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
