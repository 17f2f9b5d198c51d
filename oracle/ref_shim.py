"""Import shim for the UNMODIFIED reference package (TEST INFRASTRUCTURE ONLY).

The reference (`/root/reference/chsimpy`) cannot be imported as-is in this image:
`ruamel.yaml`, `opensimplex`, `matplotlib` and `seaborn` are absent.  None of them
is touched by the hot path (solver.py / solution.py / timedata.py / utils.py), so we
pre-register inert stub modules and then import the reference package itself.

Used only by `tests/golden/make_golden.py` (fixture generation, in the build
container) and by `oracle/validate_oracle.py`.  Nothing in the product path, the
`-m gpu` tests, `smoke()` or `bench.py` imports this file: `/root/reference` does not
exist on the GPU box.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CHS_REFERENCE_ROOT", "/root/reference")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    if "ruamel.yaml" not in sys.modules:
        class _Ctor:
            def add_constructor(self, *a, **k):
                pass

        class YAML:
            def __init__(self, *a, **k):
                self.constructor = _Ctor()

            def register_class(self, cls):
                return cls

            def dump(self, *a, **k):
                raise RuntimeError("ruamel.yaml is stubbed")

            def load(self, *a, **k):
                raise RuntimeError("ruamel.yaml is stubbed")

        ry = _stub("ruamel.yaml", YAML=YAML)
        _stub("ruamel", yaml=ry)
    if "opensimplex" not in sys.modules:
        def noise2array(*a, **k):
            raise RuntimeError("opensimplex is stubbed")
        _stub("opensimplex", noise2array=noise2array)
    if "matplotlib" not in sys.modules:
        class _Any:
            def __getattr__(self, n):
                return _Any()

            def __call__(self, *a, **k):
                return _Any()
        mpl = _stub("matplotlib", use=lambda *a, **k: None)
        for sub in ("pyplot", "colors", "gridspec", "ticker", "cm", "animation"):
            sm = _stub("matplotlib." + sub)
            sm.__getattr__ = lambda n: _Any()  # PEP 562
            setattr(mpl, sub, sm)
    if "seaborn" not in sys.modules:
        sb = _stub("seaborn")
        sb.__getattr__ = lambda n: (lambda *a, **k: None)


def import_reference():
    """Returns the reference `chsimpy` package (unmodified sources)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "chsimpy")):
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import chsimpy  # noqa
    return chsimpy
