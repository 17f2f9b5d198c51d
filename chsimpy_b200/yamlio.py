"""YAML round-trip of Parameters / Solution scalars (reference parameters.py:66-101,
solution.py:69-92) on PyYAML.  The files carry the reference's tags (`--- !Parameters`)
so either package can read what the other wrote."""
import inspect
import re

import numpy as np
import yaml


def _plain(v):
    if isinstance(v, (np.floating,)):
        return float(v)
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.bool_,)):
        return bool(v)
    try:
        import sympy
        if isinstance(v, sympy.Float):
            return float(v)
    except ImportError:
        pass
    return v


def object_scalars(obj):
    out = {}
    for name in dir(obj):
        if name.startswith('_'):
            continue
        try:
            v = getattr(obj, name)
        except Exception:
            continue
        if callable(v):
            if getattr(v, "__name__", "") == "<lambda>":
                try:
                    src = str(inspect.getsourcelines(v)[0][0])
                except (OSError, TypeError):
                    continue
                src = re.sub(r'#[^\n]*', '', src)
                src = re.sub(r'\s+', '', src).replace('lambda', 'lambda ')
                out[name] = src
            continue
        v = _plain(v)
        if isinstance(v, (bool, int, float, str)) or v is None:
            out[name] = v
    return out


def dump_object(obj, fname, tag):
    data = object_scalars(obj)
    with open(fname, 'w') as f:
        f.write(f"--- {tag}\n")
        yaml.safe_dump(data, f, default_flow_style=False, width=1000)
        # a Solution carries its Parameters as a nested tagged mapping, as the reference's dump does
        # (solution.py:69-92: `params` is a registered class, so ruamel writes `params: !Parameters`)
        sub = getattr(obj, "params", None)
        if sub is not None and hasattr(sub, "yaml_export_scalars"):
            f.write("params: !Parameters\n")
            body = yaml.safe_dump(object_scalars(sub), default_flow_style=False, width=1000)
            f.write("".join("  " + ln + "\n" for ln in body.splitlines()))


class _Loader(yaml.SafeLoader):
    """SafeLoader that knows the reference's tags (utils.py:50-76, parameters.py:66, solution.py:69)."""


def _construct_parameters(loader, node):
    from .parameters import Parameters
    p = Parameters()
    for k, v in loader.construct_mapping(node, deep=True).items():
        if k in ("func_A0", "func_A1"):
            continue                      # dumped as source text only; the defaults stay callable
        setattr(p, k, v)
    return p


def _construct_solution(loader, node):
    from .solution import Solution
    s = Solution.__new__(Solution)        # no __init__: the file carries the derived scalars (as ruamel's loader does)
    s.U, s.timedata, s._eig = None, None, None
    for k, v in loader.construct_mapping(node, deep=True).items():
        setattr(s, k, v)
    return s


def _construct_ndarray(loader, node):
    import ast
    return np.array(ast.literal_eval(loader.construct_scalar(node).replace('\n', '')))


_Loader.add_constructor('!Parameters', _construct_parameters)
_Loader.add_constructor('!Solution', _construct_solution)
_Loader.add_constructor('!ndarray', _construct_ndarray)
_Loader.add_constructor('!numpy.float64', lambda loader, node: float(loader.construct_scalar(node)))


def load_object(fname):
    """reference utils.yaml_import: the Parameters / Solution instance a file written by
    yaml_export_scalars (of either package) describes."""
    with open(fname) as f:
        return yaml.load(f, Loader=_Loader)


def load_mapping(fname):
    with open(fname) as f:
        text = f.read()
    text = re.sub(r'^---\s*![A-Za-z_.]+\s*$', '---', text, count=1, flags=re.M)
    data = yaml.safe_load(text)
    return data if isinstance(data, dict) else {}
