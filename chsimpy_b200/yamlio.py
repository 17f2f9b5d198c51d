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
    with open(fname, 'w') as f:
        f.write(f"--- {tag}\n")
        yaml.safe_dump(object_scalars(obj), f, default_flow_style=False, width=1000)


def load_mapping(fname):
    with open(fname) as f:
        text = f.read()
    text = re.sub(r'^---\s*![A-Za-z_.]+\s*$', '---', text, count=1, flags=re.M)
    data = yaml.safe_load(text)
    return data if isinstance(data, dict) else {}
