"""Load the REAL reference NumPy code (ssdutils / utils / transforms) from
/root/reference under a stub ``tensorflow`` module -- TEST INFRASTRUCTURE ONLY.

The reference's ``utils.py:25`` imports TensorFlow at module top but only uses
it inside function bodies the hot path never calls, so an empty stand-in module
is enough.  The reference modules are registered under
private names (``_ref_utils`` ...) so they never shadow the product modules of
the same name.  /root/reference exists only in the authoring container; on the
GPU box ``available()`` is False and the tests fall back to the committed
fixtures in tests/golden/.
"""
import importlib.util
import os
import sys
import types

REF_DIR = os.environ.get('SSD_REFERENCE_DIR', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF_DIR, 'ssdutils.py'))


_cache = {}


def load():
    """Returns (ref_utils, ref_ssdutils, ref_transforms) modules."""
    if 'mods' in _cache:
        return _cache['mods']
    if not available():
        raise RuntimeError('reference tree not present at ' + REF_DIR)
    saved = {k: sys.modules.get(k) for k in ('tensorflow', 'utils', 'ssdutils', 'transforms')}
    sys.modules['tensorflow'] = types.ModuleType('tensorflow')
    mods = []
    try:
        for name in ('utils', 'ssdutils', 'transforms'):
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF_DIR, name + '.py'))
            m = importlib.util.module_from_spec(spec)
            sys.modules[name] = m          # reference modules import each other by bare name
            spec.loader.exec_module(m)
            mods.append(m)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    for name, m in zip(('_ref_utils', '_ref_ssdutils', '_ref_transforms'), mods):
        sys.modules[name] = m
    _cache['mods'] = tuple(mods)
    return _cache['mods']


def load_ap():
    """The reference's average_precision module (needs the NumPy-1 alias ``np.int``, removed in NumPy 1.24: supplied here,
    in the test process only, because the reference file cannot be edited)."""
    if 'ap' in _cache:
        return _cache['ap']
    import numpy as np
    ru, rs, _ = load()
    if not hasattr(np, 'int'):
        np.int = int
    if not hasattr(np, 'bool'):
        np.bool = bool
    saved = {k: sys.modules.get(k) for k in ('utils', 'ssdutils')}
    sys.modules['utils'] = ru; sys.modules['ssdutils'] = rs
    try:
        spec = importlib.util.spec_from_file_location('_ref_average_precision', os.path.join(REF_DIR, 'average_precision.py'))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache['ap'] = m
    return m
