"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference (bestquark/mentpy).

The reference tree lives read-only at /root/reference in the build container and does NOT exist
on the GPU box.  This module is only ever used by `oracle/gen_golden.py` (to produce the committed
fixtures under tests/golden/) and by container-only cross-check tests that skip when the tree is
absent.  Nothing in the product package (`mentpy_b200/`) may import it.

The reference unconditionally imports three packages that are not installed here (pennylane,
galois, matplotlib).  None of them is touched at run time by the numpy simulators
(mentpy/simulators/np_simulator_sv.py, np_simulator_dm.py), the templates, the causal-flow finder,
the gradients or the optimizers, so inert stand-ins are enough.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("MENTPY_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mentpy"))


def _stub(name):
    mod = types.ModuleType(name)
    sys.modules[name] = mod
    return mod


def import_reference():
    """Return the reference `mentpy` package (imported from REFERENCE_ROOT, unmodified)."""
    if "mentpy" in sys.modules and getattr(sys.modules["mentpy"], "__b200_shimmed__", False):
        return sys.modules["mentpy"]
    if not reference_available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    if "galois" not in sys.modules:
        g = _stub("galois")
        g.GF = lambda order: (lambda x: np.asarray(x).astype(int) % 2)
    if "pennylane" not in sys.modules:
        q = _stub("pennylane")
        q.math = types.SimpleNamespace(fidelity=lambda a, b: None)

        def _na(*a, **k):
            raise NotImplementedError("pennylane is not installed (stub)")

        q.device = _na
        q.qnode = lambda dev: (lambda f: f)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import mentpy  # noqa: E402

    mentpy.__b200_shimmed__ = True
    return mentpy
