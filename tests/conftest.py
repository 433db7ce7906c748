import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests need a CUDA device and the built extension: skip them (instead of failing
    349 times) where either is missing, so that CPU runs show real regressions only."""
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    from mentpy_b200 import _lib

    reason = None
    if not have_gpu:
        reason = "no CUDA device"
    elif not os.path.exists(_lib.LIB_PATH):
        reason = f"{_lib.LIB_PATH} not built"
    if reason:
        skip = pytest.mark.skip(reason=reason)
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def from_cplx(d):
    if d is None:
        return None
    a = np.asarray(d["re"], dtype=float) + 1j * np.asarray(d["im"], dtype=float)
    return a.reshape(d["shape"])


def infidelity_pure(a, b):
    """1 - |<a|b>|^2 / (<a|a><b|b>)"""
    a = np.asarray(a).reshape(-1)
    b = np.asarray(b).reshape(-1)
    return abs(1.0 - abs(np.vdot(a, b)) ** 2 / (np.vdot(a, a).real * np.vdot(b, b).real))


def dm_distance(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))))


@pytest.fixture(scope="session")
def golden_sim_cases():
    return load_golden("sim_cases.json")["cases"]


@pytest.fixture(scope="session")
def golden_structures():
    return load_golden("structures.json")["records"]


def build_spec(spec):
    """Circuit for a golden `spec` = (template | "vstack" | "hstack" | "merge", args, kwargs)."""
    import mentpy_b200 as mb

    name, args, kwargs = spec
    if name in ("vstack", "hstack"):
        return getattr(mb, name)([build_spec(a) for a in args])
    if name == "merge":
        return mb.merge(build_spec(args[0]), build_spec(args[1]), along=[tuple(x) for x in args[2]])
    return getattr(mb.templates, name)(*args, **kwargs)
