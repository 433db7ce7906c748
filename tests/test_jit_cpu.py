"""CPU: the run-time specialised state-vector kernel (mentpy_b200/csrc/sv_jit_src.inc) is generated
and compiled with NVRTC for every golden pattern it covers -- NVRTC compiles offline, so the
generated source is checked in the build container; parity of its results is a GPU test
(tests/test_cuda_jit.py)."""
import pytest

import mentpy_b200 as mb
from conftest import load_golden
from mentpy_b200 import _lib, plan as plan_mod

CASES = [c for c in load_golden("sim_cases.json")["cases"] if c["backend"] == "numpy-sv"]


def _circuit(case):
    name, args, kwargs = case["spec"]
    gs = getattr(mb.templates, name)(*args, **kwargs)
    for v in case["x_nodes"]:
        gs[v] = mb.Ment("X")
    for v, (ang, plane) in case["fixed"].items():
        gs[int(v)] = mb.Ment(ang, plane)
    return gs


def _nvrtc_available():
    lib = _lib.load()
    lowered = plan_mod.lower(mb.templates.linear_cluster(5))
    dp = plan_mod.DevicePlan(lowered, host_only=True)
    try:
        return dp.jit_compile_check() > 0
    except NotImplementedError:
        return False
    finally:
        del dp, lib


@pytest.mark.parametrize("case", CASES, ids=[f"{c['spec'][0]}{c['spec'][1]}-w{c['window_size']}-s{c['seed']}" for c in CASES])
def test_specialised_kernel_compiles_for_golden_patterns(case):
    if not _nvrtc_available():
        pytest.skip("libnvrtc not found: " + _lib.load().mbqc_jit_info().decode())
    lowered = plan_mod.lower(_circuit(case), window_size=case["window_size"])
    dp = plan_mod.DevicePlan(lowered, host_only=True)
    for out_form in (_lib.OUT_SV, _lib.OUT_DM):
        size = dp.jit_compile_check(out_form=out_form, cta=128)
        if lowered.window <= _lib.MAX_WINDOW_REG and lowered.window >= 2:
            assert size > 1000, "pattern is in scope (reference schedule, window <= 5) but no kernel was built"
        else:
            assert size == 0


def test_host_only_plan_cannot_run():
    import ctypes as C

    lowered = plan_mod.lower(mb.templates.grid_cluster(2, 6))
    dp = plan_mod.DevicePlan(lowered, host_only=True)
    lib = _lib.load()
    assert lib.mbqc_plan_window(dp.handle) == 3 and lib.mbqc_plan_num_steps(dp.handle) == 10
    rc = lib.mbqc_run_batch_sv(dp.handle, C.c_void_p(16), 10, None, _lib.INPUT_PLUS, 4, C.c_void_p(16), _lib.OUT_SV, None, None)
    assert rc == _lib.MBQC_E_ARG and b"host-only" in lib.mbqc_last_error()


DM_CASES = [c for c in load_golden("sim_cases.json")["cases"] if c["backend"] == "numpy-dm"]


@pytest.mark.parametrize("case", DM_CASES, ids=[f"{c['spec'][0]}{c['spec'][1]}-w{c['window_size']}-s{c['seed']}" for c in DM_CASES])
def test_specialised_dm_kernel_compiles_for_golden_patterns(case):
    """dm_jit_src.inc + the preamble of dm_jit_gen.h (layout schedule, sign tables) compile for
    every golden density-matrix pattern in scope: window 2..5, no plane-Z node."""
    if not _nvrtc_available():
        pytest.skip("libnvrtc not found: " + _lib.load().mbqc_jit_info().decode())
    lowered = plan_mod.lower(_circuit(case), window_size=case["window_size"], mixed=True)
    noise = plan_mod.noise_from_kraus(plan_mod.kraus_set("depolarizing", p=0.01)) if case["seed"] % 2 else None
    dp = plan_mod.DevicePlan(lowered, noise=noise, host_only=True)
    size = dp.jit_compile_check(out_form=200)
    has_z = any(st.plane == _lib.PLANE_Z for st in lowered.steps)
    if 2 <= lowered.window <= 5 and not has_z:
        assert size > 1000
    else:
        assert size == 0


def test_gradient_kernel_variants_and_dm_layouts_compile():
    """Every form of the specialised gradient kernel -- plain, and the replicated-result forms that
    store into all GPUs' copies (coalesced stores, bulk copies with full / read-only wait, multimem
    stores to a multicast address) -- and both register / lane layouts of the density-matrix kernel."""
    if not _nvrtc_available():
        pytest.skip("libnvrtc not found: " + _lib.load().mbqc_jit_info().decode())
    dp = plan_mod.DevicePlan(plan_mod.lower(mb.templates.grid_cluster(4, 5)), host_only=True)
    for form in (100, 101, 102, 103, 104):
        assert dp.jit_compile_check(out_form=form) > 1000
    dm = plan_mod.DevicePlan(plan_mod.lower(mb.templates.grid_cluster(3, 8), mixed=True), host_only=True)
    sizes = {lb: dm.jit_compile_check(out_form=200, cta=lb) for lb in (1, 2)}
    assert sizes[1] > 1000 and sizes[2] > sizes[1]  # 16 entries per lane unroll to more code than 4
    gs = mb.templates.linear_cluster(5)
    gs[2] = mb.ControlMent(gs[0].outcome, None, "XY", 0, "X")
    ctl = plan_mod.DevicePlan(plan_mod.lower(gs, mixed=True), host_only=True)
    assert ctl.jit_compile_check(out_form=200) == 0  # controlled steps stay on the general kernels
