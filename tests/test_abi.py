"""The C-ABI library loads without a GPU and exports every symbol include/cemk.h declares."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from manipulator_mujoco_b200 import _lib
from manipulator_mujoco_b200.kmodel import KModel


@pytest.fixture(scope="module")
def lib():
    _lib.build_library()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "cemk.h")).read()
    declared = set(re.findall(r"\b(cemk_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layout_matches(lib):
    assert lib.cemk_sizeof_kmodel() == C.sizeof(KModel)
    assert lib.cemk_version() >= 1


def test_errors_are_reported_not_thrown(lib):
    h = C.c_void_p()
    assert lib.cemk_create(None, 0, 0, C.byref(h)) == -1            # CEMK_ERR_ARG
    assert b"null" in lib.cemk_last_error()
    km = KModel()
    assert lib.cemk_create(C.byref(km), 16, 0, C.byref(h)) == -3    # CEMK_ERR_MODEL (size mismatch)
    assert lib.cemk_launch_count(None) == 0


def test_no_cpu_fallback_in_product_path():
    """The package never imports the oracle, and the planner refuses to run without CUDA."""
    import torch
    pkg = os.path.join(ROOT, "manipulator_mujoco_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
    if not torch.cuda.is_available():
        from manipulator_mujoco_b200 import cem_planner
        with pytest.raises(RuntimeError):
            cem_planner(num_dof=6, num_batch=8, num_steps=8, timestep=0.05, maxiter_cem=1, num_elite=0.5,
                        w_pos=1.0, w_rot=1.0, w_col=1.0, maxiter_projection=1)
