"""MJCF compiler + model tables: ids / counts pinned by the reference, consistency of derived tables."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from manipulator_mujoco_b200 import kmodel as KM
from manipulator_mujoco_b200.mjcf import compile_mjcf, host_kinematics, host_mass_matrix, load_model

IDS = json.load(open(os.path.join(GOLDEN, "scene_ids.json")))
REF_XML = "/root/reference/sampling_based_planner/ur5e_hande_mjx/scene.xml"


def test_counts_and_ids(mc):
    assert [mc.geom_id(f"robot_{i}") for i in range(10)] == IDS["robot_geom_ids"]      # view_traj_mjx.py:54
    assert (mc.nq, mc.nv, mc.nbody, mc.ngeom) == (IDS["nq"], IDS["nv"], IDS["nbody"], IDS["ngeom"])
    assert mc.ncon == IDS["ncon"] and len(mc.pair_geom) == IDS["npair"]
    from oracle.oracle import robot_slot_mask
    assert robot_slot_mask(mc).sum() == IDS["nrobot_slots"]                             # mask.sum() of mjx_planner.py:115
    assert mc.opt["iterations"] == 1 and mc.opt["ls_iterations"] == 5 and mc.opt["eulerdamp"] == 0


def test_forward_kinematics_pins(mc):
    tcp = mc.site_id("tcp")
    q = mc.qpos0.copy()
    xpos, xquat, xmat = host_kinematics(mc, q)
    b = mc.site_body[tcp]
    np.testing.assert_allclose(xpos[b] + xmat[b] @ mc.site_pos[tcp], IDS["tcp_at_zero"], atol=1e-9)
    q[:6] = IDS["init_pos"]
    xpos, xquat, xmat = host_kinematics(mc, q)
    np.testing.assert_allclose(xpos[b] + xmat[b] @ mc.site_pos[tcp], IDS["tcp_at_init_pos"], atol=1e-5)
    np.testing.assert_allclose(xquat[mc.body_id("hande")], IDS["xquat_hande_at_init_pos"], atol=1e-5)


@pytest.mark.skipif(not os.path.exists(REF_XML), reason="reference checkout not present (GPU box)")
def test_packaged_constants_match_the_xml(mc):
    fresh = compile_mjcf(REF_XML)
    for k, v in fresh.d.items():
        if isinstance(v, np.ndarray):
            np.testing.assert_allclose(v, mc.d[k], rtol=0, atol=1e-12, err_msg=k)
        elif k != "xml":
            assert v == mc.d[k], k


def test_hande_inertia_from_geoms(mc):
    """No <inertial> on `hande`: mass from coupler + hande meshes + robot_0 capsule at density 1000."""
    m = mc.body_mass[mc.body_id("hande")]
    cap = 1000 * (np.pi * 0.04 ** 2 * 0.14 + 4 / 3 * np.pi * 0.04 ** 3)
    assert abs(cap - 0.9718) < 1e-3
    assert 1.30 < m < 1.40 and m > cap


def test_mass_matrix_spd_and_invweights(mc):
    q = mc.qpos0.copy()
    q[:6] = IDS["init_pos"]
    M, _, _ = host_mass_matrix(mc, q)
    np.testing.assert_allclose(M, M.T, atol=1e-12)
    assert np.linalg.eigvalsh(M).min() > 0
    assert abs(mc.body_invweight0[mc.body_id("target_0"), 0] - 1 / 0.064) < 1e-9      # free box: 1/m
    assert np.all(mc.dof_invweight0 > 0)


def test_kmodel_flattening(mc):
    km, info = KM.build_kmodel(mc, 0.05)
    assert km.nl == 6 and km.ncap == 10 and km.nsbox == 6 and km.has_box == 1
    assert km.nrpair == 107 and km.nbpair == 7 and km.nslot_robot == 187
    # merged link masses == sum of the bodies welded to each link
    masses = [km.l_inertia[i][6] for i in range(6)]
    assert abs(sum(masses) - mc.body_mass[3:12].sum()) < 1e-5
    # the pass-major pair table covers every robot slot exactly once
    seen = np.zeros(187, dtype=int)
    for e in range(KM.MAXRPAIR):
        x = km.rp[e]
        ty, a, b, sl = (x & 15) - 1, (x >> 4) & 15, (x >> 8) & 15, x >> 16
        if ty == KM.KP_NONE:
            continue
        n = 1 if ty == KM.KP_CAP_CAP else 2
        seen[sl:sl + n] += 1
        if ty == KM.KP_CAP_BOX:                       # pass = box, lane = capsule
            assert e // KM.LANE_GROUP == b and e % KM.LANE_GROUP == a and e // KM.LANE_GROUP < km.ncbpass
        else:
            assert e // KM.LANE_GROUP >= km.ncbpass
    assert (seen == 1).all()
    assert km.ncbpass == 7
    assert C.sizeof(km) % 4 == 0


def test_unsupported_models_are_refused(mc):
    import copy
    bad = copy.deepcopy(mc)
    bad.d["opt"] = dict(mc.opt, iterations=4)
    with pytest.raises(NotImplementedError):
        KM.build_kmodel(bad, 0.05)
