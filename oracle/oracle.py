"""ctypes front-end of the CPU oracle (oracle/mjstep.c) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, bench.py's ``cpu_baseline`` / ``--impl reference`` leg and ``__graft_entry__.smoke()``
import this module.  PARITY UNPINNED against real MJX (see the header of mjstep.c).

The planner-level algebra of the reference (projection filter, cost, elite selection, mean /
covariance update; ``mjx_planner.py:181-335``) is restated in numpy float64 in
``oracle/planner_ref.py``; this file only exposes the rigid-body step / rollout.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAXB, MAXJ, MAXV, MAXQ, MAXG, MAXP, MAXCON = 24, 12, 16, 20, 24, 128, 256

_d, _i = C.c_double, C.c_int


class OModel(C.Structure):
    _fields_ = [
        ("nq", _i), ("nv", _i), ("nbody", _i), ("njnt", _i), ("ngeom", _i), ("npair", _i), ("ncon", _i),
        ("iterations", _i), ("ls_iterations", _i), ("tcp_body", _i), ("hande_body", _i), ("capbox_mode", _i),
        ("timestep", _d), ("tolerance", _d), ("ls_tolerance", _d), ("impratio", _d), ("meaninertia", _d),
        ("gravity", _d * 3), ("tcp_pos", _d * 3),
        ("body_parent", _i * MAXB), ("body_jnt", _i * MAXB), ("body_rootid", _i * MAXB), ("body_weldid", _i * MAXB),
        ("body_pos", _d * 3 * MAXB), ("body_quat", _d * 4 * MAXB), ("body_mass", _d * MAXB), ("body_ipos", _d * 3 * MAXB),
        ("body_inertia", _d * 9 * MAXB), ("body_gravcomp", _d * MAXB), ("body_invweight0", _d * MAXB),
        ("jnt_type", _i * MAXJ), ("jnt_body", _i * MAXJ), ("jnt_qposadr", _i * MAXJ), ("jnt_dofadr", _i * MAXJ),
        ("jnt_limited", _i * MAXJ),
        ("jnt_axis", _d * 3 * MAXJ), ("jnt_range", _d * 2 * MAXJ), ("jnt_armature", _d * MAXJ), ("jnt_damping", _d * MAXJ),
        ("jnt_solref", _d * 2 * MAXJ), ("jnt_solimp", _d * 5 * MAXJ), ("jnt_margin", _d * MAXJ),
        ("dof_invweight0", _d * MAXV),
        ("geom_type", _i * MAXG), ("geom_body", _i * MAXG),
        ("geom_pos", _d * 3 * MAXG), ("geom_quat", _d * 4 * MAXG), ("geom_size", _d * 3 * MAXG),
        ("geom_friction", _d * 3 * MAXG), ("geom_solref", _d * 2 * MAXG), ("geom_solimp", _d * 5 * MAXG),
        ("pair_g1", _i * MAXP), ("pair_g2", _i * MAXP), ("pair_slotadr", _i * MAXP), ("pair_nslot", _i * MAXP),
        ("slot_robot", _i * MAXCON),
    ]


def build(force=False):
    """Compile oracle/mjstep.c (gcc) into oracle/_build/; a no-op when the .so files are fresh."""
    out = os.path.join(_HERE, "_build")
    src = os.path.join(_HERE, "mjstep.c")
    libs = [os.path.join(out, n) for n in ("liboracle_f64.so", "liboracle_f32.so", "liboracle_count.so")]
    hdr = os.path.join(_HERE, "count_real.h")
    fresh = all(os.path.exists(p) and os.path.getmtime(p) >= max(os.path.getmtime(src), os.path.getmtime(hdr)) for p in libs)
    if force or not fresh:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return libs


def _fill(dst, src):
    a = np.asarray(src)
    flat = np.ctypeslib.as_array(dst).reshape(-1)
    flat[:a.size] = a.reshape(-1)


def robot_slot_mask(mc, robot_names=None):
    """Boolean mask over the ncon contact slots: pair contains a ``robot_i`` geom (mjx_planner.py:113-115)."""
    if robot_names is None:
        robot_names = [f"robot_{i}" for i in range(10)]
    ids = {mc.geom_id(n) for n in robot_names if n in mc.geom_names}
    mask = np.zeros(mc.ncon, dtype=bool)
    for (g1, g2), a, n in zip(mc.pair_geom, mc.pair_slotadr, mc.pair_nslot):
        if g1 in ids or g2 in ids:
            mask[a:a + n] = True
    return mask


class Oracle:
    """CPU reference stepper for a compiled scene (ModelConsts)."""

    def __init__(self, mc, timestep, dtype="f64", tcp_site="tcp", hande_body="hande", capbox_mode=1):
        libs = build()
        # "count": the op-counting build (float64 arithmetic + exact shortcuts, see count_real.h / tools/count_flops.py)
        self.lib = C.CDLL(libs[{"f64": 0, "f32": 1, "count": 2}[dtype]])
        assert self.lib.oracle_sizeof_model() == C.sizeof(OModel), "struct layout mismatch"
        self.mc = mc
        m = OModel()
        col = [g for g in range(mc.ngeom) if mc.geom_collides[g]]
        remap = {g: i for i, g in enumerate(col)}
        assert mc.nbody <= MAXB and mc.njnt <= MAXJ and mc.nv <= MAXV and len(col) <= MAXG
        assert len(mc.pair_geom) <= MAXP and mc.ncon <= MAXCON
        m.nq, m.nv, m.nbody, m.njnt, m.ngeom = mc.nq, mc.nv, mc.nbody, mc.njnt, len(col)
        m.npair, m.ncon = len(mc.pair_geom), mc.ncon
        m.iterations, m.ls_iterations = mc.opt["iterations"], mc.opt["ls_iterations"]
        sid = mc.site_id(tcp_site)
        m.tcp_body, m.hande_body = int(mc.site_body[sid]), mc.body_id(hande_body)
        m.timestep = timestep
        m.capbox_mode = int(capbox_mode)      # capsule_box restatement (mjstep.c): 1 = has_support gate, 0 = round-1
        m.tolerance, m.ls_tolerance = mc.opt["tolerance"], mc.opt["ls_tolerance"]
        m.impratio, m.meaninertia = mc.opt["impratio"], mc.meaninertia
        _fill(m.gravity, mc.opt["gravity"])
        _fill(m.tcp_pos, mc.site_pos[sid])
        _fill(m.body_parent, mc.body_parent)
        _fill(m.body_jnt, mc.body_jntadr)
        _fill(m.body_rootid, mc.body_rootid)
        _fill(m.body_weldid, mc.body_weldid)
        _fill(m.body_pos, mc.body_pos)
        _fill(m.body_quat, mc.body_quat)
        _fill(m.body_mass, mc.body_mass)
        _fill(m.body_ipos, mc.body_ipos)
        _fill(m.body_inertia, mc.body_inertia)
        _fill(m.body_gravcomp, mc.body_gravcomp)
        _fill(m.body_invweight0, mc.body_invweight0[:, 0])
        _fill(m.jnt_type, mc.jnt_type)
        _fill(m.jnt_body, mc.jnt_body)
        _fill(m.jnt_qposadr, mc.jnt_qposadr)
        _fill(m.jnt_dofadr, mc.jnt_dofadr)
        _fill(m.jnt_limited, mc.jnt_limited)
        _fill(m.jnt_axis, mc.jnt_axis)
        _fill(m.jnt_range, mc.jnt_range)
        _fill(m.jnt_armature, mc.jnt_armature)
        _fill(m.jnt_damping, mc.jnt_damping)
        _fill(m.jnt_solref, np.tile([0.02, 1.0], (mc.njnt, 1)))
        _fill(m.jnt_solimp, np.tile([0.9, 0.95, 0.001, 0.5, 2.0], (mc.njnt, 1)))
        _fill(m.jnt_margin, mc.jnt_margin)
        _fill(m.dof_invweight0, mc.dof_invweight0)
        _fill(m.geom_type, mc.geom_type[col])
        _fill(m.geom_body, mc.geom_body[col])
        _fill(m.geom_pos, mc.geom_pos[col])
        _fill(m.geom_quat, mc.geom_quat[col])
        _fill(m.geom_size, mc.geom_size[col])
        _fill(m.geom_friction, mc.geom_friction[col])
        _fill(m.geom_solref, mc.geom_solref[col])
        _fill(m.geom_solimp, mc.geom_solimp[col])
        _fill(m.pair_g1, [remap[g] for g in mc.pair_geom[:, 0]])
        _fill(m.pair_g2, [remap[g] for g in mc.pair_geom[:, 1]])
        _fill(m.pair_slotadr, mc.pair_slotadr)
        _fill(m.pair_nslot, mc.pair_nslot)
        self.mask = robot_slot_mask(mc)
        _fill(m.slot_robot, self.mask.astype(np.int32))
        self.m = m
        self.nrobot = int(self.mask.sum())

    # ------------------------------------------------------------------
    def forward(self, qpos, qvel, warm=None):
        mc = self.mc
        qpos = np.ascontiguousarray(qpos, dtype=np.float64)
        qvel = np.ascontiguousarray(qvel, dtype=np.float64)
        warm = np.zeros(mc.nv) if warm is None else np.ascontiguousarray(warm, dtype=np.float64)
        out = dict(qacc=np.zeros(mc.nv), M=np.zeros((mc.nv, mc.nv)), qfrc_bias=np.zeros(mc.nv),
                   qfrc_passive=np.zeros(mc.nv), con_dist=np.zeros(mc.ncon), xpos=np.zeros((mc.nbody, 3)),
                   xquat=np.zeros((mc.nbody, 4)), site_tcp=np.zeros(3), con_pos=np.zeros((mc.ncon, 3)),
                   con_frame=np.zeros((mc.ncon, 9)))
        nefc = C.c_int(0)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self.lib.oracle_forward(C.byref(self.m), p(qpos), p(qvel), p(warm), p(out["qacc"]), p(out["M"]),
                                p(out["qfrc_bias"]), p(out["qfrc_passive"]), p(out["con_dist"]), p(out["xpos"]),
                                p(out["xquat"]), p(out["site_tcp"]), p(out["con_pos"]), p(out["con_frame"]),
                                C.byref(nefc))
        out["nefc"] = nefc.value
        return out

    STAGES = ("kinematics", "com_pos + CRBA + factor", "velocity + passive + RNE + qacc_smooth", "narrow phase",
              "constraint rows", "Newton + line search", "Euler", "not counted (I/O, dense unobservable work)")

    def read_counts(self):
        """Op counts per stage since the last call (dtype="count" build only)."""
        out = (C.c_longlong * 8)()
        self.lib.oracle_count_read(out)
        return dict(zip(self.STAGES, [int(v) for v in out]))

    def initial_warmstart(self):
        """qacc of the constructor's ``mjx.forward`` at qpos0 (mjx_planner.py:107) = first warm start."""
        return self.forward(self.mc.qpos0, np.zeros(self.mc.nv))["qacc"]

    def rollout(self, thetadot, q0, v0, num_dof=6, warm=None, nthreads=0, want_collision=True, want_state=False):
        """compute_rollout_batch (mjx_planner.py:123,266-274). thetadot [B, num_dof*T]."""
        mc = self.mc
        thetadot = np.ascontiguousarray(thetadot, dtype=np.float64)
        B = thetadot.shape[0]
        T = thetadot.shape[1] // num_dof
        q0 = np.ascontiguousarray(q0, dtype=np.float64)
        v0 = np.ascontiguousarray(v0, dtype=np.float64)
        warm = self.initial_warmstart() if warm is None else np.ascontiguousarray(warm, dtype=np.float64)
        qpos_init = np.ascontiguousarray(mc.qpos0, dtype=np.float64)
        qvel_init = np.zeros(mc.nv)
        theta = np.zeros((B, num_dof * T))
        eef_pos = np.zeros((B, T, 3))
        eef_rot = np.zeros((B, T, 4))
        collision = np.zeros((B, T, self.nrobot)) if want_collision else None
        qpos_out = np.zeros((B, T, mc.nq)) if want_state else None
        qacc_out = np.zeros((B, T, mc.nv)) if want_state else None
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        self.lib.oracle_rollout(C.byref(self.m), B, T, num_dof, p(thetadot), p(q0), p(v0), p(qpos_init),
                                p(qvel_init), p(warm), p(theta), p(eef_pos), p(eef_rot), p(collision),
                                p(qpos_out), p(qacc_out), int(nthreads))
        if want_state:
            return theta, eef_pos, eef_rot, collision, qpos_out, qacc_out
        return theta, eef_pos, eef_rot, collision


def collide(kind, p1, m1, s1, p2, m2, s2, dtype="f64"):
    """Run one oracle collider in isolation (unit tests).  kind: plane_capsule, capsule_capsule,
    capsule_box, plane_box, box_box.  Returns dist [n], pos [n,3], frame [n,3,3]."""
    libs = build()
    lib = C.CDLL(libs[0] if dtype == "f64" else libs[1])
    code = ["plane_capsule", "capsule_capsule", "capsule_box", "plane_box", "box_box", "capsule_box_legacy"].index(kind)
    a = lambda x, n: np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1)[:n] if np.size(x) >= n
                                          else np.concatenate([np.asarray(x, dtype=np.float64).reshape(-1), np.zeros(n - np.size(x))]))
    P1, M1, S1, P2, M2, S2 = a(p1, 3), a(m1, 9), a(s1, 3), a(p2, 3), a(m2, 9), a(s2, 3)
    dist, pos, frame = np.zeros(4), np.zeros(12), np.zeros(36)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    n = lib.oracle_collide(code, p(P1), p(M1), p(S1), p(P2), p(M2), p(S2), p(dist), p(pos), p(frame))
    return dist[:n].copy(), pos.reshape(4, 3)[:n].copy(), frame.reshape(4, 3, 3)[:n].copy()
