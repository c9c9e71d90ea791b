"""Build + call the CPU emulation build of the rollout core (tests/emu/emu.cpp).  Test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "_emu_build", "libcemk_emu.so")
SRCS = [os.path.join(ROOT, "tests", "emu", "emu.cpp")] + [os.path.join(ROOT, "manipulator_mujoco_b200", "csrc", n)
                                                          for n in ("rollout_core.h", "warp_dsl.h", "kmodel.h")]


def build(race=False):
    """race=True: the race-checking emulation (-DCEMK_EMU_RACE, csrc/warp_dsl.h) as a second library."""
    out = OUT.replace(".so", "_race.so") if race else OUT
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in SRCS):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-fPIC", "-shared", "-fopenmp", "-std=c++17", "-Wno-unknown-pragmas"] + (["-DCEMK_EMU_RACE"] if race else [])
                   + ["-o", out, SRCS[0]], check=True, capture_output=True)
    return out


class Emu:
    def __init__(self, km, race=False):
        from manipulator_mujoco_b200.kmodel import KModel
        self.lib = C.CDLL(build(race))
        assert self.lib.emu_race_enabled() == int(race)
        assert self.lib.emu_sizeof_kmodel() == C.sizeof(KModel)
        self.km = km

    def rollout(self, td, q0, v0, tpos, trot, w=(20.0, 3.0, 80.0), nc=48, km=None):
        km = km or self.km
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        td = f(td)
        B, T = td.shape[0], td.shape[1] // 6
        q0, v0, tpos, trot = f(q0), f(v0), f(tpos), f(trot)
        out = dict(theta=np.zeros((B, 6 * T), np.float32), cost4=np.zeros((B, 4), np.float32), eef_pos=np.zeros((B, T, 3), np.float32),
                   eef_rot=np.zeros((B, T, 4), np.float32), collision=np.zeros((B, T, km.nslot_robot), np.float32),
                   qacc=np.zeros((B, T, 12), np.float32), flags=np.zeros(B, np.int32))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self.lib.emu_rollout(C.byref(km), B, T, p(td), p(q0), p(v0), p(tpos), p(trot), C.c_float(w[0]), C.c_float(w[1]), C.c_float(w[2]),
                             p(out["theta"]), p(out["cost4"]), p(out["eef_pos"]), p(out["eef_rot"]), p(out["collision"]), p(out["qacc"]),
                             p(out["flags"]), int(nc))
        return out


def capsule_box(cpos, cmat, csize, bpos, bmat, bsize):
    """The kernel's capsule_box<true> (csrc/rollout_core.h) on one pair -> dist [2], pos [2,3], normal [2,3]."""
    lib = C.CDLL(build())
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    cpos, cmat = np.asarray(cpos, float), np.asarray(cmat, float).reshape(3, 3)
    A, B = f(cpos - cmat[:, 2] * csize[1]), f(cpos + cmat[:, 2] * csize[1])
    bp, bm, bs = f(bpos), f(np.asarray(bmat, float).reshape(-1)), f(bsize)
    d, p, n = np.zeros(2, np.float32), np.zeros(6, np.float32), np.zeros(6, np.float32)
    q = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.emu_capsule_box(q(A), q(B), C.c_float(csize[0]), q(bp), q(bm), q(bs), q(d), q(p), q(n))
    return d, p.reshape(2, 3), n.reshape(2, 3)


def race_stats(emu):
    """(write-write conflicts, lane blocks checked) since the last call; race build only."""
    out = (C.c_longlong * 5)()
    emu.lib.emu_race_stats(out)
    emu.first_conflict = dict(line=int(out[2]), word=int(out[3]), lanes=(int(out[4]) // 100, int(out[4]) % 100)) if out[0] else None
    return int(out[0]), int(out[1])


def race_selftest(which, race):
    lib = C.CDLL(build(race))
    out = np.zeros(33, np.float32)
    lib.emu_race_selftest(int(which), out.ctypes.data_as(C.c_void_p))
    return out[:16].copy(), int(out[16])
