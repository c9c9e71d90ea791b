"""The rollout kernel's source (csrc/rollout_core.h), compiled for the CPU by tests/emu, against the
oracle.  Catches indexing / physics mistakes in the kernel logic without a GPU; the real parity
tests (tests/test_gpu_*.py) run the CUDA build on the B200."""
import copy

import numpy as np
import pytest

from conftest import Q0, TARGET_POS, TARGET_ROT, planner_inputs
from emu_util import Emu


@pytest.fixture(scope="module")
def emu(mc, oracle64):
    from manipulator_mujoco_b200.kmodel import build_kmodel
    km, _ = build_kmodel(mc, 0.05, warm0=oracle64.initial_warmstart())
    return Emu(km)


def test_short_rollout_matches_oracle(emu, oracle64):
    pr, z, xi, st, xif, td = planner_inputs(16, 64)
    out = emu.rollout(td, Q0, np.zeros(6), TARGET_POS, TARGET_ROT)
    oth, oep, oer, ocol = oracle64.rollout(td, Q0, np.zeros(6))
    th, ot = out["theta"].reshape(64, 6, 16), oth.reshape(64, 6, 16)
    assert np.abs(th[:, :, :4] - ot[:, :, :4]).max() < 2e-6           # before the box lands: well conditioned
    assert np.abs(out["eef_pos"][:, :5] - oep[:, :5]).max() < 2e-6
    assert np.abs(out["theta"] - oth).max() < 5e-4                    # after: line-search branch noise
    assert (np.abs(out["collision"] - ocol) > 1e-3).mean() < 1e-4
    oc = pr.compute_cost_batch(oep, oer, ocol, TARGET_POS, TARGET_ROT)
    np.testing.assert_allclose(out["cost4"][:, 1], oc[1], rtol=2e-4)
    np.testing.assert_allclose(out["cost4"][:, 2], oc[2], rtol=2e-4, atol=1e-4)
    np.testing.assert_allclose(out["cost4"][:, 0], 20 * out["cost4"][:, 1] + 3 * out["cost4"][:, 2] + 80 * out["cost4"][:, 3], rtol=1e-5)
    assert out["flags"].max() == 0


def test_teacher_forced_steps_on_contact_states(emu, oracle64, oracle32):
    """Single steps from states of a long oracle rollout where the robot is in contact.
    Narrow phase must agree to rounding.  The solver output qacc must agree wherever the reference
    algorithm itself is precision-stable (float32 and float64 builds of the oracle agree)."""
    T, B = 100, 96
    pr, z, xi, st, xif, td = planner_inputs(T, B, seed=1)
    oth, oep, oer, ocol, oqp, oqa = oracle64.rollout(td, Q0, np.zeros(6), want_state=True)
    has = np.where((ocol < 0).any(axis=(1, 2)))[0]
    assert len(has) >= 10
    warm0 = oracle64.initial_warmstart()
    n = stable = 0
    worst_col = 0.0
    dev = []
    for s in has[:24]:
        qvbox = np.zeros(6)
        for t in range(T):
            qpos = np.concatenate([Q0, oracle64.mc.qpos0[6:]]) if t == 0 else oqp[s, t - 1]
            warm = warm0 if t == 0 else oqa[s, t - 1]
            qvel = np.zeros(12)
            qvel[6:] = qvbox
            qvel[:6] = td[s].reshape(6, T)[:, t]
            qvbox = qvbox + 0.05 * oqa[s, t, 6:]
            if not (ocol[s, t] < 0).any() or t % 3:
                continue
            r64, r32 = oracle64.forward(qpos, qvel, warm), oracle32.forward(qpos, qvel, warm)
            km = copy.copy(emu.km)
            for i in range(13):
                km.qpos0[i] = qpos[i]
            for i in range(12):
                km.warm0[i], km.qvel0[i] = warm[i], qvel[i]
            out = emu.rollout(qvel[:6].reshape(1, 6), qpos[:6], qvel[:6], TARGET_POS, TARGET_ROT, km=km)
            worst_col = max(worst_col, np.abs(out["collision"][0, 0] - r64["con_dist"][oracle64.mask]).max())
            scale = max(1.0, np.abs(r64["qacc"]).max())
            n += 1
            if np.abs(r32["qacc"] - r64["qacc"]).max() < 1e-3 * scale:
                stable += 1
                dev.append(np.abs(out["qacc"][0, 0] - r64["qacc"]).max() / scale)
    assert n > 50 and stable > 0.3 * n
    assert worst_col < 1e-5
    dev = np.array(dev)
    assert (dev < 5e-2).all(), np.sort(dev)[-5:]                  # inside the line-search bracket-flip jump
    assert (dev < 1e-3).mean() >= 0.85, np.sort(dev)[-10:]


def test_contact_capacity_variants_agree(emu):
    """NC = 20 (fast kernel) and NC = 48 (debug instantiation) are the same code: identical results
    whenever nothing overflows."""
    pr, z, xi, st, xif, td = planner_inputs(30, 32, seed=2)
    a = emu.rollout(td, Q0, np.zeros(6), TARGET_POS, TARGET_ROT, nc=20)
    b = emu.rollout(td, Q0, np.zeros(6), TARGET_POS, TARGET_ROT, nc=48)
    assert a["flags"].max() == 0
    np.testing.assert_array_equal(a["theta"], b["theta"])
    np.testing.assert_array_equal(a["cost4"], b["cost4"])


def test_robot_box_coupled_contact(emu, oracle64, oracle32, mc):
    """States where a robot capsule touches the free box (target_0): the Hessian couples the robot
    and box blocks (full 12x12 Cholesky path of the kernel)."""
    rng = np.random.default_rng(5)
    names = mc.geom_names
    slots = []
    for (g1, g2), a, n in zip(mc.pair_geom, mc.pair_slotadr, mc.pair_nslot):
        if names[g2] == "target_0" and (names[g1] or "").startswith("robot_"):
            slots += list(range(a, a + n))
    found = checked = 0
    dev = []
    for _ in range(200):
        qr = Q0 + rng.normal(size=6) * 0.3
        tcp = oracle64.forward(np.concatenate([qr, mc.qpos0[6:]]), np.zeros(12))["site_tcp"]
        qbox = np.concatenate([tcp + rng.normal(size=3) * 0.01, [1.0, 0, 0, 0]])     # box floating at the tool tip
        q = np.concatenate([qr, qbox])
        v = np.zeros(12)
        v[:6] = rng.normal(size=6) * 0.2
        r64 = oracle64.forward(q, v)
        if not (r64["con_dist"][slots] < 0).any():
            continue
        found += 1
        r32 = oracle32.forward(q, v)
        km = copy.copy(emu.km)
        for i in range(13):
            km.qpos0[i] = q[i]
        for i in range(12):
            km.warm0[i], km.qvel0[i] = 0.0, v[i]
        out = emu.rollout(v[:6].reshape(1, 6), q[:6], v[:6], TARGET_POS, TARGET_ROT, km=km)
        np.testing.assert_allclose(out["collision"][0, 0], r64["con_dist"][oracle64.mask], atol=1e-5)
        scale = max(1.0, np.abs(r64["qacc"]).max())
        if np.abs(r32["qacc"] - r64["qacc"]).max() < 1e-3 * scale:
            checked += 1
            dev.append(np.abs(out["qacc"][0, 0] - r64["qacc"]).max() / scale)
        if found >= 12:
            break
    assert found >= 5 and checked >= 2
    # The line search accepts a bracket point on a rounding-level comparison: a state on
    # which the oracle's own float32 build happens to agree with float64 can still flip for another float32
    # evaluation order, which moves qacc by ~1e-2 of its scale.  All states stay inside that jump, most match closely.
    dev = np.array(dev)
    assert (dev < 5e-2).all(), dev
    assert (dev < 1e-3).mean() >= 0.7, dev


def test_contact_spill_area_is_bit_identical_to_all_shared_memory(emu, oracle64, mc):
    """More simultaneous contacts than the fast kernel keeps in shared memory (20): the extra ones live in the
    per-sample global spill area.  Same code, same arithmetic -- results must equal the 48-contact
    all-in-shared-memory instantiation bit for bit, on single steps and over a horizon."""
    rng = np.random.default_rng(3)
    deep = []
    for _ in range(4000):
        q = rng.uniform(-3, 3, 6)
        r = oracle64.forward(np.concatenate([q, mc.qpos0[6:]]), np.zeros(12))
        n = int((r["con_dist"] < 0).sum())
        if n > 20:
            deep.append((n, q))
        if len(deep) >= 3:
            break
    assert deep, "no state with more than 20 contacts found"
    T = 6
    for n, q in deep:
        td = rng.normal(size=(1, 6 * T)) * 0.3
        a = emu.rollout(td, q, np.zeros(6), TARGET_POS, TARGET_ROT, nc=20)
        b = emu.rollout(td, q, np.zeros(6), TARGET_POS, TARGET_ROT, nc=48)
        assert a["flags"][0] == 0 and b["flags"][0] == 0
        for key in ("theta", "cost4", "qacc", "collision", "eef_pos"):
            np.testing.assert_array_equal(a[key].view(np.int32), b[key].view(np.int32), err_msg=f"{key} ({n} contacts)")
        # and the spilled contacts take part in the solve: the oracle agrees on the first step's distances
        r64 = oracle64.forward(np.concatenate([q, mc.qpos0[6:]]), np.concatenate([td[0, ::T], np.zeros(6)]))
        np.testing.assert_allclose(a["collision"][0, 0], r64["con_dist"][oracle64.mask], atol=2e-5)


def test_free_box_contacts_in_odd_poses_match_oracle(emu, oracle64, oracle32, mc):
    """The free box pushed into / over the edges of the static boxes in random orientations: exercises the
    cooperative box-box narrow phase beyond the resting face contact (edge-edge, clipped incident faces,
    more than four polygon vertices).  The box's own acceleration must match the oracle's wherever the solve
    is precision-stable; the active-contact count is cross-checked through the constraint force."""
    rng = np.random.default_rng(12)
    names = mc.geom_names
    bslots = []
    for (g1, g2), a, n in zip(mc.pair_geom, mc.pair_slotadr, mc.pair_nslot):
        if "target_0" in (names[g1], names[g2]) and not (names[g1] or "").startswith("robot_") and not (names[g2] or "").startswith("robot_"):
            bslots += list(range(a, a + n))
    anchors = [np.array([0.546, 0.0, 0.425]), np.array([-0.2, -0.15, 0.5]), np.array([-0.3, 0.3, 0.5]), np.array([0.0, 0.625, 0.425]),
               np.array([0.3, 0.3, 0.44]), np.array([0.0, 0.0, 0.445])]
    seen = {}
    checked = 0
    dev = []
    for trial in range(400):
        c = anchors[trial % len(anchors)] + rng.normal(size=3) * np.array([0.03, 0.03, 0.02])
        quat = rng.normal(size=4); quat /= np.linalg.norm(quat)
        if trial % 3 == 0:
            quat = np.array([1.0, 0, 0, 0]) + rng.normal(size=4) * 0.05; quat /= np.linalg.norm(quat)
        q = np.concatenate([Q0, c, quat])
        v = np.zeros(12); v[6:] = rng.normal(size=6) * 0.1
        r64 = oracle64.forward(q, v)
        nact = int((r64["con_dist"][bslots] < 0).sum())
        if nact == 0:
            continue
        seen[nact] = seen.get(nact, 0) + 1
        r32 = oracle32.forward(q, v)
        km = copy.copy(emu.km)
        for i in range(13):
            km.qpos0[i] = q[i]
        for i in range(12):
            km.warm0[i], km.qvel0[i] = 0.0, v[i]
        out = emu.rollout(np.zeros((1, 6)), Q0, np.zeros(6), TARGET_POS, TARGET_ROT, km=km)
        scale = max(1.0, np.abs(r64["qacc"]).max())
        if np.abs(r32["qacc"] - r64["qacc"]).max() < 1e-3 * scale:
            checked += 1
            dev.append(np.abs(out["qacc"][0, 0] - r64["qacc"]).max() / scale)
    dev = np.array(dev)
    print("active free-box contacts -> cases:", dict(sorted(seen.items())), "checked", checked, "max dev", dev.max())
    assert checked >= 40 and len(seen) >= 3, (checked, seen)          # 1 (edge-edge), 2-3 (clipped), 4 (face) contacts all occur
    assert (dev < 5e-2).all(), np.sort(dev)[-5:]
    assert (dev < 1e-3).mean() >= 0.85, np.sort(dev)[-10:]


def test_spill_area_with_robot_box_coupling(emu, oracle64, mc):
    """Deep in the table *and* touching the free box: the spill-capable instantiation of the solve together with
    the coupled 12x12 system.  Bit-identical to the all-in-shared-memory instantiation over a short horizon."""
    rng = np.random.default_rng(8)
    hits = 0
    for q in ([1.181, 0.789, -0.949, -0.051, 2.391, -0.535], [1.168, 0.758, -0.921, -0.796, 0.168, 2.145]):
        q = np.array(q)
        tcp = oracle64.forward(np.concatenate([q, mc.qpos0[6:]]), np.zeros(12))["site_tcp"]
        for _ in range(6):
            qbox = np.concatenate([tcp + rng.normal(size=3) * 0.015, [1.0, 0, 0, 0]])
            r = oracle64.forward(np.concatenate([q, qbox]), np.zeros(12))
            nact = int((r["con_dist"] < 0).sum())
            if nact <= 20:
                continue
            km = copy.copy(emu.km)
            for i in range(13):
                km.qpos0[i] = np.concatenate([q, qbox])[i]
            for i in range(12):
                km.warm0[i], km.qvel0[i] = 0.0, 0.0
            T = 5
            td = rng.normal(size=(1, 6 * T)) * 0.2
            a = emu.rollout(td, q, np.zeros(6), TARGET_POS, TARGET_ROT, nc=20, km=km)
            b = emu.rollout(td, q, np.zeros(6), TARGET_POS, TARGET_ROT, nc=48, km=km)
            for key in ("theta", "cost4", "qacc", "collision"):
                np.testing.assert_array_equal(a[key].view(np.int32), b[key].view(np.int32), err_msg=f"{key} ({nact} contacts)")
            np.testing.assert_allclose(a["collision"][0, 0], r["con_dist"][oracle64.mask], atol=2e-5)
            hits += 1
    assert hits >= 2


def test_kernel_capsule_box_matches_oracle_fuzz():
    """The kernel's capsule-box collider (far-field early-out, closed-form clip, filtered edge loop) against the
    oracle's literal restatement of MJX _capsule_convex on random poses concentrated around contact: distances of
    both slots (incl. the +1 sentinel), and position / normal of every penetrating slot."""
    from emu_util import capsule_box
    from oracle.oracle import collide
    from test_oracle_colliders import rot
    rng = np.random.default_rng(7)
    n_near = n_act = n_edge = n_flip = 0
    dev_p, dev_n = [], []
    for trial in range(6000):
        bs = rng.uniform(0.02, 0.3, 3)
        r, hl = rng.uniform(0.02, 0.06), rng.uniform(0.02, 0.2)
        Rb = rot(rng.normal(size=3), rng.uniform(0, 3)) if trial % 2 else np.eye(3)
        Rc = rot(rng.normal(size=3), rng.uniform(0, 3))
        # capsule centre near the box surface: pick a point on a random face / edge / corner shell
        p = rng.uniform(-1, 1, 3) * bs
        k = rng.integers(3)
        p[k] = np.sign(p[k] or 1.0) * bs[k]
        if trial % 3 == 0:
            k2 = (k + 1) % 3
            p[k2] = np.sign(p[k2] or 1.0) * bs[k2]
        off = rng.normal(size=3)
        c_local = p + off / np.linalg.norm(off) * rng.uniform(0, 1.5) * (r + hl * rng.uniform(0, 1))
        bpos = rng.uniform(-0.5, 0.5, 3)
        cpos = bpos + Rb @ c_local
        d_o, p_o, f_o = collide("capsule_box", cpos, Rc, [r, hl, 0], bpos, Rb, bs)
        # position / normal of an edge contact come out of MJX's regularised closest-segment routine, which is
        # float32-sensitive (~1e-3 on the normal for near-parallel segments): compare those in the same precision
        _, p_o, f_o = collide("capsule_box", cpos, Rc, [r, hl, 0], bpos, Rb, bs, dtype="f32")
        d_k, p_k, n_k = capsule_box(cpos, Rc, [r, hl, 0], bpos, Rb, bs)
        sent_o, sent_k = d_o == 1, d_k == 1
        if (sent_o != sent_k).any():
            # the +1 / real-distance switch is a genuine discontinuity of MJX's collider (support == 0 or
            # edge penetration == 0): float32 may sit on the other side only within rounding of it
            real = np.where(sent_o, d_k, d_o)[sent_o != sent_k]
            assert np.abs(real).max() < 1e-5, (trial, d_o, d_k)
            n_flip += 1
            continue
        n_near += (~sent_o).any()
        np.testing.assert_allclose(d_k, d_o, atol=2e-6, err_msg=str(trial))
        for j in range(2):
            if d_o[j] < -1e-4:
                n_act += 1
                n_edge += abs(abs(f_o[j][0] @ Rb[:, 0]) - 1) > 1e-6 and abs(abs(f_o[j][0] @ Rb[:, 1]) - 1) > 1e-6 and abs(abs(f_o[j][0] @ Rb[:, 2]) - 1) > 1e-6
                dev_p.append(np.abs(p_k[j] - p_o[j]).max())
                dev_n.append(np.abs(n_k[j] - f_o[j][0]).max())
    dev_p, dev_n = np.array(dev_p), np.array(dev_n)
    assert dev_p.max() < 2e-4 and np.percentile(dev_p, 99) < 2e-5, (dev_p.max(), np.percentile(dev_p, 99))
    assert dev_n.max() < 5e-3 and np.percentile(dev_n, 99) < 2e-4, (dev_n.max(), np.percentile(dev_n, 99))
    assert n_near > 1500 and n_act > 1000 and n_edge > 100 and n_flip < 10, (n_near, n_act, n_edge, n_flip)


def test_lane_parallel_edge_stage_equals_the_serial_one():
    """The near pass evaluates the 12 box edges of a pair on 12 lanes (capbox_edge_lane, written in the edge's own coordinates
    with component selects) and reduces to the first maximum; capsule_box<true> -- the version fuzzed against the oracle in
    tests/test_oracle_colliders.py -- walks them serially (capbox_edges).  Same contact on random near poses, edge hits included."""
    import ctypes as C
    from emu_util import build
    lib = C.CDLL(build())
    rng = np.random.default_rng(8)
    n_edge = 0
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for _ in range(4000):
        size = f(rng.uniform(0.02, 0.3, 3))
        r = float(rng.uniform(0.02, 0.06))
        # segments hovering around a box edge / corner region, so the edge stage runs and often wins
        c = rng.uniform(-1, 1, 3) * (size + 1.2 * r)
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        hl = rng.uniform(0.02, 0.2)
        a, b = f(c - d * hl), f(c + d * hl)
        outs = [np.zeros(2, np.float32), np.zeros(6, np.float32), np.zeros(6, np.float32), np.zeros(2, np.float32), np.zeros(6, np.float32), np.zeros(6, np.float32)]
        lib.emu_capsule_box_lanes(p(a), p(b), C.c_float(r), p(size), *[p(o) for o in outs])
        d2, p6, n6, d2s, p6s, n6s = outs
        np.testing.assert_allclose(d2, d2s, atol=2e-6)
        if d2s[0] < 0 or d2s[1] < 0:
            np.testing.assert_allclose(p6, p6s, atol=1e-4)      # (the foot point on a near-parallel edge moves with the summation order)
            np.testing.assert_allclose(n6, n6s, atol=2e-3)
        n_edge += int(abs(n6s[:3]).max() < 0.999)              # slot 0 carries an edge contact (normal not a face normal)
    assert n_edge > 100
