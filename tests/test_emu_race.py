"""Race check of the rollout core's lane blocks without a GPU (compute-sanitizer's racecheck is not available on this pool).

The kernel is written against csrc/warp_dsl.h, whose rule 1 says: inside one LANES block a lane reads only scratch written
in earlier blocks (or by itself) and writes only words no other lane touches.  The race-checking emulation
(-DCEMK_EMU_RACE) enforces exactly that: every lane of a block runs against the scratch as it was when the block started,
the writes are merged when it ends, and two lanes writing different values to one word are counted.  If the kernel obeys
the rule, this mode and the plain lane-after-lane emulation agree bit for bit; a missing fence makes them differ."""
import numpy as np
import pytest

from conftest import Q0, TARGET_POS, TARGET_ROT, planner_inputs
from emu_util import Emu, race_selftest, race_stats


@pytest.fixture(scope="module")
def emus(mc, oracle64):
    from manipulator_mujoco_b200.kmodel import build_kmodel
    km, _ = build_kmodel(mc, 0.05, warm0=oracle64.initial_warmstart())
    return Emu(km), Emu(km, race=True)


def test_checker_sees_a_misfenced_block_and_a_double_write():
    ok_plain, _ = race_selftest(0, False)
    ok_race, waw0 = race_selftest(0, True)
    assert np.array_equal(ok_plain, ok_race) and waw0 == 0            # correctly fenced: both modes agree
    bad_plain, _ = race_selftest(1, False)
    bad_race, _ = race_selftest(1, True)
    assert not np.array_equal(bad_plain, bad_race)                    # same-block read of a neighbour's write: detected
    _, waw = race_selftest(2, True)
    assert waw == 1                                                   # two lanes, one word, different values: counted


def _same(a, b):
    for k in a:
        assert np.array_equal(a[k].view(np.int32) if a[k].dtype == np.float32 else a[k], b[k].view(np.int32) if b[k].dtype == np.float32 else b[k]), k


def test_planner_rollouts_obey_the_lane_block_rules(emus, oracle64):
    """Contact-rich planner samples over a long horizon (robot contacts, the box landing, coupled solves)."""
    plain, race = emus
    T, B = 60, 96
    _, _, _, _, _, td = planner_inputs(T, B, seed=1)
    _, _, _, ocol = oracle64.rollout(td, Q0, np.zeros(6))
    has = np.where((ocol < 0).any(axis=(1, 2)))[0]
    sel = np.concatenate([has[:12], np.setdiff1d(np.arange(B), has)[:4]])
    assert len(has) >= 6
    a = plain.rollout(td[sel], Q0, np.zeros(6), TARGET_POS, TARGET_ROT, nc=20)
    race_stats(race)
    b = race.rollout(td[sel], Q0, np.zeros(6), TARGET_POS, TARGET_ROT, nc=20)
    waw, blocks = race_stats(race)
    assert blocks > 1000 * len(sel) and waw == 0
    _same(a, b)


def test_deep_collision_states_obey_the_rules_in_both_instantiations(emus, mc):
    """Start states inside the table / obstacles: dozens of contacts, the spill area (fast instantiation, NC = 20) and
    the all-in-shared-memory instantiation (NC = 48), robot-box coupling."""
    plain, race = emus
    rng = np.random.default_rng(11)
    T = 6
    for q0 in (np.array([1.5, -0.6, 1.9, -1.25, -1.6, 0.0]), np.array([2.3, -1.0, 1.9, -2.4, -1.6, 0.3]), np.array([-2.6, -0.75, 1.6, -2.4, -1.6, 0.0])):
        td = rng.uniform(-0.6, 0.6, size=(4, 6 * T))
        for nc in (20, 48):
            a = plain.rollout(td, q0, np.zeros(6), TARGET_POS, TARGET_ROT, nc=nc)
            race_stats(race)
            b = race.rollout(td, q0, np.zeros(6), TARGET_POS, TARGET_ROT, nc=nc)
            waw, blocks = race_stats(race)
            assert waw == 0 and blocks > 0
            _same(a, b)
