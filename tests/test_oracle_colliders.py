"""Colliders of the oracle in isolation, against brute-force geometry where the MJX semantics
coincide with true signed distance, and the documented sentinel / slot semantics elsewhere."""
import numpy as np
import pytest

from oracle.oracle import collide

I3 = np.eye(3)


def rot(axis, ang):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return I3 + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def seg_points(c, R, hl, n=400):
    return c[None] + np.linspace(-hl, hl, n)[:, None] * R[:, 2][None]


def test_capsule_capsule_vs_bruteforce():
    rng = np.random.default_rng(0)
    for _ in range(50):
        c1, c2 = rng.uniform(-.5, .5, 3), rng.uniform(-.5, .5, 3)
        R1, R2 = rot(rng.normal(size=3), rng.uniform(0, 3)), rot(rng.normal(size=3), rng.uniform(0, 3))
        s1, s2 = [0.05, rng.uniform(.02, .3), 0], [0.04, rng.uniform(.02, .3), 0]
        d, pos, fr = collide("capsule_capsule", c1, R1, s1, c2, R2, s2)
        a, b = seg_points(c1, R1, s1[1]), seg_points(c2, R2, s2[1])
        brute = np.sqrt(((a[:, None] - b[None]) ** 2).sum(-1)).min() - s1[0] - s2[0]
        assert abs(d[0] - brute) < 2e-3
        assert abs(np.linalg.norm(fr[0][0]) - 1) < 1e-9
        np.testing.assert_allclose(fr[0] @ fr[0].T, I3, atol=1e-9)


def test_plane_capsule():
    c, R = np.array([0.1, 0.2, 0.3]), rot([1, 0, 0], 1.2)
    d, pos, fr = collide("plane_capsule", [0, 0, 0], I3, [0, 0, 0.05], c, R, [0.06, 0.2, 0])
    ends = [c + R[:, 2] * 0.2, c - R[:, 2] * 0.2]
    np.testing.assert_allclose(d, [e[2] - 0.06 for e in ends], atol=1e-12)
    np.testing.assert_allclose(fr[0][0], [0, 0, 1], atol=1e-12)
    # first tangent follows the capsule axis projected on the plane
    t = R[:, 2] - R[2, 2] * np.array([0, 0, 1.0])
    np.testing.assert_allclose(fr[0][1], t / np.linalg.norm(t), atol=1e-12)
    # ... unless the capsule is within 30 degrees of the plane normal (|projection| < 0.5): world y
    d2, _, fr2 = collide("plane_capsule", [0, 0, 0], I3, [0, 0, 0.05], c, rot([1, 0, 0], 0.4), [0.06, 0.2, 0])
    np.testing.assert_allclose(fr2[0][1], [0, 1, 0], atol=1e-12)


def test_capsule_box_far_field_is_the_sentinel():
    """has_support gate (MJX _capsule_convex): a segment hovering over a face with the face plane separating the
    radius-inflated capsule from the box reports dist = +1 in both slots -- no far-field distance.  The round-1
    restatement (kept as "capsule_box_legacy") reported height - radius there."""
    box = [0.3, 0.2, 0.1]
    rng = np.random.default_rng(1)
    for _ in range(50):
        c = np.array([rng.uniform(-.2, .2), rng.uniform(-.1, .1), rng.uniform(0.15, 0.6)])
        R = rot([0, 1, 0], np.pi / 2 + rng.uniform(-.2, .2))           # roughly horizontal capsule
        hl = 0.05
        ends = [c - R[:, 2] * hl, c + R[:, 2] * hl]
        assert max(abs(e[0]) for e in ends) < 0.3 and max(abs(e[1]) for e in ends) < 0.2
        assert min(e[2] for e in ends) - 0.1 - 0.03 > 0
        d, pos, fr = collide("capsule_box", c, R, [0.03, hl, 0], [0, 0, 0], I3, box)
        np.testing.assert_allclose(d, [1, 1])
        d0, _, fr0 = collide("capsule_box_legacy", c, R, [0.03, hl, 0], [0, 0, 0], I3, box)
        np.testing.assert_allclose(d0, [e[2] - 0.1 - 0.03 for e in ends], atol=1e-12)


def test_capsule_box_face_contact_is_true_distance():
    """Penetrating the top face (every face plane has an inflated end point behind it): both slots measure
    (height - radius) of the clipped end points, normal capsule -> box."""
    box = [0.3, 0.2, 0.1]
    rng = np.random.default_rng(2)
    for _ in range(50):
        c = np.array([rng.uniform(-.2, .2), rng.uniform(-.1, .1), rng.uniform(0.105, 0.125)])
        R = rot([0, 1, 0], np.pi / 2 + rng.uniform(-.05, .05))
        hl = 0.05
        ends = [c - R[:, 2] * hl, c + R[:, 2] * hl]
        if min(e[2] for e in ends) - 0.1 - 0.03 >= 0:
            continue
        d, pos, fr = collide("capsule_box", c, R, [0.03, hl, 0], [0, 0, 0], I3, box)
        np.testing.assert_allclose(d, [e[2] - 0.1 - 0.03 for e in ends], atol=1e-12)
        np.testing.assert_allclose(fr[0][0], [0, 0, -1], atol=1e-12)   # capsule -> box
        np.testing.assert_allclose(pos[0][:2], ends[0][:2], atol=1e-12)


def test_capsule_box_one_end_in_contact_reports_the_other_ends_distance():
    """has_support only needs ONE inflated end point behind each face plane: a tilted capsule touching the top
    face with its lower end reports the (positive) face distance of its upper end in the other slot."""
    R = rot([0, 1, 0], np.pi / 2 - 0.5)
    c = np.array([0.0, 0.0, 0.14])
    hl, r = 0.05, 0.03
    ends = [c - R[:, 2] * hl, c + R[:, 2] * hl]
    d, _, _ = collide("capsule_box", c, R, [r, hl, 0], [0, 0, 0], I3, [0.3, 0.2, 0.1])
    np.testing.assert_allclose(d, [e[2] - 0.1 - r for e in ends], atol=1e-12)
    assert d.min() < 0 < d.max()


def test_capsule_box_sentinel_when_clip_fails():
    """Whole segment outside one side plane of the best face: penetration -1 => dist = +1."""
    d, _, _ = collide("capsule_box", [1.0, 0, 0.5], rot([0, 1, 0], np.pi / 2), [0.03, 0.05, 0], [0, 0, 0], I3, [0.1, 0.1, 0.1])
    np.testing.assert_allclose(d, [1, 1])


def test_capsule_box_edge_contact_replaces_slot0():
    # vertical capsule just outside the +x side of the top face, touching the top edge: the capsule point lies in
    # front of both faces adjacent to that edge (+x and +z), so it is a shallow edge contact although no face has support
    d, pos, fr = collide("capsule_box", [0.12, 0, 0.16], I3, [0.05, 0.05, 0], [0, 0, 0], I3, [0.1, 0.1, 0.1])
    # closest point of the segment (its lower end at z=.11) to the edge x=.1,z=.1: distance sqrt(.02^2+.01^2)
    assert abs(d[0] - (np.hypot(0.02, 0.01) - 0.05)) < 1e-9
    assert d[0] < 0 and d[1] == 1
    n = np.array([0.1 - 0.12, 0, 0.1 - 0.11]); n /= np.linalg.norm(n)
    np.testing.assert_allclose(fr[0][0], n, atol=1e-9)               # capsule -> edge
    np.testing.assert_allclose(pos[0], 0.5 * (np.array([0.1, 0, 0.1]) + np.array([0.12, 0, 0.11]) + n * 0.05), atol=1e-9)


def test_capsule_box_edge_needs_the_voronoi_region():
    # same capsule moved over the face (x < .1): its closest point is not in front of the +x face, the edge does not
    # qualify; the face contact (has_support holds) is reported instead
    d, pos, fr = collide("capsule_box", [0.095, 0, 0.16], I3, [0.05, 0.05, 0], [0, 0, 0], I3, [0.1, 0.1, 0.1])
    np.testing.assert_allclose(d[0], 0.11 - 0.1 - 0.05, atol=1e-12)
    np.testing.assert_allclose(fr[0][0], [0, 0, -1], atol=1e-12)


def test_box_box_resting_face_contact():
    d, pos, fr = collide("box_box", [0, 0, 0], I3, [0.546, 0.625, 0.025], [0.1, 0.1, 0.04], rot([0, 0, 1], 0.3), [0.02, 0.02, 0.02])
    np.testing.assert_allclose(d, -0.005, atol=1e-12)
    np.testing.assert_allclose(fr[0][0], [0, 0, 1], atol=1e-12)
    assert len({tuple(np.round(p, 6)) for p in pos}) == 4              # four distinct corners
    np.testing.assert_allclose(pos[:, 2], 0.02, atol=1e-12)            # on the reference face (ties prefer the axes of geom2)


def test_box_box_separated_is_inactive():
    d, _, _ = collide("box_box", [0, 0, 0], I3, [0.1, 0.1, 0.1], [0.5, 0, 0], rot([1, 1, 0], 0.7), [0.05, 0.05, 0.05])
    assert (d > 0).all()


def test_plane_box_penetrating_vertices():
    d, pos, fr = collide("plane_box", [0, 0, 0], I3, [0, 0, 0.05], [0, 0, 0.015], I3, [0.02, 0.02, 0.02])
    np.testing.assert_allclose(np.sort(d), [-0.005] * 4, atol=1e-12)
