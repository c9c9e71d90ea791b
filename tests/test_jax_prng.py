"""jax.random restatement: host key arithmetic, the numpy oracle of the normal
draws and -- on the GPU -- the device generator behind compute_xi_samples (mjx_planner.py:313-316).

Pins (jax itself is not installable here):
 * Threefry-2x32-20 known answers of the Random123 distribution (also jax's own tests/random_test.py);
 * outputs printed in the JAX documentation ("Pseudorandom numbers" tutorial, PRNGKey(42); jax.random.split
   docstring era, PRNGKey(0)) for the original counter layout.
The partitionable layout (default since jax 0.5, the reference pins 0.5.3) has no published vector that
could be recalled; it differs from the original one only in the counters fed to the same block function.
"""
import ctypes as C

import numpy as np
import pytest

from manipulator_mujoco_b200 import jax_prng
from oracle import jax_random_ref as ref


@pytest.mark.parametrize("key, ctr, out", [
    ((0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6b200159, 0x99ba4efe)),
    ((0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x1cb996fc, 0xbb002be7)),
    ((0x13198a2e, 0x03707344), (0x243f6a88, 0x85a308d3), (0xc4923a9c, 0x483df7a0)),
])
def test_threefry_known_answers(key, ctr, out):
    assert ref.threefry2x32_scalar(*key, *ctr) == out
    a, b = jax_prng.threefry2x32(key, [ctr[0]], [ctr[1]])
    assert (int(a[0]), int(b[0])) == out


def test_key_and_split_original_layout_match_jax_docs():
    np.testing.assert_array_equal(jax_prng.PRNGKey(0), [0, 0])
    np.testing.assert_array_equal(jax_prng.PRNGKey(42), [0, 42])
    np.testing.assert_array_equal(jax_prng.PRNGKey((7 << 32) | 9), [7, 9])
    np.testing.assert_array_equal(jax_prng.split(jax_prng.PRNGKey(0), partitionable=False),
                                  [[4146024105, 967050713], [2718843009, 1272950319]])


def test_normal_original_layout_matches_jax_tutorial():
    k = jax_prng.PRNGKey(42)
    np.testing.assert_allclose(ref.normal(k, 1, partitionable=False), [-0.18471177], rtol=2e-7)
    np.testing.assert_allclose(ref.normal(k, 3, partitionable=False), [0.18693547, -1.2806505, -1.5593132], rtol=2e-7)
    sub = jax_prng.split(k, 3, partitionable=False)
    np.testing.assert_allclose([ref.normal(s, 1, partitionable=False)[0] for s in sub], [-0.04838832, 0.10796154, -1.2226542], rtol=2e-7)


def test_partitionable_layout_is_the_same_block_function_on_flat_counters():
    k = jax_prng.PRNGKey(3)
    sp = jax_prng.split(k, 4)
    for i in range(4):
        assert tuple(int(v) for v in sp[i]) == ref.threefry2x32_scalar(0, 3, 0, i)
    bits = ref.random_bits(k, 7)
    a, b = jax_prng.threefry2x32(k, np.zeros(7, np.uint32), np.arange(7, dtype=np.uint32))
    np.testing.assert_array_equal(bits, a ^ b)
    z = ref.normal(k, 20000)
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.03 and np.isfinite(z).all()


def test_as_key_accepts_seed_or_raw_key():
    np.testing.assert_array_equal(jax_prng.as_key(5), [0, 5])
    np.testing.assert_array_equal(jax_prng.as_key(np.array([1, 2])), [1, 2])
    with pytest.raises(ValueError):
        jax_prng.as_key([1, 2, 3])


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("original", [0, 1])
def test_device_normal_matches_oracle(original):
    import torch
    from manipulator_mujoco_b200 import _lib
    from manipulator_mujoco_b200 import cem_planner
    pl = cem_planner(num_dof=6, num_batch=64, num_steps=16, timestep=0.05, maxiter_cem=1, num_elite=0.05, w_pos=20.0, w_rot=3.0,
                     w_col=80.0, maxiter_projection=10)
    lib, h = pl._lib, pl._h
    for key, n in (((0, 0), 66 * 31), ((0, 42), 3), ((0xDEADBEEF, 0x12345678), 4097)):
        out = torch.empty(n, device="cuda")
        _lib.check(lib.cemk_jax_normal(h, key[0], key[1], original, n, 0, n, C.c_void_p(out.data_ptr()), None), lib)
        torch.cuda.synchronize()
        want = ref.normal(np.array(key, dtype=np.uint32), n, partitionable=not original)
        got = out.cpu().numpy()
        # same bits, same polynomial; log1p / sqrt may differ in the last place
        np.testing.assert_allclose(got, want, rtol=3e-6, atol=1e-7)
        # a slice generated with an offset is the same stream (what every rank of a sharded planner relies on)
        if n > 100:
            part = torch.empty(50, device="cuda")
            _lib.check(lib.cemk_jax_normal(h, key[0], key[1], original, n, 37, 50, C.c_void_p(part.data_ptr()), None), lib)
            assert torch.equal(part, out[37:87])
    if original:
        out = torch.empty(3, device="cuda")
        _lib.check(lib.cemk_jax_normal(h, 0, 42, 1, 3, 0, 3, C.c_void_p(out.data_ptr()), None), lib)
        np.testing.assert_allclose(out.cpu().numpy(), [0.18693547, -1.2806505, -1.5593132], rtol=3e-6)     # JAX tutorial


@pytest.mark.gpu
def test_planner_samples_follow_the_reference_key_chain():
    """cem_planner.key = PRNGKey(0); compute_cem splits once (:388), compute_xi_samples splits again (:314)
    and draws multivariate_normal(key, mean, cov + 0.003 I, (B,)) with the Cholesky method (:315)."""
    from manipulator_mujoco_b200 import cem_planner
    B = 100
    pl = cem_planner(num_dof=6, num_batch=B, num_steps=16, timestep=0.05, maxiter_cem=1, num_elite=0.05, w_pos=20.0, w_rot=3.0,
                     w_col=80.0, maxiter_projection=10)
    np.testing.assert_array_equal(pl.key, [0, 0])
    k1 = jax_prng.split(pl.key)[0]
    rng = np.random.default_rng(1)
    A = rng.normal(size=(66, 66))
    cov = (A @ A.T / 66 + np.eye(66)).astype(np.float32)
    mean = rng.normal(size=66).astype(np.float32)
    xi, k2 = pl.compute_xi_samples(k1, mean, cov)
    np.testing.assert_array_equal(k2, jax_prng.split(k1)[0])
    want = ref.multivariate_normal(k2, mean.astype(np.float64), cov.astype(np.float64) + 0.003 * np.eye(66), B)
    np.testing.assert_allclose(xi.cpu().numpy(), want, rtol=0, atol=3e-5)
    # the key never advances on the planner object (:80, :388): every tick reuses the same draws
    np.testing.assert_array_equal(pl.key, [0, 0])
