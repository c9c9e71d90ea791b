"""TEST INFRASTRUCTURE -- numpy restatement of jax.random.normal / multivariate_normal (jax==0.5.3,
the reference's requirements.txt:9; call site mjx_planner.py:313-316) used to check the CUDA
generator `cemk_jax_normal`.  PARITY UNPINNED against jax itself (not installable here); pinned only
by the Threefry-2x32 known-answer vectors (tests/test_jax_prng.py).  Never imported by the product.

Follows jax/_src/prng.py (threefry_2x32, _threefry_random_bits_partitionable / _original),
jax/_src/random.py (_uniform, _normal_real, _multivariate_normal with method='cholesky') and XLA's
float32 erf_inv (xla/client/lib/math.cc, Giles' polynomial).
"""
import numpy as np

_ROT = (13, 15, 26, 6, 17, 29, 16, 24)
M32 = 0xFFFFFFFF


def threefry2x32_scalar(k0, k1, x0, x1):
    """Plain-int Threefry-2x32-20 (one block), written independently of the package's vectorised version."""
    ks = (k0, k1, (k0 ^ k1 ^ 0x1BD11BDA) & M32)
    x0 = (x0 + ks[0]) & M32
    x1 = (x1 + ks[1]) & M32
    for g in range(5):
        for r in _ROT[4 * (g % 2):4 * (g % 2) + 4]:
            x0 = (x0 + x1) & M32
            x1 = ((x1 << r) | (x1 >> (32 - r))) & M32
            x1 ^= x0
        x0 = (x0 + ks[(g + 1) % 3]) & M32
        x1 = (x1 + ks[(g + 2) % 3] + g + 1) & M32
    return x0, x1


def random_bits(key, n, partitionable=True):
    out = np.empty(n, dtype=np.uint32)
    k0, k1 = int(key[0]), int(key[1])
    if partitionable:
        for e in range(n):
            a, b = threefry2x32_scalar(k0, k1, 0, e)
            out[e] = a ^ b
    else:
        half = (n + 1) // 2
        for j in range(half):
            a, b = threefry2x32_scalar(k0, k1, j, half + j if half + j < n else 0)
            out[j] = a
            if half + j < n:
                out[half + j] = b
    return out


def erfinv32(x):
    x = x.astype(np.float32)
    w = -np.log1p(-x * x).astype(np.float32)
    lt = w < np.float32(5)
    w = np.where(lt, w - np.float32(2.5), np.sqrt(w, dtype=np.float32) - np.float32(3)).astype(np.float32)
    a = np.array([2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503,
                  -0.00417768164, 0.246640727, 1.50140941], dtype=np.float32)
    b = np.array([-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613,
                  0.00943887047, 1.00167406, 2.83297682], dtype=np.float32)
    p = np.where(lt, a[0], b[0]).astype(np.float32)
    for i in range(1, 9):
        p = (np.where(lt, a[i], b[i]) + p * w).astype(np.float32)
    return (p * x).astype(np.float32)


def normal(key, n, partitionable=True):
    bits = random_bits(key, n, partitionable)
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1)
    lo = np.nextafter(np.float32(-1), np.float32(0))
    u = np.maximum(lo, (f * (np.float32(1) - lo) + lo).astype(np.float32))
    return (np.float32(np.sqrt(2)) * erfinv32(u)).astype(np.float32)


def multivariate_normal(key, mean, cov, B, partitionable=True):
    n = mean.shape[0]
    z = normal(key, B * n, partitionable).reshape(B, n).astype(np.float64)
    L = np.linalg.cholesky(cov.astype(np.float64))
    return mean + z @ L.T
