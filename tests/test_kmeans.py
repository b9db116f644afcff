"""compute_cvt_centroids(backend="gpu") (reference qdax/core/containers/mapelites_repertoire.py:30-72 calls
scikit-learn KMeans on the host -- not reproducible bit for bit): the GPU Lloyd iterations against the NumPy
restatement of the same rule (bit-exact), plus the properties any CVT must have."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import jax_prng as jr  # noqa: E402
from oracle import qdax_numpy as qn  # noqa: E402


def inertia(x, cent):
    c = qn.get_cells_indices(x, cent)
    return float(((x - cent[c]).astype(np.float64) ** 2).sum())


def test_oracle_lloyd_is_a_descent_and_reaches_a_fixed_point():
    rng = np.random.default_rng(0)
    x = rng.random((4000, 2)).astype(np.float32)
    prev = inertia(x, x[:64])
    for iters in (1, 2, 5, 20):
        cent, ran = qn.lloyd_cvt_centroids(x, 64, iters)
        cur = inertia(x, cent)
        assert cur <= prev * (1 + 1e-6) and cent.min() >= 0.0 and cent.max() < 1.0
        prev = cur
    cent, ran = qn.lloyd_cvt_centroids(x, 64, 500)
    assert ran < 500                                   # converged: assignments stopped changing
    # running longer from the same start changes nothing: it is a fixed point
    assert np.allclose(qn.lloyd_cvt_centroids(x, 64, ran + 5)[0], cent)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("N,K,Dd,iters", [(20000, 256, 2, 12), (3000, 300, 3, 40), (40000, 2048, 16, 6), (30000, 1500, 32, 4)])
def test_gpu_lloyd_bit_exact_vs_oracle(dev, N, K, Dd, iters):
    from qdax_b200 import random as qr
    from qdax_b200.core.containers.mapelites_repertoire import lloyd_cvt_centroids

    x = qr.uniform(jr.key(N + K), (N, Dd), device=dev)
    cent, ran = lloyd_cvt_centroids(x, K, iters)
    ref, ran_ref = qn.lloyd_cvt_centroids(x.cpu().numpy(), K, iters)
    assert ran == ran_ref
    assert np.array_equal(cent.cpu().numpy(), ref)


@pytest.mark.gpu
def test_compute_cvt_centroids_gpu_backend(dev):
    from qdax_b200.core.containers.mapelites_repertoire import compute_cvt_centroids

    cent = compute_cvt_centroids(2, 20000, 128, minval=[-1.0, 0.0], maxval=[1.0, 2.0], key=jr.key(0), device=dev, backend="gpu")
    c = cent.cpu().numpy()
    assert c.shape == (128, 2) and np.isfinite(c).all()
    assert c[:, 0].min() >= -1.0 and c[:, 0].max() <= 1.0 and c[:, 1].min() >= 0.0 and c[:, 1].max() <= 2.0
    # a CVT of the uniform density is close to uniform itself: every quadrant of the box holds about a quarter of the centroids
    quad = ((c[:, 0] > 0).astype(int) * 2 + (c[:, 1] > 1).astype(int))
    assert np.bincount(quad, minlength=4).min() >= 16
    # and it beats the initial guess by a wide margin, like scikit-learn's result on the same samples
    from qdax_b200 import random as qr
    x = qr.uniform(qr.split(jr.key(0))[1], (20000, 2), device=dev).cpu().numpy()
    unit = (c - np.array([-1.0, 0.0], np.float32)) / 2.0
    assert inertia(x, unit.astype(np.float32)) < 0.75 * inertia(x, x[:128])
    sk = compute_cvt_centroids(2, 20000, 128, minval=[-1.0, 0.0], maxval=[1.0, 2.0], key=jr.key(0), device=dev).cpu().numpy()
    sk_unit = ((sk - np.array([-1.0, 0.0], np.float32)) / 2.0).astype(np.float32)
    assert inertia(x, unit.astype(np.float32)) < 1.05 * inertia(x, sk_unit)
