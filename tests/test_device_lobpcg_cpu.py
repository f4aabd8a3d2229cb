"""`cluster.device_lobpcg` (the LOBPCG recurrence of the cut with device-resident block vectors) against scipy's own
iteration, with the device replaced by a numpy stand-in of the five block-vector operations: the port of the control flow
(active mask, restarts, implicit / explicit Gram matrices, final Rayleigh-Ritz) is checked on the CPU; the CUDA kernels
behind `ctx.bv_*` are checked on the GPU (tests/test_gpu_cluster.py compares the cut's labels with sklearn's)."""
import numpy as np
import pytest
from scipy.sparse.csgraph import laplacian as csgraph_laplacian
from scipy.sparse.linalg import lobpcg

from phertilizer_b200.cluster import device_lobpcg


class NumpyBlocks:
    """numpy stand-in for CopyDistance.bv_*: ten k x 3 slots and a dense operator."""

    def __init__(self, L):
        self.L = L
        self._k = L.shape[0]
        self.slots = np.zeros((10, self._k, 3))
        self.applies = 0

    def bv_set(self, slot, X):
        self.slots[slot] = X

    def bv_get(self, slot):
        return self.slots[slot].copy()

    def bv_apply(self, src, dst):
        self.applies += 1
        self.slots[dst] = self.L @ self.slots[src]

    def bv_gram(self, left, right):
        Lm = np.concatenate([self.slots[s] for s in left], axis=1)
        Rm = np.concatenate([self.slots[s] for s in right], axis=1)
        return Lm.T @ Rm

    def bv_combine(self, dst, terms):
        self.slots[dst] = sum(self.slots[s] @ np.asarray(C, float).reshape(3, 3) for s, C in terms)


def normalised_laplacian(k, rng, blocks=2):
    lab = rng.integers(0, blocks, k)
    pts = rng.standard_normal((k, 2)) + 4.0 * lab[:, None]
    d = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1))
    W = np.exp(-d / (0.2 * d.max()))
    L, dd = csgraph_laplacian(W, normed=True, return_diag=True)
    np.fill_diagonal(L, 1.0)
    return L, dd


@pytest.mark.parametrize("k,seed", [(120, 0), (400, 1), (900, 2), (64, 3)])
def test_device_lobpcg_follows_scipy(k, seed):
    rng = np.random.default_rng(seed)
    L, dd = normalised_laplacian(k, rng)
    X0 = np.random.RandomState(seed).standard_normal(size=(k, 3))
    X0[:, 0] = dd
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lam_ref, V_ref = lobpcg(L, X0.copy(), tol=None, largest=False, maxiter=2000)
    ctx = NumpyBlocks(L)
    lam, V = device_lobpcg(ctx, X0.copy(), maxiter=2000)
    assert np.allclose(lam, lam_ref, rtol=1e-8, atol=1e-10)
    # same invariant subspace: the eigenvectors agree up to sign wherever the eigenvalues are separated
    for j in range(3):
        c = abs(float(V[:, j] @ V_ref[:, j])) / (np.linalg.norm(V[:, j]) * np.linalg.norm(V_ref[:, j]))
        assert c > 1 - 1e-6, (j, c)
    # and the cut sklearn derives from them is the same
    from sklearn.cluster._spectral import cluster_qr
    from sklearn.utils.extmath import _deterministic_vector_sign_flip

    def cut(Vm):
        e = Vm.T[:2] / dd
        return cluster_qr(_deterministic_vector_sign_flip(e).T)

    assert np.array_equal(cut(V), cut(V_ref))
