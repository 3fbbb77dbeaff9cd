"""CPU oracle for the ensemble PCA (randomized SVD).  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/src/svd.py:12-82 (Halko et al. fixed-rank randomized SVD) with one
addition: the Gaussian test matrix can be injected so that results are reproducible and
comparable with the CUDA path (the reference draws it from the unseeded global np.random
stream at src/svd.py:51).

PARITY STATUS: pinned.  tests/golden/make_golden.py imports the reference's own
src/svd.py in the build container, runs it with a seeded global stream, and commits the
outputs as tests/golden/rsvd_*.npz; tests/test_oracle_cpu.py checks this restatement against
those fixtures bit-for-bit (same NumPy/BLAS) or to FP32 round-off.
"""
import numpy as np


def randomized_svd(X, p, k=None, q=1, omega=None, rng=np.random):
    """Same call shape as src/svd.py:12; evaluation order as src/svd.py:51-68.

    Note the reference evaluates ``X @ X.T @ Y`` left to right, i.e. forms the (m, m) Gram
    matrix first (src/svd.py:56).
    """
    if k is None:
        k = p
    if omega is None:
        omega = rng.normal(size=(X.shape[1], p + k)).astype(np.float32)     # svd.py:51
    Y = X @ omega                                                            # svd.py:52
    for _ in range(q):
        Y = X @ X.T @ Y                                                      # svd.py:56
    Q, _ = np.linalg.qr(Y, mode='reduced')                                   # svd.py:59
    B = Q.T @ X                                                              # svd.py:60
    U, S, V = np.linalg.svd(B, full_matrices=False)                          # svd.py:63
    U = Q @ U                                                                # svd.py:64
    return U[:, :p], S[:p], V[:p, :]                                         # svd.py:66-68


def k_basis(S, Vh, p, m):
    """src/model.py:101  K = diag(S[:p]) @ Vh[:p] / sqrt(m)."""
    return np.diag(S[:p]) @ Vh[:p] / np.sqrt(m)
