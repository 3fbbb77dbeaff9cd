"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's ensemble initialisation passes.

Follows /root/reference/src/model.py line by line (plain NumPy, as the reference itself):
  :60-64    column mean / std(ddof=1) with the sd_threshold clamp
  :71-72    y_std = (y_sim - mu_y) / sd_y
  :219-223  w = (pinv(K)^T y_std^T)^T ; pc_prec = 1 / var(y_std - w K)
and SEPIA's default lamWOs prior terms.  Only tests/, __graft_entry__.smoke() and bench.py's
CPU baseline may import this module; the product path (gladsgp_b200/) never does.
Pinning: these are the reference's own in-tree NumPy expressions, evaluated by NumPy here; there is no recorded
numerical output for them in the reference tree (the ensemble is not committed), so the check is restatement vs CUDA.
"""
import numpy as np


def column_stats(y_sim, sd_threshold=1e-6):
    """src/model.py:60-64."""
    mu_y = np.mean(y_sim, axis=0)
    sd_y = np.std(y_sim, ddof=1, axis=0)
    sd_y[sd_y < sd_threshold] = sd_threshold
    return mu_y, sd_y


def standardize(y_sim, mu_y, sd_y):
    """src/model.py:72 (and SepiaData.standardize_y with given y_mean / y_sd)."""
    return (y_sim - mu_y) / sd_y


def pc_weights_and_precision(y_std, K):
    """src/model.py:219-223, in float64 so that it can serve as the reference value for the FP64-accumulating
    device pass (the reference evaluates the same lines in the dtype of its inputs, float32)."""
    K = np.asarray(K, dtype=np.float64)
    y_std = np.asarray(y_std, dtype=np.float64)
    w = np.dot(np.linalg.pinv(K).T, y_std.T).T
    y_hat = np.dot(w, K)
    pc_resid = y_std - y_hat
    pc_var = np.var(pc_resid)
    return w, 1.0 / pc_var, float(np.sum(pc_resid ** 2))
