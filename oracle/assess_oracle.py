"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's test-error post-processing.

Follows /root/reference/experiments/synthetic/analysis/assess_all_models.py:
  :489-500  per batch of test designs: ypreds = preds.get_y() (float32), one noise draw per (sample, design) scaled by sd_y,
            mean over samples, (q, 1-q) quantiles over samples of ypreds + error_preds;
  :523-538  RMSE per design, MAPE per design with outputs below the 10 % quantile of the whole test matrix masked out,
            coverage fraction, mean lower / upper limit; :498-500, :537-538 integrated interval width.
The enclosing function (compute_test_error) loads models through sepia / matplotlib imports and cannot be called here, so
these NumPy lines are restated one for one (parity of this row: restated from in-tree lines, not executed against them).

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this module.
"""
import numpy as np


def fields(w, K, sd_y, mu_y, noise, quantile):
    """w (nsamp, npred, pu) f32; K (pu, n_y) f32; noise (nsamp, npred) f32 -> ypred_mean, ypred_lq, ypred_uq (npred, n_y) f32."""
    w = np.asarray(w, dtype=np.float32)
    ypreds = (np.tensordot(w, np.asarray(K, dtype=np.float32), axes=[[2], [0]]) * sd_y + mu_y).astype(np.float32)   # get_y()
    error_preds = np.zeros(ypreds.shape, dtype=np.float32)
    for l_pred in range(ypreds.shape[1]):
        for l_sample in range(ypreds.shape[0]):
            error_preds[l_sample][l_pred] = sd_y * noise[l_sample, l_pred]                   # :493-497
    ypred_mean = np.mean(ypreds, axis=0)                                                     # :498
    ypred_lq = np.quantile(ypreds + error_preds, quantile, axis=0)                           # :499
    ypred_uq = np.quantile(ypreds + error_preds, 1 - quantile, axis=0)                       # :500
    return ypred_mean, ypred_lq, ypred_uq


def error_statistics(ypred_mean, ypred_lq, ypred_uq, y_test):
    """assess_all_models.py:523-538 -> dict of per-design arrays (+ the scalar 10 % quantile and integrated width)."""
    pred_resid = ypred_mean - y_test
    pred_rmse = np.sqrt(np.mean(pred_resid ** 2, axis=1))
    lq = np.quantile(y_test, 0.1)
    inner_mape = np.abs(pred_resid / y_test)
    inner_mape[y_test < lq] = np.nan
    pred_mape = np.nanmean(inner_mape, axis=1)
    is_covered = np.logical_and(y_test >= ypred_lq, y_test <= ypred_uq)
    frac_covered = is_covered.sum(axis=1) / is_covered.shape[1]
    pred_lq = np.mean(ypred_lq, axis=1)
    pred_uq = np.mean(ypred_uq, axis=1)
    integrated_ci = np.mean(ypred_uq - ypred_lq)          # mean over equal batches of the batch means of (uq - lq), :521,:537
    return dict(rmse=pred_rmse, mape=pred_mape, lq=pred_lq, uq=pred_uq, frac_covered=frac_covered,
                integrated_ci=integrated_ci, mape_floor=lq)
