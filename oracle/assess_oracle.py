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
    """w (nsamp, npred, pu) f32; K (pu, n_y) f32; noise (nsamp, npred) f32 -> mean, lower, upper fields (npred, n_y) f32.
    get_y() in float32, one noise value per (sample, design) scaled by sd_y (:489-497), mean and quantiles over samples (:498-500)."""
    w = np.asarray(w, dtype=np.float32)
    y_s = (np.tensordot(w, np.asarray(K, dtype=np.float32), axes=[[2], [0]]) * sd_y + mu_y).astype(np.float32)
    z_s = y_s + (np.asarray(noise, dtype=np.float32)[:, :, None] * np.asarray(sd_y, dtype=np.float32)[None, None, :]).astype(np.float32)
    return np.mean(y_s, axis=0), np.quantile(z_s, quantile, axis=0), np.quantile(z_s, 1 - quantile, axis=0)


def error_statistics(y_mean, y_lo, y_hi, y_test):
    """Per-design error columns of performance_n{m}_p{p}.csv (:523-538) + the 10 % quantile used as MAPE floor and the
    integrated interval width."""
    resid = y_mean - y_test
    floor = np.quantile(y_test, 0.1)
    ape = np.where(y_test < floor, np.nan, np.abs(resid / y_test))
    inside = (y_test >= y_lo) & (y_test <= y_hi)
    return dict(rmse=np.sqrt(np.mean(resid ** 2, axis=1)), mape=np.nanmean(ape, axis=1), lq=np.mean(y_lo, axis=1),
                uq=np.mean(y_hi, axis=1), frac_covered=inside.sum(axis=1) / inside.shape[1],
                integrated_ci=np.mean(y_hi - y_lo), mape_floor=floor)
