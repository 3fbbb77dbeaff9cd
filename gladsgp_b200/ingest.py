"""Device-resident ensemble ingest: the streaming passes of the reference's init_model /
fit_models around the PCA (/root/reference/src/model.py:60-73, 218-224) as CUDA kernels (csrc/ggp_ingest.cu).

The ensemble is uploaded once; column statistics, standardisation, the rSVD passes, the projection on the basis
(`w`), the residual precision and the lamWOs prior terms are all computed from that device copy.  Host copies of
y_std are only materialised when a caller reads `data.sim_data.y_std`.
"""
import numpy as np

from . import _lib, ops

# ensembles with at least this many elements take the device path in SepiaData / SepiaModel
DEVICE_MIN_ELEMS = 1 << 22
_STAGE_BYTES = 64 << 20


def use_device(n_elems):
    return n_elems >= DEVICE_MIN_ELEMS


def upload(a):
    """Host float32 C-contiguous array -> device tensor, through two pinned staging buffers (the DMA of one chunk
    overlaps the host copy of the next; page-locking a multi-GB array in place costs more than the copy)."""
    torch = _lib.require_cuda()
    if torch.is_tensor(a):
        return a.to(device='cuda', dtype=torch.float32)
    a = np.asarray(a)
    if a.dtype != np.float32 or not a.flags['C_CONTIGUOUS']:
        a = np.ascontiguousarray(a, dtype=np.float32)
    out = torch.empty(a.shape, dtype=torch.float32, device='cuda')
    flat_h = a.reshape(-1)
    flat_d = out.view(-1)
    n = flat_h.size
    if a.nbytes <= _STAGE_BYTES:
        flat_d.copy_(torch.from_numpy(flat_h))
        return out
    per = _STAGE_BYTES // 4
    stage = [torch.empty(per, dtype=torch.float32).pin_memory() for _ in range(2)]
    done = [None, None]
    i = 0
    for o in range(0, n, per):
        k = min(per, n - o)
        if done[i] is not None:
            done[i].synchronize()
        stage[i][:k].numpy()[:] = flat_h[o:o + k]
        flat_d[o:o + k].copy_(stage[i][:k], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        done[i] = ev
        i ^= 1
    torch.cuda.synchronize()
    return out


def column_stats(y_dev, transposed=False, sd_threshold=1e-6, ddof=1):
    """src/model.py:60-64 on the device: (mean, sd) float32 device vectors; sd < threshold is replaced by it."""
    return ops.colstats(y_dev, transposed=transposed, ddof=ddof, sd_floor=sd_threshold)


def project_basis(ystd_dev, K_dev):
    """w = (pinv(K)^T y_std^T)^T and the residual sums of y_std - w K, from one pass over y_std (+ one over K).

    pinv(K) = K^T (K K^T)^-1 for a basis of full row rank, so w = (y_std K^T)(K K^T)^-1; with P = y_std K^T and
    G = K K^T:  ||y_std - w K||^2 = sum(y_std^2) - 2 sum(w * P) + sum((w G) * w)  and
    sum(y_std - w K) = sum(y_std) - sum_p (sum_i w_ip)(sum_c K_pc).  (src/model.py:219-223; SepiaModel.__init__)"""
    m, n = ystd_dev.shape
    pu = K_dev.shape[0]
    P = ops.project(ystd_dev, K_dev).cpu().numpy()
    Pk = ops.project(K_dev, K_dev).cpu().numpy()
    YK, s_y, ss_y = P[:, :pu], float(P[:, pu].sum()), float(P[:, pu + 1].sum())
    G, ksum = Pk[:, :pu], Pk[:, pu]
    G = 0.5 * (G + G.T)
    w = np.linalg.solve(G, YK.T).T
    resid_ss = max(ss_y - 2.0 * float(np.sum(w * YK)) + float(np.sum((w @ G) * w)), 0.0)
    resid_sum = s_y - float(w.sum(axis=0) @ ksum)
    return dict(w=w, G=G, resid_ss=resid_ss, resid_sum=resid_sum, n_elems=m * n)


def pc_precision_from(proj):
    """1 / np.var(y_std - w K)  (src/model.py:221-223)."""
    N = proj['n_elems']
    return 1.0 / (proj['resid_ss'] / N - (proj['resid_sum'] / N) ** 2)
