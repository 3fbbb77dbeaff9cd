"""Placeholder for `from sepia import SepiaPlot` (imported, never used, by
/root/reference/experiments/synthetic/analysis/assess_all_models.py:30).  Plotting is outside the
hot path."""


def __getattr__(name):
    raise NotImplementedError('sepia.SepiaPlot.%s: plotting is out of scope of gladsgp_b200' % name)
