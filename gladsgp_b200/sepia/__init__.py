"""Host-side mirror of the SEPIA class surface that GladsGP drives.

Only the sim-only ("eta-only") emulator model is implemented -- the one mode the reference uses
(/root/reference/examples/04_GP_emulation_multivariate_ensemble.ipynb:256).  All heavy arithmetic
runs in the sm_100a kernels behind libgladsgp_b200.so; there is no NumPy fallback for it.
"""
from . import SepiaData as _SepiaData_mod            # noqa: F401
from . import SepiaPrior as _SepiaPrior_mod          # noqa: F401
from . import SepiaMCMC as _SepiaMCMC_mod            # noqa: F401
from . import SepiaModel as _SepiaModel_mod          # noqa: F401
from . import SepiaPredict as _SepiaPredict_mod      # noqa: F401
from . import SepiaPlot                              # noqa: F401
from .SepiaParam import SepiaParam                   # noqa: F401  (class, as `from sepia import SepiaParam`)
