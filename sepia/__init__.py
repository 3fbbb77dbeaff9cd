"""Drop-in `sepia` package: re-exports the gladsgp_b200 mirror of the SEPIA classes that
timghill/GladsGP imports, so that the reference's src/model.py and analysis scripts
run unchanged on the B200-native path."""
import sys as _sys

from gladsgp_b200.sepia import SepiaData as _d, SepiaModel as _m, SepiaPrior as _pr, SepiaMCMC as _mc
from gladsgp_b200.sepia import SepiaPredict as _pd, SepiaPlot as _pl
from gladsgp_b200.sepia import SepiaParam as _SepiaParamClass
import gladsgp_b200.sepia.SepiaParam as _pm   # noqa: F401  (sys.modules entry of the submodule)

for _name, _mod in (('SepiaData', _d), ('SepiaModel', _m), ('SepiaPrior', _pr), ('SepiaMCMC', _mc),
                    ('SepiaPredict', _pd), ('SepiaPlot', _pl)):
    _sys.modules[__name__ + '.' + _name] = _mod
    globals()[_name] = _mod
_sys.modules[__name__ + '.SepiaParam'] = _sys.modules['gladsgp_b200.sepia.SepiaParam']
SepiaParam = _SepiaParamClass          # `from sepia import SepiaParam` is called as a class (src/model.py:15,227)
