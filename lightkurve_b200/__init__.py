"""lightkurve_b200 - B200-native periodogram-and-detrending engine behind the lightkurve API.

Only the hot path of lightkurve is here: Lomb-Scargle and BLS
periodograms, ``flatten`` and ``RegressionCorrector.correct`` - as hand-written sm_100a
CUDA kernels behind a C ABI (``include/lkb200.h``), with a thin mirror of the reference's
``LightCurve`` / ``LightCurveCollection`` / ``Periodogram`` Python surface on top.
"""
from . import engine  # noqa: F401
from .utils import LightkurveWarning, validate_method  # noqa: F401
from . import units  # noqa: F401
from .lightcurve import LightCurve, FoldedLightCurve  # noqa: F401
from .collections import LightCurveCollection  # noqa: F401
from .periodogram import Periodogram, LombScarglePeriodogram, BoxLeastSquaresPeriodogram  # noqa: F401
from . import correctors  # noqa: F401
from . import seismology  # noqa: F401
from .seismology import Seismology  # noqa: F401
from .correctors import DesignMatrix, DesignMatrixCollection, RegressionCorrector  # noqa: F401

__version__ = "0.1.0"
