"""Correctors on the hot path: DesignMatrix(Collection) and RegressionCorrector
(/root/reference/src/lightkurve/correctors/{designmatrix,regressioncorrector}.py)."""
from .designmatrix import DesignMatrix, DesignMatrixCollection  # noqa: F401
from .regressioncorrector import RegressionCorrector  # noqa: F401
from .metrics import overfit_metric_lombscargle  # noqa: F401
