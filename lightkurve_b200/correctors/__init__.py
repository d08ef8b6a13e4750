"""Correctors on the hot path: DesignMatrix(Collection) and RegressionCorrector
(/root/reference/src/lightkurve/correctors/{designmatrix,regressioncorrector}.py)."""
from .designmatrix import (DesignMatrix, DesignMatrixCollection, SparseDesignMatrix,  # noqa: F401
                           SparseDesignMatrixCollection, create_spline_matrix, create_sparse_spline_matrix)
from .regressioncorrector import RegressionCorrector  # noqa: F401
from .metrics import overfit_metric_lombscargle  # noqa: F401
from .cbvcorrector import CBVCorrector, CotrendingBasisVectors  # noqa: F401
