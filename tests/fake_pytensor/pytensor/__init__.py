"""TEST-ONLY stand-in for the parts of the PyTensor Op protocol that pymc_statespace_b200/pytensor_op.py touches.

PyTensor is not installable in the build image (no wheel, no network), so `KalmanFilterOp.make_node / infer_shape /
connection_pattern / L_op` and `KalmanFilterGradOp.perform` could never run there.  This package implements the
protocol as the reference relies on it (precedent: the reference's own SolveDiscreteARE Op,
pymc_statespace/utils/pytensor_scipy.py:11-60) in ~150 lines: typed variables, Apply nodes, `Op.__call__`,
a graph evaluator that calls `perform`, and a reverse-mode `grad` that calls `L_op` with `DisconnectedType`
cotangents for outputs the cost does not depend on.  It is NOT PyTensor: it checks our wiring against the protocol as
we understand it, nothing more.  tests/test_pytensor_shim.py puts this directory on sys.path for its own duration.
"""
from . import gradient, graph, tensor  # noqa: F401
from .evaluate import function  # noqa: F401
from .gradient import grad  # noqa: F401

__version__ = "0.0-shim"
