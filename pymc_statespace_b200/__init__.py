"""pymc_statespace_b200 - B200-native (sm_100a) Kalman-filter log-likelihood and gradient behind the
filter plugin surface of jessegrabowski/pymc_statespace.  See DESIGN.md / INTEGRATION.md."""
from .engine import BatchedKalman, KalmanNumericalError, fp64_peak_tflops, lyapunov_backward, lyapunov_forward, rts_smoother  # noqa: F401

__version__ = "0.1.0"
