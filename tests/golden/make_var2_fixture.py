"""VAR(2) MLE coefficients printed (4 decimals) in reference examples/'VARMAX Example.ipynb' cell 7
(statsmodels VARMAX(order=(2,0), trend='n') on log-diff macrodata; Log Likelihood 1950.186, :363).
Writes tests/golden/var2_macrodata_params.npz.  Run from the repo root."""
import numpy as np

A1 = np.array([[-0.2170, 0.6959, 0.0197], [0.1379, 0.3248, -0.0233], [-3.0702, 4.2246, 0.4428]])
A2 = np.array([[0.0607, 0.3283, -0.0167], [0.0649, 0.3640, -0.0095], [-0.4455, 0.2350, 0.0191]])
L = np.array([[0.0075, 0.0, 0.0], [0.0042, 0.0056, 0.0], [0.0282, -0.0203, 0.0215]])
np.savez("tests/golden/var2_macrodata_params.npz", A1=A1, A2=A2, L=L)
