"""CPU oracle for the Kalman logp(+grad) hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pymc_statespace_b200/`` may import
this package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and
there only as the checker / the reported CPU baseline.

Contents
--------
kalman_numpy   literal NumPy/SciPy restatement of the five reference filters
               (``/root/reference/pymc_statespace/filters/kalman_filter.py``).
kalman_torch   the same equations in torch-fp64 so autograd supplies the
               gradient oracle (the reference's gradient is PyTensor autodiff
               of the same graph).
solvers        Lyapunov / DARE forward + the adjoint formulas the reference
               uses (``utils/pytensor_scipy.py:39-60``).
models         theta -> system-matrix maps of BayesianARMA / BayesianVARMAX /
               BayesianLocalLevel (``models/*.py``).
kalman_c.c     plain-C port (forward + hand-written adjoint, OpenMP over draws)
               used as the timed CPU baseline; built into ``oracle/_build/``.

Parity status: the reference cannot be imported in this image (PyTensor, PyMC
and statsmodels are absent), and its tests hold no stored
golden vectors (they compare against statsmodels live).  The oracle is pinned
on what IS available offline - see ``tests/test_oracle_pins.py`` - and is
otherwise "parity unpinned" (gradient values, multivariate values).
"""
