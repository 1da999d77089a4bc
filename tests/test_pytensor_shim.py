"""Row N2 of VERDICT r1: the PyTensor adapter (pymc_statespace_b200/pytensor_op.py) executed through the Op protocol.

PyTensor itself cannot be installed here, so these tests drive `make_node -> perform -> L_op -> GradOp.perform`
through tests/fake_pytensor, a minimal stand-in of the protocol (typed variables, Apply, Op.__call__, an evaluator
that calls `perform`, a reverse-mode `grad` that calls `L_op` with DisconnectedType cotangents and honours
`connection_pattern`).  What is checked is OUR wiring - argument order, optional c / d, output types, shapes,
disconnected inputs / outputs - and, on the GPU, that the numbers coming out of the Op pair equal the oracle's."""
import importlib
import os
import sys

import numpy as np
import pytest

from tests.helpers import random_system, rel_err

SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fake_pytensor")


@pytest.fixture()
def shim():
    """pymc_statespace_b200.pytensor_op re-imported against the stand-in; restored afterwards."""
    sys.path.insert(0, SHIM)
    for k in [k for k in sys.modules if k == "pytensor" or k.startswith("pytensor.")]:
        del sys.modules[k]
    import pymc_statespace_b200.pytensor_op as pop

    pop = importlib.reload(pop)
    assert pop.HAVE_PYTENSOR
    import pytensor

    yield pop, pytensor
    sys.path.remove(SHIM)
    for k in [k for k in sys.modules if k == "pytensor" or k.startswith("pytensor.")]:
        del sys.modules[k]
    importlib.reload(pop)


def _symbolic_inputs(pt, with_c, with_d):
    names = ["data", "a0", "P0", "T", "Z", "R", "H", "Q"] + (["c"] if with_c else []) + (["d"] if with_d else [])
    return names, [pt.dtensor3(n) if n == "data" else pt.dmatrix(n) for n in names]


@pytest.mark.parametrize("with_c,with_d", [(False, False), (True, True), (False, True), (True, False)])
def test_graph_construction_and_protocol_wiring(shim, with_c, with_d):
    pop, pytensor = shim
    import pytensor.tensor as pt
    from pytensor.gradient import DisconnectedType

    from pymc_statespace_b200.filters import FILTER_FACTORY

    names, ins = _symbolic_inputs(pt, with_c, with_d)
    flt = FILTER_FACTORY["standard"]()
    kw = dict(zip(names[8:], ins[8:]))
    outs = flt.build_graph(*ins[:8], **kw)          # PyTensor variables in -> symbolic outputs of ONE Op
    assert len(outs) == 6 and [o.ndim for o in outs] == [3, 3, 3, 3, 0, 1]
    node = outs[4].owner
    assert isinstance(node.op, pop.KalmanFilterOp) and node.op.has_c == with_c and node.op.has_d == with_d
    assert [i.name for i in node.inputs] == names   # argument order at the seam (kalman_filter.py:126-128)
    assert node.op == pop.KalmanFilterOp("standard", True, with_c, with_d) and hash(node.op) == hash(
        pop.KalmanFilterOp("standard", True, with_c, with_d))                                   # __props__ equality
    shapes = node.op.infer_shape(None, node, [(10, 1, 1), (3, 1)] + [None] * (len(names) - 2))
    assert shapes == [(10, 3, 1), (11, 3, 1), (10, 3, 3), (11, 3, 3), (), (10,)]
    pat = node.op.connection_pattern(node)
    assert len(pat) == len(names) and pat[0] == [False] * 6 and all(p == [False] * 4 + [True, True] for p in pat[1:])
    # d logp / d inputs: one GradOp node, one cotangent per matrix input, data disconnected
    g = pytensor.grad(outs[4], ins)
    assert g[0] is None and all(gi is not None for gi in g[1:])
    gnode = g[1].owner
    assert isinstance(gnode.op, pop.KalmanFilterGradOp) and len(gnode.outputs) == len(names) - 1
    assert [i.name for i in gnode.inputs[:len(names)]] == names and len(gnode.inputs) == len(names) + 2
    assert [o.ndim for o in gnode.outputs] == [2] * (len(names) - 1)
    assert gnode.op.infer_shape(None, gnode, [(10, 1, 1)] + [(k, k) for k in range(1, len(names) + 2)]) == [
        (k, k) for k in range(1, len(names))]
    # a cotangent arriving on a moment output is refused (only log_likelihood / ll_obs are differentiable)
    with pytest.raises(NotImplementedError):
        node.op.L_op(node.inputs, node.outputs, [pt.dtensor3("g")] + [DisconnectedType()() for _ in range(5)])
    # steady_state / univariate reject time-varying inputs at graph-construction time, whatever optional inputs exist
    tv = list(ins)
    tv[names.index("T")] = pt.dtensor3("T")
    with pytest.raises(ValueError):
        FILTER_FACTORY["univariate"]().build_graph(*tv[:8], **dict(zip(names[8:], tv[8:])))
    if with_d:
        tv = list(ins)
        tv[names.index("d")] = pt.dtensor3("d")   # used to be labelled "c" when only d was given
        with pytest.raises(ValueError):
            FILTER_FACTORY["steady_state"]().build_graph(*tv[:8], **dict(zip(names[8:], tv[8:])))


@pytest.mark.gpu
@pytest.mark.parametrize("kind,dims,with_c,with_d", [("standard", (3, 2, 2), True, True), ("standard", (2, 1, 1), False, True),
                                                     ("univariate", (3, 2, 2), False, False), ("single", (2, 1, 1), True, False),
                                                     ("cholesky", (4, 3, 2), False, False), ("steady_state", (2, 1, 1), False, False)])
def test_op_pair_evaluates_to_oracle_values_and_gradients(shim, kind, dims, with_c, with_d):
    """pytensor.function([...], outputs + grads) through the stand-in: `KalmanFilterOp.perform` (forward kernels) and
    `KalmanFilterGradOp.perform` (loglik-only forward + adjoint kernel) against the numpy / torch-autograd oracle."""
    pop, pytensor = shim
    import pytensor.tensor as pt

    from oracle import kalman_numpy as kn
    from oracle import kalman_torch as kt
    from pymc_statespace_b200.filters import FILTER_FACTORY

    m, p, r = dims
    rng = np.random.default_rng(sum(dims))
    args = list(random_system(rng, m, p, r, 20, n_missing=2))
    c = rng.normal(size=(m, 1)) if with_c else None
    d = rng.normal(size=(p, 1)) if with_d else None
    names, ins = _symbolic_inputs(pt, with_c, with_d)
    vals = args + ([c] if with_c else []) + ([d] if with_d else [])
    outs = FILTER_FACTORY[kind]().build_graph(*ins[:8], **dict(zip(names[8:], ins[8:])))
    w = rng.normal(size=20)
    for cost, gobs in ((outs[4], None), (pt._Scale()(outs[5], w), w)):   # logp, and a weighted sum over ll_obs
        grads = pytensor.grad(cost, ins) if gobs is None else None
        if gobs is not None:
            # seed the vector cotangent of ll_obs directly (the stand-in has no reductions)
            from pytensor.gradient import DisconnectedType

            node = outs[5].owner
            og = [DisconnectedType()() for _ in range(5)] + [pt.as_tensor_variable(w)]
            grads = node.op.L_op(node.inputs, node.outputs, og)
        f = pytensor.function(ins, list(outs) + [g for g in grads[1:]])
        res = f(*vals)
        ref = kn.kalman_filter(kind, *args, c=c, d=d)
        for a, b in zip(res[:6], ref):
            assert rel_err(a, b) < 1e-8
        _, gref = kt.loglik_and_grads(kind, *args, c=c, d=d, g_ll_obs=gobs)
        tol = 1e-7 if kind == "steady_state" else 1e-8
        for name, got in zip(names[1:], res[6:]):
            assert got.shape == np.asarray(vals[names.index(name)]).shape
            scale = max(np.abs(gref[name]).max(), 1e-12 * max(np.abs(v).max() for v in gref.values()))
            assert np.abs(got - gref[name]).max() / scale < tol, (kind, name)
