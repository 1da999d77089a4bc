import numpy as np

from ..graph.basic import Apply, Variable
from ..graph.op import Op
from . import type as type  # noqa: F401,PLC0414
from .type import TensorType


def as_tensor_variable(x):
    if isinstance(x, Variable):
        return x
    a = np.asarray(x, dtype=np.float64)
    return Variable(TensorType("float64", shape=a.shape), value=a, name="const")


def _typed(name, ndim):
    def make(name_=None):
        return Variable(TensorType("float64", shape=(None,) * ndim), name=name_ or name)
    return make


dscalar, dvector, dmatrix, dtensor3 = _typed("s", 0), _typed("v", 1), _typed("m", 2), _typed("t3", 3)


def zeros(shape, dtype="float64"):
    return as_tensor_variable(np.zeros(shape, dtype=dtype))


class _ZerosLike(Op):
    def make_node(self, x):
        return Apply(self, [x], [x.type()])

    def perform(self, node, inputs, output_storage):
        output_storage[0][0] = np.zeros_like(inputs[0])


def zeros_like(x):
    return _ZerosLike()(as_tensor_variable(x))


class _Scale(Op):  # cost = w * x (+ ...): enough to build scalar costs in the tests
    def make_node(self, x, w):
        return Apply(self, [as_tensor_variable(x), as_tensor_variable(w)], [as_tensor_variable(x).type()])

    def perform(self, node, inputs, output_storage):
        output_storage[0][0] = np.asarray(inputs[0] * inputs[1])

    def L_op(self, inputs, outputs, output_grads):
        return [_Scale()(output_grads[0], inputs[1]), None]
