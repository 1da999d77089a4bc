from ..graph.basic import Variable


class TensorType:
    def __init__(self, dtype, shape):
        self.dtype, self.shape = dtype, tuple(shape)

    @property
    def ndim(self):
        return len(self.shape)

    def __call__(self, name=None):
        return Variable(self, name=name)

    def __repr__(self):
        return f"TensorType({self.dtype}, {self.shape})"
