class Variable:
    def __init__(self, type_, owner=None, index=None, name=None, value=None):
        self.type, self.owner, self.index, self.name, self.value = type_, owner, index, name, value

    @property
    def ndim(self):
        return self.type.ndim

    @property
    def dtype(self):
        return getattr(self.type, "dtype", None)

    def __repr__(self):
        return f"Var({self.name or id(self)}: {self.type})"


class Apply:
    def __init__(self, op, inputs, outputs):
        self.op, self.inputs, self.outputs = op, list(inputs), list(outputs)
        for i, o in enumerate(self.outputs):
            o.owner, o.index = self, i
