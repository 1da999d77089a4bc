import numpy as np


def _eval(v, env, cache):
    if id(v) in env:
        return env[id(v)]
    if v.owner is None:
        if v.value is None:
            raise KeyError(f"no value for input {v}")
        return v.value
    node = v.owner
    if id(node) not in cache:
        ins = [np.asarray(_eval(i, env, cache)) for i in node.inputs]
        storage = [[None] for _ in node.outputs]
        node.op.perform(node, ins, storage)
        cache[id(node)] = [s[0] for s in storage]
    return cache[id(node)][v.index]


def function(inputs, outputs):
    """pytensor.function(inputs, outputs) for graphs of Ops with a Python `perform`."""
    single = not isinstance(outputs, (list, tuple))
    outs = [outputs] if single else list(outputs)

    def f(*values):
        env = {id(i): np.asarray(val, dtype=np.float64) for i, val in zip(inputs, values)}
        cache = {}
        res = [np.asarray(_eval(o, env, cache)) for o in outs]
        return res[0] if single else res

    return f
