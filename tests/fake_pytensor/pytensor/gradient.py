from .graph.basic import Variable


class DisconnectedType:
    ndim = None

    def __call__(self, name=None):
        return Variable(self, name=name or "disconnected")

    def __repr__(self):
        return "DisconnectedType"


def grad(cost, wrt):
    """Reverse mode over ONE Apply level per output (enough for cost = f(Op outputs)): walks back from `cost`,
    calls each node's L_op with DisconnectedType cotangents for outputs that did not receive one, honours
    connection_pattern (a disconnected input must not receive a connected gradient)."""
    cot = {id(cost): (cost, None)}  # id -> (variable, cotangent variable or None for "seed = 1")
    order, seen = [], set()

    def visit(v):
        if v.owner is not None and id(v.owner) not in seen:
            seen.add(id(v.owner))
            for i in v.owner.inputs:
                visit(i)
            order.append(v.owner)

    visit(cost)
    from . import tensor as pt

    grads = {id(cost): pt.as_tensor_variable(1.0)}
    for node in reversed(order):
        og = [grads.get(id(o), None) for o in node.outputs]
        if all(g is None for g in og):
            continue
        og = [DisconnectedType()() if g is None else g for g in og]
        ig = node.op.L_op(node.inputs, node.outputs, og)
        assert len(ig) == len(node.inputs), "L_op must return one entry per input"
        pattern = node.op.connection_pattern(node) if hasattr(node.op, "connection_pattern") else None
        for k, (inp, g) in enumerate(zip(node.inputs, ig)):
            if g is None:
                continue
            if pattern is not None and not any(pattern[k]):
                assert isinstance(g.type, DisconnectedType), f"input {k} is declared disconnected but got a gradient"
                continue
            if isinstance(g.type, DisconnectedType):
                continue
            assert id(inp) not in grads, "the shim does not accumulate fan-out (not needed by the tests)"
            grads[id(inp)] = g
    return [grads.get(id(w)) for w in wrt]
