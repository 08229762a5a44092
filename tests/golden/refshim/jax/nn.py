import torch as _t


def silu(x): return x * _t.sigmoid(x)
def sigmoid(x): return _t.sigmoid(x)


def logsumexp(a, b=None, axis=None):
    m = _t.max(a) if axis is None else _t.amax(a, dim=axis, keepdim=True)
    e = _t.exp(a - m)
    if b is not None:
        e = e * b
    s = _t.sum(e) if axis is None else _t.sum(e, dim=axis)
    return _t.log(s) + (m if axis is None else _t.squeeze(m, dim=axis))


def softmax(x, axis=-1): return _t.softmax(x, dim=axis)
