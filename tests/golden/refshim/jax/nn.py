import torch as _t


def silu(x): return x * _t.sigmoid(x)
def sigmoid(x): return _t.sigmoid(x)
