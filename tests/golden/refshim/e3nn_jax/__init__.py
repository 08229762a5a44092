"""e3nn_jax stand-in: scatter_sum only."""
import torch as _t


def scatter_sum(data, *, dst, output_size):
    out = _t.zeros((output_size,) + tuple(data.shape[1:]), dtype=data.dtype)
    return out.index_add(0, dst.long(), data)
