"""distrax stand-in: Distribution, ScalarAffine, Lambda, Transformed with distrax's published semantics
(Transformed.log_prob(y) = base.log_prob(inverse(y)) + inverse_log_det_jacobian(y); samples = forward(base samples))."""
import math as _math

import torch as _t


class Distribution:
    def _sample_n(self, key, n):
        raise NotImplementedError

    def log_prob(self, value):
        raise NotImplementedError

    def _sample_n_and_log_prob(self, key, n):
        samples = self._sample_n(key, n)
        return samples, self.log_prob(samples)

    def sample(self, *, seed, sample_shape=()):
        shape = (sample_shape,) if isinstance(sample_shape, int) else tuple(sample_shape)
        n = int(_math.prod(shape)) if shape else 1
        samples = self._sample_n(seed, n)
        return samples.reshape(shape + tuple(samples.shape[1:]))

    def sample_and_log_prob(self, *, seed, sample_shape=()):
        shape = (sample_shape,) if isinstance(sample_shape, int) else tuple(sample_shape)
        n = int(_math.prod(shape)) if shape else 1
        samples, lp = self._sample_n_and_log_prob(seed, n)
        return samples.reshape(shape + tuple(samples.shape[1:])), lp.reshape(shape)


class ScalarAffine:
    def __init__(self, shift, scale):
        self.shift, self.scale = shift, scale

    def forward(self, x): return self.scale * x + self.shift
    def inverse(self, y): return (y - self.shift) / self.scale
    def forward_log_det_jacobian(self, x): return _t.log(_t.abs(self.scale)) + _t.zeros_like(x)
    def inverse_log_det_jacobian(self, y): return -_t.log(_t.abs(self.scale)) + _t.zeros_like(y)


class Lambda:
    def __init__(self, forward, inverse, forward_log_det_jacobian, inverse_log_det_jacobian, event_ndims_in=0, event_ndims_out=0):
        self.forward, self.inverse = forward, inverse
        self.forward_log_det_jacobian, self.inverse_log_det_jacobian = forward_log_det_jacobian, inverse_log_det_jacobian


class Transformed(Distribution):
    def __init__(self, distribution, bijector):
        self.distribution, self.bijector = distribution, bijector

    def _sample_n(self, key, n):
        return self.bijector.forward(self.distribution._sample_n(key, n))

    def _sample_n_and_log_prob(self, key, n):
        x, lp_x = self.distribution._sample_n_and_log_prob(key, n)
        return self.bijector.forward(x), lp_x - self.bijector.forward_log_det_jacobian(x)

    def log_prob(self, value):
        x = self.bijector.inverse(value)
        return self.distribution.log_prob(x) + self.bijector.inverse_log_det_jacobian(value)
