"""A minimal TensorFlow-1.15 / TFP-0.8 API shim on torch float64 -- TEST INFRASTRUCTURE ONLY.

Purpose: execute the UNMODIFIED reference source (/root/reference/SVGPVAE_model.py, utils.py) in this
container, where TensorFlow cannot be installed (Python 3.12, no wheel, no network), in order to
generate golden vectors that pin oracle/ to the reference's own code (tests/golden/make_reference_golden.py).
Only the ~40 tf.* symbols the SVGP path touches are provided; each follows the documented TF 1.15
semantics of that op (eager evaluation instead of graph construction -- the ops are pure functions, so
the values are the same).  Every floating dtype maps to torch.float64: the vectors are a high-precision
execution of the reference's op sequence, which is what an fp64 oracle has to match.  Because tensors are
torch tensors, torch autograd differentiates THROUGH the reference code, which stands in for tf.gradients.

What this does NOT pin: the arithmetic inside TensorFlow / TFP themselves.  The three
tfp.math.psd_kernels used by the reference are restated below from the TFP 0.8.0 definitions
(exponentiated_quadratic.py, exp_sin_squared.py, linear.py: _apply with feature_ndims = 1).
"""
import builtins
import sys
import types

import numpy as np
import torch

F64 = torch.float64


class _DType:
    def __init__(self, name, torch_dtype):
        self.name, self.torch = name, torch_dtype

    def __repr__(self):
        return "tf." + self.name


float32, float64 = _DType("float32", F64), _DType("float64", F64)
int32, int64 = _DType("int32", torch.int64), _DType("int64", torch.int64)
newaxis = None


def _dt(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, _DType):
        return dtype.torch
    if isinstance(dtype, torch.dtype):
        return F64 if dtype.is_floating_point else torch.int64
    return F64 if np.issubdtype(np.dtype(dtype), np.floating) else torch.int64


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, (torch.Size, tuple, list)) and all(isinstance(v, (int, np.integer)) for v in x):
        t = torch.tensor(list(x))
    else:
        t = torch.as_tensor(np.asarray(x))
    want = _dt(dtype) if dtype is not None else (F64 if t.dtype.is_floating_point else t.dtype)
    return t.to(want)


# torch tensors answer the two TF tensor methods the reference calls
torch.Tensor.get_shape = lambda self: self.shape


def constant(value, dtype=None, name=None):
    return _t(value, dtype).detach().clone()


VARIABLES = {}


def Variable(initial_value=None, name=None, dtype=None, trainable=True):
    v = _t(initial_value, dtype).detach().clone().requires_grad_(True)
    VARIABLES[name if name is not None else "var%d" % len(VARIABLES)] = v
    return v


def cast(x, dtype=None):
    return _t(x, dtype)


def shape(x):
    return tuple(x.shape)


def log(x):
    return torch.log(_t(x))


def exp(x):
    return torch.exp(_t(x))


def sqrt(x):
    return torch.sqrt(_t(x))


def matmul(a, b, transpose_a=False, transpose_b=False):
    a = a.transpose(-1, -2) if transpose_a else a
    b = b.transpose(-1, -2) if transpose_b else b
    return torch.matmul(a, b)


def multiply(a, b):
    return a * b


def expand_dims(x, axis):
    return torch.unsqueeze(x, axis)


def reduce_sum(x, axis=None, keepdims=False):
    if isinstance(x, (list, tuple)):
        x = torch.stack([_t(v) for v in x])
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def reduce_mean(x, axis=None, keepdims=False):
    return x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims)


def transpose(x, perm=None):
    return x.permute(*perm) if perm is not None else x.permute(*reversed(builtins.range(x.dim())))


def trace(x):
    return torch.diagonal(x, dim1=-2, dim2=-1).sum(-1)


def gather(params, indices):
    return params[indices]


def repeat(x, repeats, axis):
    r = torch.as_tensor(repeats).reshape(-1).long()
    if r.numel() == 1:
        return torch.repeat_interleave(x, int(r[0]), dim=axis)
    return torch.repeat_interleave(x, r, dim=axis)


def segment_mean(data, segment_ids):
    ids = torch.as_tensor(segment_ids).long()
    n = int(ids.max()) + 1
    sums = torch.zeros((n,) + tuple(data.shape[1:]), dtype=data.dtype).index_add(0, ids, data)
    return sums / torch.bincount(ids, minlength=n).to(data.dtype).reshape((n,) + (1,) * (data.dim() - 1))


def concat(values, axis=0):
    return torch.cat(list(values), dim=axis)


def stack(values, axis=0):
    return torch.stack(list(values), dim=axis)


def clip_by_value(x, lo, hi):
    return torch.clamp(x, lo, hi)          # gradient 0 outside [lo, hi], as tf.clip_by_value


def stop_gradient(x):
    return x.detach()


def _tf_range(n, dtype=None):
    return torch.arange(int(n)).to(_dt(dtype) or torch.int64)


def tile(x, multiples):
    return x.repeat(*[int(m) for m in multiples])


class _Linalg(types.ModuleType):
    @staticmethod
    def inv(x):
        return torch.linalg.inv(x)                     # LU-based explicit inverse, like tf.linalg.inv

    @staticmethod
    def cholesky(x):
        return torch.linalg.cholesky(x)

    @staticmethod
    def matvec(a, b):
        return torch.matmul(a, b.unsqueeze(-1)).squeeze(-1)       # broadcasts batch dimensions

    @staticmethod
    def diag_part(x):
        return torch.diagonal(x, dim1=-2, dim2=-1)

    @staticmethod
    def diag(x):
        return torch.diag_embed(x)

    @staticmethod
    def set_diag(x, d):
        return x - torch.diag_embed(torch.diagonal(x, dim1=-2, dim2=-1)) + torch.diag_embed(d)

    trace = staticmethod(trace)


class _Math(types.ModuleType):
    @staticmethod
    def reciprocal_no_nan(x):
        safe = torch.where(x == 0, torch.ones_like(x), x)
        return torch.where(x == 0, torch.zeros_like(x), 1.0 / safe)

    @staticmethod
    def reduce_euclidean_norm(x, axis=None, keepdims=False):
        return torch.sqrt((x * x).sum(dim=axis, keepdim=keepdims))

    log = staticmethod(log)
    multiply = staticmethod(multiply)


class _Random(types.ModuleType):
    @staticmethod
    def normal(shape, dtype=None, seed=None):
        return torch.randn(*shape, dtype=F64, generator=torch.Generator().manual_seed(0))


# ---------------------------------------------------------------------------------------------
# tfp.math.psd_kernels 0.8.0, feature_ndims = 1
# ---------------------------------------------------------------------------------------------
def _pair(x1, x2):
    return x1.unsqueeze(-2), x2.unsqueeze(-3)          # (..., e1, 1, f), (..., 1, e2, f)


class _Kernel:
    def matrix(self, x1, x2):
        a, b = _pair(_t(x1), _t(x2))
        return self._apply(a, b)

    def apply(self, x1, x2):
        return self._apply(_t(x1), _t(x2))


class ExponentiatedQuadratic(_Kernel):
    """exp(-||x - y||^2 / (2 l^2)) * amplitude^2; amplitude / length_scale None -> 1."""

    def __init__(self, amplitude=None, length_scale=None, feature_ndims=1, name=None):
        self.amplitude, self.length_scale = amplitude, length_scale

    def _apply(self, a, b):
        e = -0.5 * ((a - b) ** 2).sum(-1)
        if self.length_scale is not None:
            e = e / _t(self.length_scale) ** 2
        if self.amplitude is not None:
            e = e + 2.0 * torch.log(_t(self.amplitude))
        return torch.exp(e)


class ExpSinSquared(_Kernel):
    """amplitude^2 exp(-2 sum_f sin^2(pi |x - y| / period) / l^2)."""

    def __init__(self, amplitude=None, length_scale=None, period=None, feature_ndims=1, name=None):
        self.amplitude, self.length_scale, self.period = amplitude, length_scale, period

    def _apply(self, a, b):
        d = np.pi * torch.abs(a - b)
        if self.period is not None:
            d = d / _t(self.period)
        e = -2.0 * (torch.sin(d) ** 2).sum(-1)
        if self.length_scale is not None:
            e = e / _t(self.length_scale) ** 2
        if self.amplitude is not None:
            e = e + 2.0 * torch.log(_t(self.amplitude))
        return torch.exp(e)


class Linear(_Kernel):
    """sum_f x_f y_f (bias_variance, slope_variance, shift all None)."""

    def __init__(self, bias_variance=None, slope_variance=None, shift=None, feature_ndims=1, name=None):
        assert bias_variance is None and slope_variance is None and shift is None

    def _apply(self, a, b):
        return (a * b).sum(-1)


class _Anything:
    """Stands in for every symbol the SVGP path never executes (keras layers, matplotlib, ...)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __mro_entries__(self, bases):
        return (object,)


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


def install():
    """Register the shim modules in sys.modules (call before importing the reference)."""
    tf = _Stub("tensorflow")
    g = globals()
    for k in ("float32", "float64", "int32", "int64", "newaxis", "constant", "Variable", "cast", "shape", "log", "exp", "sqrt",
              "matmul", "multiply", "expand_dims", "reduce_sum", "reduce_mean", "transpose", "trace", "gather", "repeat", "stack",
              "clip_by_value", "stop_gradient", "tile", "segment_mean", "concat"):
        setattr(tf, k, g[k])
    tf.range = _tf_range
    tf.linalg, tf.math, tf.random = _Linalg("tensorflow.linalg"), _Math("tensorflow.math"), _Random("tensorflow.random")
    tfp = _Stub("tensorflow_probability")
    tfp.math = _Stub("tensorflow_probability.math")
    pk = types.ModuleType("tensorflow_probability.math.psd_kernels")
    pk.ExponentiatedQuadratic, pk.ExpSinSquared, pk.Linear = ExponentiatedQuadratic, ExpSinSquared, Linear
    tfp.math.psd_kernels = pk
    mods = {"tensorflow": tf, "tensorflow_probability": tfp, "tensorflow.python": _Stub("tensorflow.python"),
            "tensorflow.python.ops": _Stub("tensorflow.python.ops")}
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "seaborn"):
        mods[name] = _Stub(name)
    sys.modules.update(mods)
    return tf
