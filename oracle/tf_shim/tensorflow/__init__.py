"""TEST INFRASTRUCTURE (oracle) -- a numpy stand-in for the few TensorFlow-1.x graph-API calls the
reference's inference code makes, so that the reference's OWN UNMODIFIED files

    /root/reference/HM-16.5_Test_AI/bin/net_CNN.py
    /root/reference/HM-16.5_Test_AI/bin/video_to_cu_depth.py
    /root/reference/ETH-CNN_Training_LDP/net_CTU64.py

can be imported and executed in this container (TensorFlow is not installable: no network).
It is used ONLY by oracle/make_golden.py to generate tests/golden/*.npz and by the CPU tests that
re-check them when /root/reference is present.  Nothing in the product imports it.

Semantics restated from TensorFlow's public op definitions (float32 throughout):
  conv2d NHWC/HWIO, VALID           avg_pool (window == stride, divisible extents)
  resize_nearest_neighbor, align_corners=False  (src = floor(dst * in / out))
  leaky_relu(x) = maximum(0.2 * x, x)           sigmoid, relu, matmul, concat, reshape
  cond(pred, f, g): both branches traced at build time, one evaluated at run time
  Variable auto-naming: "Variable", "Variable_1", ... in creation order (names the checkpoints use)
  train.Saver(var_list).restore(sess, prefix): by-name lookup in a V2 bundle.
The graph is lazy (placeholders are fed at Session.run), values are memoised per run.
"""
from __future__ import annotations

import os as _os
import sys as _sys

import builtins as _builtins

import numpy as _np

_builtin_slice = _builtins.slice

float32 = _np.float32
_F = _np.float32

_here = _os.path.dirname(_os.path.abspath(__file__))
_repo = _os.path.abspath(_os.path.join(_here, "..", "..", ".."))
if _repo not in _sys.path:
    _sys.path.insert(0, _repo)
from oracle import tf_bundle as _tf_bundle  # noqa: E402

_TRAINABLE = []
_NAME_COUNTS = {}


def _reset_default_graph():
    del _TRAINABLE[:]
    _NAME_COUNTS.clear()
    _SCOPED_VARS.clear()
    del _SCOPE[:]


reset_default_graph = _reset_default_graph


def _as_tensor(v):
    if isinstance(v, Tensor):
        return v
    return Tensor(lambda env, _v=v: _np.asarray(_v, dtype=_F) if not isinstance(_v, (_builtins.bool, _np.bool_)) else _v, (),
                  "Const")


class Tensor(object):
    def __init__(self, fn, inputs, op, name=None, static_shape=None):
        self._fn = fn
        self._inputs = tuple(inputs)
        self.op_type = op
        self.name = name or op
        self.static_shape = static_shape

    def _eval(self, env):
        key = id(self)
        if key in env:
            return env[key]
        val = self._fn(env, *[i._eval(env) for i in self._inputs])
        env[key] = val
        return val

    def __repr__(self):
        return "<shim Tensor %s op=%s shape=%s>" % (self.name, self.op_type, self.static_shape)

    # arithmetic (float32, numpy broadcasting == TF broadcasting for the shapes used)
    def _bin(self, other, f, op, rev=False):
        o = _as_tensor(other)
        a, b = (o, self) if rev else (self, o)
        return Tensor(lambda env, x, y: f(x, y), (a, b), op)

    def __add__(self, o): return self._bin(o, lambda x, y: (x + y).astype(_F), "Add")
    def __radd__(self, o): return self._bin(o, lambda x, y: (x + y).astype(_F), "Add", True)
    def __sub__(self, o): return self._bin(o, lambda x, y: (x - y).astype(_F), "Sub")
    def __rsub__(self, o): return self._bin(o, lambda x, y: (x - y).astype(_F), "Sub", True)
    def __mul__(self, o): return self._bin(o, lambda x, y: (x * y).astype(_F), "Mul")
    def __rmul__(self, o): return self._bin(o, lambda x, y: (x * y).astype(_F), "Mul", True)
    def __truediv__(self, o): return self._bin(o, lambda x, y: (x / y).astype(_F), "RealDiv")
    def __rtruediv__(self, o): return self._bin(o, lambda x, y: (x / y).astype(_F), "RealDiv", True)
    __div__ = __truediv__
    def __neg__(self): return Tensor(lambda env, x: (-x).astype(_F), (self,), "Neg")
    def __lt__(self, o): return self._bin(o, lambda x, y: x < y, "Less")
    def __gt__(self, o): return self._bin(o, lambda x, y: x > y, "Greater")
    def __le__(self, o): return self._bin(o, lambda x, y: x <= y, "LessEqual")
    def __ge__(self, o): return self._bin(o, lambda x, y: x >= y, "GreaterEqual")
    def __getitem__(self, idx): return Tensor(lambda env, x: x[idx], (self,), "StridedSlice")
    __hash__ = object.__hash__


class Variable(Tensor):
    def __init__(self, initial_value, name=None, trainable=True):
        base = name or "Variable"
        n = _NAME_COUNTS.get(base, 0)
        _NAME_COUNTS[base] = n + 1
        uniq = base if n == 0 else "%s_%d" % (base, n)
        init = _as_tensor(initial_value)
        self._value = None
        self._init = init
        Tensor.__init__(self, lambda env: self._read(), (), "VariableV2", name=uniq + ":0")
        self.var_name = uniq
        if trainable:
            _TRAINABLE.append(self)

    def _read(self):
        if self._value is None:
            raise RuntimeError("variable %s used before restore/initialisation" % self.var_name)
        return self._value

    def load(self, value):
        self._value = _np.ascontiguousarray(value, dtype=_F)


def trainable_variables():
    return list(_TRAINABLE)


# ----------------------------------------------------------------------------- variable scopes / get_variable
_SCOPE = []
_SCOPED_VARS = {}


class _VarScope(object):
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        _SCOPE.append(self.name)
        return self

    def __exit__(self, *a):
        _SCOPE.pop()
        return False

    def reuse_variables(self):
        pass


def variable_scope(name_or_scope, reuse=None, **kw):
    return _VarScope(name_or_scope if isinstance(name_or_scope, str) else name_or_scope.name)


def get_variable_scope():
    return _VarScope("/".join(_SCOPE))


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True):
    full = "/".join(_SCOPE + [name])
    if full in _SCOPED_VARS:
        return _SCOPED_VARS[full]
    v = Variable.__new__(Variable)
    v._value = None
    v._init = constant(0.0, shape=shape)
    Tensor.__init__(v, lambda env, _v=v: _v._read(), (), "VariableV2", name=full + ":0", static_shape=shape)
    v.var_name = full
    _SCOPED_VARS[full] = v
    if trainable:
        _TRAINABLE.append(v)
    return v


def placeholder(dtype, shape=None, name=None):
    t = Tensor(None, (), "Placeholder", name=name, static_shape=shape)

    def fn(env, _t=t):
        raise RuntimeError("placeholder %r was not fed" % (_t,))
    t._fn = fn
    return t


def constant(value, dtype=None, shape=None, name=None):
    arr = _np.asarray(value, dtype=_F)
    if shape is not None:
        arr = _np.full(tuple(shape), arr, dtype=_F) if arr.ndim == 0 else arr.reshape(shape)
    return Tensor(lambda env, _a=arr: _a, (), "Const")


def truncated_normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None, name=None):
    shp = tuple(shape)
    return Tensor(lambda env: _np.clip(_np.random.standard_normal(shp), -2, 2).astype(_F) * _F(stddev) + _F(mean),
                  (), "TruncatedNormal")


def zeros(shape, dtype=None, name=None):
    parts = [_as_tensor(s) for s in shape] if isinstance(shape, (list, tuple)) else [_as_tensor(shape)]
    return Tensor(lambda env, *d: _np.zeros(tuple(int(x) for x in d), dtype=_F), parts, "Zeros")


def shape(x, name=None):
    return Tensor(lambda env, v: _np.asarray(v.shape, dtype=_np.int64), (x,), "Shape")


def scalar_mul(scalar, x):
    # tf.scalar_mul(scalar, x) == scalar * x with the scalar converted to x's dtype
    s = _F(scalar)
    return Tensor(lambda env, v: (s * v).astype(_F), (x,), "Mul")


def reshape(x, shape_, name=None):
    shp = tuple(int(s) for s in shape_)
    return Tensor(lambda env, v: v.reshape(shp), (_as_tensor(x),), "Reshape")


def concat(values, axis, name=None):
    return Tensor(lambda env, *v: _np.concatenate(v, axis=axis).astype(_F), [_as_tensor(t) for t in values], "ConcatV2")


def matmul(a, b, name=None):
    return Tensor(lambda env, x, y: _np.matmul(x, y).astype(_F), (a, b), "MatMul")


def multiply(a, b, name=None):
    if isinstance(b, (list, tuple)):   # tf.multiply(1.0, [t0, t1, ...]) packs the list into one tensor first
        b = stack(list(b), axis=0)
    return _as_tensor(a) * b


def slice(input_, begin, size, name=None):  # noqa: A001
    def fn(env, v):
        idx = tuple(_builtin_slice(b, None if sz == -1 else b + sz) for b, sz in zip(begin, size))
        return v[idx]
    return Tensor(fn, (_as_tensor(input_),), "Slice")


def split(value, num_or_size_splits, axis=0, name=None):
    n = int(num_or_size_splits)
    return [Tensor(lambda env, v, _i=i: _np.split(v, n, axis=axis)[_i], (_as_tensor(value),), "Split") for i in range(n)]


def one_hot(indices, depth, axis=-1, dtype=None, name=None):
    def fn(env, v):
        idx = _np.asarray(v).astype(_np.int64)
        out = (idx[..., None] == _np.arange(depth)).astype(_F)
        if axis not in (-1, idx.ndim):
            out = _np.moveaxis(out, -1, axis)
        return out
    return Tensor(fn, (_as_tensor(indices),), "OneHot")


def to_int32(x, name=None):
    return Tensor(lambda env, v: _np.asarray(v).astype(_np.int32), (_as_tensor(x),), "Cast")


bool = "bool"  # noqa: A001  (tf.bool, only ever passed to tf.cast)


def cond(pred, true_fn=None, false_fn=None, name=None, fn1=None, fn2=None):
    t_branch = _as_tensor((true_fn or fn1)())
    f_branch = _as_tensor((false_fn or fn2)())
    p = _as_tensor(pred)

    def fn(env):
        return t_branch._eval(env) if _builtins.bool(p._eval(env)) else f_branch._eval(env)
    return Tensor(fn, (), "Cond")


def count_nonzero(x, axis=None, name=None):
    return Tensor(lambda env, v: _np.int64(_np.count_nonzero(v)), (_as_tensor(x),), "CountNonzero")


def _lazy_unary(f, op):
    def g(x, *a, **k):
        return Tensor(lambda env, v: f(v), (_as_tensor(x),), op)
    return g


log = _lazy_unary(lambda v: _np.log(v).astype(_F), "Log")
round = _lazy_unary(lambda v: _np.round(v).astype(_F), "Round")  # noqa: A001
to_float = _lazy_unary(lambda v: _np.asarray(v, dtype=_F), "Cast")


def cast(x, dtype=None, name=None):
    if dtype == "bool":
        return Tensor(lambda env, v: _np.asarray(v).astype(_np.bool_), (_as_tensor(x),), "Cast")
    return Tensor(lambda env, v: _np.asarray(v, dtype=_F), (_as_tensor(x),), "Cast")


def equal(a, b, name=None):
    return _as_tensor(a)._bin(b, lambda x, y: x == y, "Equal")


def reduce_sum(x, axis=None, keep_dims=False, name=None):
    return Tensor(lambda env, v: _np.sum(v, axis=None if axis is None else tuple(_np.atleast_1d(axis)),
                                          keepdims=keep_dims, dtype=_F), (_as_tensor(x),), "Sum")


def reduce_mean(x, axis=None, keep_dims=False, name=None):
    return Tensor(lambda env, v: _np.mean(_np.asarray(v, dtype=_F), axis=None if axis is None else tuple(_np.atleast_1d(axis)),
                                           keepdims=keep_dims, dtype=_F), (_as_tensor(x),), "Mean")


def tile(x, multiples, name=None):
    return Tensor(lambda env, v: _np.tile(v, multiples), (_as_tensor(x),), "Tile")


def stack(values, axis=0, name=None):
    return Tensor(lambda env, *v: _np.stack(v, axis=axis), [_as_tensor(t) for t in values], "Pack")


# ----------------------------------------------------------------------------- tf.nn
class _NN(object):
    @staticmethod
    def conv2d(input, filter, strides, padding, use_cudnn_on_gpu=None, data_format=None, name=None):  # noqa: A002
        if padding != "VALID":
            raise NotImplementedError("shim conv2d: only VALID is used by the reference")
        sh, sw = int(strides[1]), int(strides[2])

        def fn(env, x, w):
            b, h, ww, cin = x.shape
            kh, kw, wcin, cout = w.shape
            assert wcin == cin
            oh, ow = (h - kh) // sh + 1, (ww - kw) // sw + 1
            s0, s1, s2, s3 = x.strides
            win = _np.lib.stride_tricks.as_strided(
                x, shape=(b, oh, ow, kh, kw, cin), strides=(s0, s1 * sh, s2 * sw, s1, s2, s3), writeable=False)
            out = _np.matmul(_np.ascontiguousarray(win).reshape(b * oh * ow, kh * kw * cin), w.reshape(kh * kw * cin, cout))
            return out.reshape(b, oh, ow, cout).astype(_F)
        return Tensor(fn, (_as_tensor(input), _as_tensor(filter)), "Conv2D")

    @staticmethod
    def avg_pool(value, ksize, strides, padding, data_format="NHWC", name=None):
        kh, kw = int(ksize[1]), int(ksize[2])
        if (int(strides[1]), int(strides[2])) != (kh, kw):
            raise NotImplementedError("shim avg_pool: window must equal stride")

        def fn(env, x):
            b, h, w, c = x.shape
            if h % kh or w % kw:
                raise NotImplementedError("shim avg_pool: extents must divide")
            s = x.reshape(b, h // kh, kh, w // kw, kw, c).sum(axis=(2, 4), dtype=_F)
            return (s / _F(kh * kw)).astype(_F)
        return Tensor(fn, (_as_tensor(value),), "AvgPool")

    @staticmethod
    def max_pool(value, ksize, strides, padding, data_format="NHWC", name=None):
        raise NotImplementedError("shim: max_pool is dead code in the reference's forward path")

    @staticmethod
    def leaky_relu(features, alpha=0.2, name=None):
        a = _F(alpha)
        return Tensor(lambda env, x: _np.maximum(a * x, x).astype(_F), (_as_tensor(features),), "LeakyRelu")

    @staticmethod
    def relu(features, name=None):
        return Tensor(lambda env, x: _np.maximum(x, _F(0)).astype(_F), (_as_tensor(features),), "Relu")

    @staticmethod
    def sigmoid(x, name=None):
        return Tensor(lambda env, v: (_F(1) / (_F(1) + _np.exp(-v, dtype=_F))).astype(_F), (_as_tensor(x),), "Sigmoid")

    @staticmethod
    def tanh(x, name=None):
        return Tensor(lambda env, v: _np.tanh(v).astype(_F), (_as_tensor(x),), "Tanh")

    @staticmethod
    def dropout(x, keep_prob, noise_shape=None, seed=None, name=None):
        def fn(env, v, kp):
            raise RuntimeError("shim: dropout branch must never run at inference (isdrop=0)")
        return Tensor(fn, (_as_tensor(x), _as_tensor(keep_prob)), "Dropout")

    @staticmethod
    def l2_loss(t, name=None):
        return Tensor(lambda env, v: _F(0.5) * _np.sum(v * v, dtype=_F), (_as_tensor(t),), "L2Loss")


nn = _NN()
sigmoid = _NN.sigmoid
tanh = _NN.tanh


# ----------------------------------------------------------------------------- tf.image
class _Image(object):
    @staticmethod
    def resize_nearest_neighbor(images, size, align_corners=False, name=None):
        if align_corners:
            raise NotImplementedError
        oh, ow = int(size[0]), int(size[1])

        def fn(env, x):
            b, h, w, c = x.shape
            # legacy kernel (align_corners=False, half_pixel_centers=False): src = floor(dst * in/out)
            iy = _np.minimum(_np.floor(_np.arange(oh) * (_F(h) / _F(oh))).astype(_np.int64), h - 1)
            ix = _np.minimum(_np.floor(_np.arange(ow) * (_F(w) / _F(ow))).astype(_np.int64), w - 1)
            return x[:, iy][:, :, ix]
        return Tensor(fn, (_as_tensor(images),), "ResizeNearestNeighbor")


image = _Image()


# ----------------------------------------------------------------------------- tf.train / Session
# ----------------------------------------------------------------------------- tf.contrib.rnn (one-step use only)
class _LSTMStateTuple(tuple):
    def __new__(cls, c, h):
        return tuple.__new__(cls, (c, h))

    c = property(lambda self: self[0])
    h = property(lambda self: self[1])


class _LSTMCell(object):
    """tf.contrib.rnn.LSTMCell (TF 1.x rnn_cell_impl.LSTMCell.call) without peepholes / projection:
    gates (i, j, f, o) = split([x, h_prev] kernel + bias); c = sigmoid(f + forget_bias) c_prev + sigmoid(i) tanh(j),
    clipped to +-cell_clip; h = sigmoid(o) tanh(c).  Variables: <scope>/lstm_cell/{kernel,bias}."""

    def __init__(self, num_units, forget_bias=1.0, cell_clip=None, state_is_tuple=True, **kw):
        self.n, self.forget_bias, self.cell_clip = int(num_units), _F(forget_bias), cell_clip

    def __call__(self, inputs, state):
        c_prev, h_prev = state
        n = self.n
        with variable_scope("lstm_cell"):
            in_dim = None
            kernel_holder = {}

            def fn(env, x, c0, h0):
                k = kernel_holder["k"]._eval(env)
                b = kernel_holder["b"]._eval(env)
                z = (_np.concatenate([x, h0], axis=1).astype(_F) @ k + b).astype(_F)
                i, j, f, o = _np.split(z, 4, axis=1)
                sig = lambda v: (_F(1) / (_F(1) + _np.exp(-v, dtype=_F))).astype(_F)
                c = (sig(f + self.forget_bias) * c0 + sig(i) * _np.tanh(j)).astype(_F)
                if self.cell_clip is not None:
                    c = _np.clip(c, _F(-self.cell_clip), _F(self.cell_clip))
                h = (sig(o) * _np.tanh(c)).astype(_F)
                return (c, h)
            # input width is only known from the checkpoint: kernel is [in + n, 4n]
            kernel_holder["k"] = get_variable("kernel", None)
            kernel_holder["b"] = get_variable("bias", [4 * n])
            both = Tensor(fn, (_as_tensor(inputs), _as_tensor(c_prev), _as_tensor(h_prev)), "LSTMCell")
        c_new = Tensor(lambda env, v: v[0], (both,), "LSTMCell_c")
        h_new = Tensor(lambda env, v: v[1], (both,), "LSTMCell_h")
        return h_new, _LSTMStateTuple(c_new, h_new)


class _DropoutWrapper(object):
    def __init__(self, cell, input_keep_prob=1.0, output_keep_prob=1.0, **kw):
        self.cell, self.keep = cell, _as_tensor(output_keep_prob)

    def __call__(self, inputs, state):
        out, new_state = self.cell(inputs, state)
        keep = self.keep

        def fn(env, v, kp):
            if float(kp) != 1.0:
                raise RuntimeError("shim: dropout must be inactive at inference (isdrop=0)")
            return v   # x / 1 * floor(1 + uniform[0,1)) == x
        return Tensor(fn, (out, keep), "Dropout"), new_state


class _MultiRNNCell(object):
    def __init__(self, cells, state_is_tuple=True):
        self.cells = list(cells)

    def __call__(self, inputs, state):
        cur = inputs
        new_states = []
        with variable_scope("multi_rnn_cell"):
            for i, cell in enumerate(self.cells):
                with variable_scope("cell_%d" % i):
                    cur, st = cell(cur, state[i])
                    new_states.append(st)
        return cur, tuple(new_states)


class _Rnn(object):
    LSTMCell = _LSTMCell
    DropoutWrapper = _DropoutWrapper
    MultiRNNCell = _MultiRNNCell
    LSTMStateTuple = _LSTMStateTuple


class _Contrib(object):
    rnn = _Rnn()


contrib = _Contrib()


class _SaverDef(object):
    V1 = 1
    V2 = 2


class _Saver(object):
    def __init__(self, var_list=None, write_version=None, max_to_keep=None, **kw):
        self._vars = list(var_list) if var_list is not None else list(_TRAINABLE)

    def restore(self, sess, save_path):
        bundle = _tf_bundle.read_bundle(save_path, verify_crc=True)
        for v in self._vars:
            if v.var_name not in bundle:
                raise KeyError("tensor %s not found in checkpoint %s" % (v.var_name, save_path))
            arr = bundle[v.var_name]
            v.load(arr)

    def save(self, *a, **k):
        raise NotImplementedError


class _Optimizer(object):
    def __init__(self, *a, **k):
        pass

    def minimize(self, loss, var_list=None, **k):
        return Tensor(lambda env: None, (), "NoOp")


class _Train(object):
    Saver = _Saver
    SaverDef = _SaverDef
    MomentumOptimizer = _Optimizer
    GradientDescentOptimizer = _Optimizer
    AdamOptimizer = _Optimizer

    @staticmethod
    def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False, name=None):
        return _as_tensor(learning_rate)


train = _Train()


class Session(object):
    def __init__(self, *a, **k):
        pass

    def run(self, fetches, feed_dict=None):
        env = {}
        for ph, val in (feed_dict or {}).items():
            env[id(ph)] = _np.asarray(val, dtype=_F)
        if isinstance(fetches, (list, tuple)):
            return [f._eval(env) for f in fetches]
        return fetches._eval(env)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def global_variables_initializer():
    def fn(env):
        for v in _TRAINABLE:
            v.load(v._init._eval({}))
    return Tensor(fn, (), "NoOp")


__version__ = "1.x-numpy-shim"
