"""NumPy-backed stand-ins for the two third-party modules the reference imports but that cannot
be installed here (``mxnet-cu90==1.3.0.post0`` and the un-pinned ``midi`` / python-midi).

Used ONLY by tests/golden/make_golden.py, in the build container, to execute the reference's own
source files (MIDIUtil/*.py, VarAutoEncoder/{data,loss,transformer,model}.py) and record golden
vectors.  It states — as narrowly as possible — the MXNet 1.3 operator semantics the reference
relies on (Dense = x W^T + b, LayerNorm eps 1e-5 over the last axis, softmax default axis -1,
NumPy-style broadcasting with left-padded dims, NDArray advanced indexing, fused-LSTM gate order
i,f,g,o).  Nothing here is product code and nothing at run time on the GPU box imports it.
"""
import contextlib
import sys
import types

import numpy as np


# ------------------------------------------------------------------ NDArray
class NDArray:
    __array_priority__ = 100.0

    def __init__(self, a):
        self.a = np.asarray(a, dtype=np.float32) if not isinstance(a, np.ndarray) or a.dtype != np.float64 \
            else a.astype(np.float32)
        if self.a.dtype != np.float32 and self.a.dtype.kind == "f":
            self.a = self.a.astype(np.float32)

    # -- basics
    @property
    def shape(self):
        return tuple(self.a.shape)

    @property
    def size(self):
        return int(self.a.size)

    @property
    def dtype(self):
        return self.a.dtype

    @property
    def ndim(self):
        return self.a.ndim

    def asnumpy(self):
        return np.array(self.a)

    def asscalar(self):
        assert self.a.size == 1
        return self.a.reshape(-1)[0]

    def as_in_context(self, ctx):
        return self

    def astype(self, dt):
        return NDArray(self.a.astype(dt))

    def copy(self):
        return NDArray(self.a.copy())

    def __len__(self):
        return self.a.shape[0]

    def __repr__(self):
        return "NDArray(%r)" % (self.a,)

    # -- arithmetic (MXNet: elementwise, broadcast when shapes differ)
    @staticmethod
    def _v(o):
        return o.a if isinstance(o, NDArray) else o

    def __add__(self, o): return NDArray(self.a + self._v(o))
    __radd__ = __add__
    def __sub__(self, o): return NDArray(self.a - self._v(o))
    def __rsub__(self, o): return NDArray(self._v(o) - self.a)
    def __mul__(self, o): return NDArray(self.a * self._v(o))
    __rmul__ = __mul__
    def __truediv__(self, o): return NDArray(self.a / self._v(o))
    def __rtruediv__(self, o): return NDArray(self._v(o) / self.a)
    def __neg__(self): return NDArray(-self.a)
    def __iadd__(self, o):
        self.a = (self.a + self._v(o)).astype(np.float32)
        return self
    # comparisons return 0/1 float arrays, as MXNet does
    def __eq__(self, o): return NDArray((self.a == self._v(o)).astype(np.float32))
    def __ne__(self, o): return NDArray((self.a != self._v(o)).astype(np.float32))
    def __gt__(self, o): return NDArray((self.a > self._v(o)).astype(np.float32))
    def __lt__(self, o): return NDArray((self.a < self._v(o)).astype(np.float32))
    __hash__ = None

    # -- indexing (NumPy semantics; NDArray indices are cast to int32 as MXNet does)
    @staticmethod
    def _key(k):
        if isinstance(k, NDArray):
            return k.a.astype(np.int32)
        if isinstance(k, tuple):
            return tuple(NDArray._key(x) for x in k)
        return k

    def __getitem__(self, k):
        return NDArray(self.a[self._key(k)])

    def __setitem__(self, k, v):
        self.a[self._key(k)] = self._v(v)

    # -- shape ops
    def reshape(self, *shape, **kw):
        shp = kw.get("shape", shape[0] if len(shape) == 1 and isinstance(shape[0], (list, tuple)) else shape)
        out, src = [], list(self.a.shape)
        for i, s in enumerate(shp):
            out.append(src[i] if s == 0 else s)          # MXNet: 0 copies the input dim
        return NDArray(self.a.reshape(out))

    def swapaxes(self, a, b):
        return NDArray(np.swapaxes(self.a, a, b))

    def expand_dims(self, axis):
        return NDArray(np.expand_dims(self.a, axis))

    def squeeze(self, axis=None):
        return NDArray(np.squeeze(self.a, axis=axis))

    def _red_axes(self, axis, exclude):
        if axis is None:
            return None
        ax = (axis,) if isinstance(axis, int) else tuple(axis)
        ax = tuple(a % self.a.ndim for a in ax)
        if exclude:
            ax = tuple(i for i in range(self.a.ndim) if i not in ax)
        return ax

    def sum(self, axis=None, exclude=False, keepdims=False):
        return NDArray(self.a.sum(axis=self._red_axes(axis, exclude), keepdims=keepdims, dtype=np.float32))

    def mean(self, axis=None, exclude=False, keepdims=False):
        return NDArray(self.a.mean(axis=self._red_axes(axis, exclude), keepdims=keepdims, dtype=np.float32))

    def norm(self):
        return NDArray(np.sqrt((self.a.astype(np.float32) ** 2).sum()))

    def take(self, idx, axis=0):
        return NDArray(np.take(self.a, NDArray._key(idx), axis=axis))


def _w(x):
    return x if isinstance(x, NDArray) else NDArray(x)


def _a(x):
    return x.a if isinstance(x, NDArray) else np.asarray(x, dtype=np.float32)


# ------------------------------------------------------------------ mx.nd
nd = types.ModuleType("mxnet.nd")
nd.NDArray = NDArray
nd.ndarray = NDArray          # transformer.py annotates with mx.nd.ndarray
nd.array = lambda a, ctx=None, dtype=None: NDArray(np.array(_a(a), dtype=np.float32))
nd.full = lambda shape, val, ctx=None, dtype=None: NDArray(np.full(shape, val, dtype=np.float32))
nd.zeros = lambda shape, ctx=None, dtype=None: NDArray(np.zeros(shape, dtype=np.float32))
nd.ones = lambda shape, ctx=None, dtype=None: NDArray(np.ones(shape, dtype=np.float32))
nd.ones_like = lambda x: NDArray(np.ones_like(_a(x)))
nd.zeros_like = lambda x: NDArray(np.zeros_like(_a(x)))
nd.where = lambda c, x, y: NDArray(np.where(_a(c) != 0, _a(x), _a(y)))
nd.sqrt = lambda x: NDArray(np.sqrt(_a(x)))
nd.log = lambda x: NDArray(np.log(_a(x)))
nd.exp = lambda x: NDArray(np.exp(_a(x)))
nd.sigmoid = lambda x: NDArray((1.0 / (1.0 + np.exp(-_a(x).astype(np.float32)))).astype(np.float32))
nd.relu = lambda x: NDArray(np.maximum(_a(x), 0))
nd.tanh = lambda x: NDArray(np.tanh(_a(x)))
nd.expand_dims = lambda x, axis: NDArray(np.expand_dims(_a(x), axis))
nd.squeeze = lambda x, axis=None: NDArray(np.squeeze(_a(x), axis=axis))
nd.repeat = lambda x, repeats, axis=None: NDArray(np.repeat(_a(x), repeats, axis=axis))
nd.broadcast_add = lambda x, y: NDArray(_a(x) + _a(y))
nd.broadcast_mul = lambda x, y: NDArray(_a(x) * _a(y))
nd.broadcast_like = lambda x, y: NDArray(np.broadcast_to(_a(x), _a(y).shape).copy())
nd.broadcast_logical_or = lambda x, y: NDArray(((_a(x) != 0) | (_a(y) != 0)).astype(np.float32))
nd.sum = lambda x, axis=None, exclude=False, keepdims=False: _w(x).sum(axis, exclude, keepdims)
nd.mean = lambda x, axis=None, exclude=False, keepdims=False: _w(x).mean(axis, exclude, keepdims)
nd.max = lambda x, axis=None: NDArray(np.max(_a(x), axis=axis))
nd.reshape = lambda x, shape: _w(x).reshape(shape=shape)
nd.concat = lambda *xs, dim=1: NDArray(np.concatenate([_a(x) for x in xs], axis=dim))
nd.concatenate = lambda xs, axis=0: NDArray(np.concatenate([_a(x) for x in xs], axis=axis))


def _softmax(x, axis=-1):
    a = _a(x)
    m = a.max(axis=axis, keepdims=True)
    e = np.exp(a - m)
    return NDArray(e / e.sum(axis=axis, keepdims=True, dtype=np.float32))


nd.softmax = _softmax


def _gemm2(a, b, transpose_a=False, transpose_b=False, alpha=1.0):
    A, B = _a(a), _a(b)
    if transpose_a:
        A = np.swapaxes(A, -1, -2)
    if transpose_b:
        B = np.swapaxes(B, -1, -2)
    return NDArray(alpha * np.matmul(A, B))


nd.linalg_gemm2 = _gemm2


def _pick(x, index, axis=-1, keepdims=False):
    a, i = _a(x), _a(index).astype(np.int64)
    out = np.take_along_axis(a, np.expand_dims(i, axis), axis=axis)
    return NDArray(out if keepdims else np.squeeze(out, axis=axis))


nd.pick = _pick


def _split(x, num_outputs, axis=1, squeeze_axis=False):
    parts = np.split(_a(x), num_outputs, axis=axis)
    if squeeze_axis:
        parts = [np.squeeze(p, axis=axis) for p in parts]
    return [NDArray(p) for p in parts]


nd.split = _split


def _sequence_mask(data, use_sequence_length=False, sequence_length=None, axis=0, value=0.0):
    a = _a(data).copy()
    if not use_sequence_length:
        return NDArray(a)
    L = _a(sequence_length)
    assert axis == 1 and a.ndim == 2       # the only form the reference uses (model.py:246-247)
    pos = np.arange(a.shape[1])[None, :]
    a[pos >= L[:, None]] = value
    return NDArray(a)


nd.SequenceMask = _sequence_mask

_RNG = {"normal": None}


def _random_normal(loc=0, scale=1.0, shape=None, ctx=None):
    eps = _RNG["normal"]
    assert eps is not None and tuple(eps.shape) == tuple(shape), "inject eps via set_normal()"
    return NDArray(loc + scale * eps)


nd.random_normal = _random_normal


def set_normal(eps):
    _RNG["normal"] = None if eps is None else np.asarray(eps, dtype=np.float32)


# F-style free functions that loss.py calls as F.*
for _n in ("where", "ones_like", "zeros_like", "log", "pick", "mean", "squeeze", "sigmoid", "broadcast_mul",
           "expand_dims", "broadcast_like", "sum", "split", "broadcast_add", "repeat", "softmax"):
    pass  # all already attributes of ``nd``; F is ``nd`` itself


# ------------------------------------------------------------------ gluon
class _Param:
    def __init__(self, name, shape):
        self.name, self.shape = name, tuple(shape)
        self.value = None

    def data(self):
        assert self.value is not None, "parameter %s not set" % self.name
        return self.value


class Block:
    def __init__(self, *args, **kwargs):
        object.__setattr__(self, "_children", [])
        object.__setattr__(self, "_params", {})

    def __setattr__(self, k, v):
        if isinstance(v, Block) and hasattr(self, "_children"):
            self._children.append((k, v))
        object.__setattr__(self, k, v)

    @contextlib.contextmanager
    def name_scope(self):
        yield

    def _reg(self, name, shape):
        p = _Param(name, shape)
        self._params[name] = p
        return p

    def collect_params(self, prefix=""):
        out = {}
        for n, p in self._params.items():
            out[prefix + n] = p
        for cn, c in self._children:
            out.update(c.collect_params(prefix + cn + "."))
        return out

    def hybridize(self, *a, **k):
        pass

    def initialize(self, *a, **k):
        pass

    def forward(self, *args):
        return self.hybrid_forward(nd, *args)

    def __call__(self, *args):
        return self.forward(*args)


HybridBlock = Block


class Dense(Block):
    def __init__(self, units, in_units=0, activation=None, flatten=True, use_bias=True):
        super().__init__()
        self.units, self.act, self.flatten = units, activation, flatten
        self.weight = self._reg("weight", (units, in_units))
        self.bias = self._reg("bias", (units,))

    def forward(self, x):
        a = _a(x)
        if self.flatten and a.ndim > 2:
            a = a.reshape(a.shape[0], -1)
        y = np.matmul(a, self.weight.data().a.T) + self.bias.data().a
        if self.act == "relu":
            y = np.maximum(y, 0)
        elif self.act is not None:
            raise NotImplementedError(self.act)
        return NDArray(y)


class Embedding(Block):
    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.weight = self._reg("weight", (input_dim, output_dim))

    def forward(self, x):
        return NDArray(self.weight.data().a[_a(x).astype(np.int64)])


class LayerNorm(Block):
    def __init__(self, axis=-1, epsilon=1e-5, in_channels=0):
        super().__init__()
        self.eps = epsilon
        self.gamma = self._reg("gamma", (in_channels,))
        self.beta = self._reg("beta", (in_channels,))

    def forward(self, x):
        a = _a(x)
        mean = a.mean(axis=-1, keepdims=True, dtype=np.float32)
        var = ((a - mean) ** 2).mean(axis=-1, keepdims=True, dtype=np.float32)
        return NDArray((a - mean) / np.sqrt(var + np.float32(self.eps)) * self.gamma.data().a + self.beta.data().a)


class Dropout(Block):
    def __init__(self, rate):
        super().__init__()
        self.rate = rate

    def forward(self, x):
        assert self.rate == 0.0, "golden vectors are generated with dropout 0"
        return x


class LSTM(Block):
    """gluon.rnn.LSTM (fused RNN op): gates i,f,g,o; states (h0,c0) each [layers,B,H]."""

    def __init__(self, hidden_size, num_layers=1, layout="TNC", dropout=0, bidirectional=False, input_size=0):
        super().__init__()
        assert not bidirectional and layout == "NTC"
        self.H, self.n = hidden_size, num_layers
        for l in range(num_layers):
            ins = input_size if l == 0 else hidden_size
            self._reg("l%d_i2h_weight" % l, (4 * hidden_size, ins))
            self._reg("l%d_h2h_weight" % l, (4 * hidden_size, hidden_size))
            self._reg("l%d_i2h_bias" % l, (4 * hidden_size,))
            self._reg("l%d_h2h_bias" % l, (4 * hidden_size,))

    def forward(self, x, states):
        a = _a(x)
        h0, c0 = _a(states[0]), _a(states[1])
        H = self.H
        sig = lambda v: 1.0 / (1.0 + np.exp(-v))
        hn, cn = [], []
        for l in range(self.n):
            Wi = self._params["l%d_i2h_weight" % l].data().a
            Wh = self._params["l%d_h2h_weight" % l].data().a
            bi = self._params["l%d_i2h_bias" % l].data().a
            bh = self._params["l%d_h2h_bias" % l].data().a
            h, c = h0[l], c0[l]
            outs = []
            for t in range(a.shape[1]):
                g = a[:, t] @ Wi.T + bi + h @ Wh.T + bh
                i, f, gg, o = sig(g[:, :H]), sig(g[:, H:2 * H]), np.tanh(g[:, 2 * H:3 * H]), sig(g[:, 3 * H:])
                c = f * c + i * gg
                h = o * np.tanh(c)
                outs.append(h)
            a = np.stack(outs, axis=1).astype(np.float32)
            hn.append(h)
            cn.append(c)
        return NDArray(a), [NDArray(np.stack(hn)), NDArray(np.stack(cn))]


class Loss(Block):
    def __init__(self, weight=None, batch_axis=0, **kw):
        super().__init__()
        self._weight, self._batch_axis = weight, batch_axis


class SoftmaxCrossEntropyLoss(Loss):
    def __init__(self, axis=-1, sparse_label=True, from_logits=False, weight=None, batch_axis=0, **kw):
        super().__init__(weight, batch_axis)
        self._axis, self._sparse_label, self._from_logits = axis, sparse_label, from_logits


class _NDArrayIter:
    """mx.io.NDArrayIter without shuffling (only what data.py touches at construction)."""

    def __init__(self, data, label=None, batch_size=1, shuffle=False):
        self.data, self.label, self.batch_size, self.shuffle = data, label, batch_size, shuffle


def install():
    """Register the stand-in ``mxnet`` and ``midi`` modules in sys.modules."""
    from oracle import smf

    mx = types.ModuleType("mxnet")
    mx.nd = nd
    mx.ndarray = nd
    nd.random = types.SimpleNamespace(multinomial=None, normal=_random_normal)
    nd.linalg = types.SimpleNamespace(maketrian=None, gemm2=_gemm2)
    mx.Context = object
    mx.cpu = lambda *a: "cpu"
    mx.gpu = lambda *a: "gpu"
    gluon = types.ModuleType("mxnet.gluon")
    gluon.Block = gluon.HybridBlock = Block
    gnn = types.ModuleType("mxnet.gluon.nn")
    gnn.Dense, gnn.Embedding, gnn.LayerNorm, gnn.Dropout = Dense, Embedding, LayerNorm, Dropout
    grnn = types.ModuleType("mxnet.gluon.rnn")
    grnn.LSTM = LSTM
    gloss = types.ModuleType("mxnet.gluon.loss")
    gloss.Loss, gloss.SoftmaxCrossEntropyLoss = Loss, SoftmaxCrossEntropyLoss
    gluon.nn, gluon.rnn, gluon.loss = gnn, grnn, gloss
    mx.gluon = gluon
    mio = types.ModuleType("mxnet.io")
    mio.NDArrayIter = _NDArrayIter
    mio.DataBatch = object
    mx.io = mio
    mx.init = types.SimpleNamespace(Xavier=lambda *a, **k: None)
    mx.metric = types.SimpleNamespace()
    for name, mod in (("mxnet", mx), ("mxnet.nd", nd), ("mxnet.gluon", gluon), ("mxnet.gluon.nn", gnn),
                      ("mxnet.gluon.rnn", grnn), ("mxnet.gluon.loss", gloss), ("mxnet.io", mio)):
        sys.modules[name] = mod

    midi = types.ModuleType("midi")
    midi.read_midifile = smf.read_midifile
    midi.NoteOnEvent, midi.NoteOffEvent, midi.SetTempoEvent = smf.NoteOnEvent, smf.NoteOffEvent, smf.SetTempoEvent
    midi.EndOfTrackEvent, midi.Pattern, midi.Track = smf.EndOfTrackEvent, smf.Pattern, smf.Track
    sys.modules["midi"] = midi
    return mx
