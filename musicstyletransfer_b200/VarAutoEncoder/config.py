"""Command-line flags and YAML-serialisable config objects.

The flag table is the reference's (VarAutoEncoder/config.py:19-75: names, types, defaults; unknown flags are
ignored through parse_known_args) plus additive flags of this implementation.  ``Config`` offers the
reference's public surface (config.py:90-222: save / load / copy / freeze / set_attrs / output_to_stream and
the ``!ClassName`` YAML tags of the on-disk ``config`` file) on a different mechanism: subclasses register
themselves in a class registry, (de)serialisation goes through a SafeLoader / SafeDumper pair that only knows
the registered classes (a model folder's ``config`` cannot construct arbitrary Python objects), and the frozen
state lives in a per-instance slot that is never part of the serialised mapping."""
import argparse
import copy
import inspect

import yaml


def str2bool(v):
    return v.lower() in ('true', '1')


arg_lists = []
parser = argparse.ArgumentParser()


def add_argument_group(name):
    arg = parser.add_argument_group(name)
    arg_lists.append(arg)
    return arg


# Network
net_arg = add_argument_group('Network')
net_arg.add_argument('--e-n-layers', type=int, default=1)
net_arg.add_argument('--e-rnn-hidden-dim', type=int, default=128)
net_arg.add_argument('--e-emb-hidden-dim', type=int, default=64)
net_arg.add_argument('--e-dropout', type=float, default=0.0)
net_arg.add_argument('--e-num-heads', type=int, default=8)
net_arg.add_argument('--latent-dim', type=int, default=64)
net_arg.add_argument('--d-n-layers', type=int, default=1)
net_arg.add_argument('--d-rnn-hidden-dim', type=int, default=128)
net_arg.add_argument('--d-dropout', type=float, default=0.0)

# Data
data_arg = add_argument_group('Data')
data_arg.add_argument('--batch-size', type=int, default=1)
data_arg.add_argument('--max-seq-len', type=int, default=64)
data_arg.add_argument('--slices-per-quarter-note', type=float, default=4)
data_arg.add_argument('--data', type=str, default='data')
data_arg.add_argument('--validation-data', type=str, default=None)
data_arg.add_argument('--minimum-pattern-length', type=int, default=16)
data_arg.add_argument('--pattern-identifier', type=str, choices=['recurring', ''], default='')

# Training / test parameters
train_arg = add_argument_group('Training')
train_arg.add_argument('--epochs', type=int, default=5000)
train_arg.add_argument('--learning-rate', type=float, default=3e-4)
train_arg.add_argument('--optimizer', type=str, default='adam')
train_arg.add_argument('--optimizer-params', type=str, default='')
train_arg.add_argument('--validation-split', type=float, default=0.1)
train_arg.add_argument('--kl-loss', type=float, default=1.0)
train_arg.add_argument('--label-smoothing', type=float, default=0.0)
train_arg.add_argument('--negative-label-downscaling', action='store_true')
train_arg.add_argument('--beam-size', type=int, default=5)
train_arg.add_argument('--sampling-type', choices=["beam-search", "sampling"], default="sampling")

# Misc
misc_arg = add_argument_group('Misc')
misc_arg.add_argument('--load-checkpoint', type=int, default=1)
misc_arg.add_argument('--checkpoint-frequency', type=int, default=5000)
misc_arg.add_argument('--sampling-frequency', type=int, default=1000)
misc_arg.add_argument('--num-checkpoints-not-improved', type=int, default=10)
misc_arg.add_argument('--out-samples', '-o', type=str, default=None)
misc_arg.add_argument('--model-output', '-m', type=str, default='models')
misc_arg.add_argument('--checkpoint', '-c', type=int, default=-1)
misc_arg.add_argument('--gpu', action='store_true')
misc_arg.add_argument('--toy', action='store_true')
misc_arg.add_argument('--visualize-samples', action='store_true')
misc_arg.add_argument('--verbose', action='store_true')

# Additive flags of the B200 implementation
b200_arg = add_argument_group('B200')
b200_arg.add_argument('--decoder-type', choices=["lstm", "transformer"], default="lstm",
                      help="lstm = the decoder scripts/train-vae.sh's --d-* flags describe (model.py:131-203); "
                           "transformer = the decoder class Model instantiates at HEAD (model.py:206-272)")
b200_arg.add_argument('--precision', choices=["fp32", "fp32x3", "tf32x3f", "bf16x3f", "bf16p3f", "tf32", "bf16"], default="bf16p3f",
                      help="precision mode (engine.PRECISIONS): bf16p3f (default) = fp32-class forward (encoder GEMMs from bf16 "
                           "hi + lo operand planes, ~2^-17 per product, compensated attention scores) with a TF32 backward; "
                           "tf32x3f = the same with the operand split inside the GEMM (3xTF32); tf32 = every tensor-core product single-pass TF32; fp32x3 = "
                           "strict fp32 on the tensor cores (3xTF32 everywhere, exact attention / LSTM); fp32 = exact FFMA; bf16 = "
                           "tf32 with the Transformer layers' GEMM operands stored as bfloat16")
b200_arg.add_argument('--cuda-graph', type=str2bool, default=True,
                      help="replay the train step from a CUDA graph per batch shape (one launch instead of ~70)")
b200_arg.add_argument('--featurisation', choices=["events", "roll"], default="events",
                      help="events = token ids (the reference's HEAD); roll = K1's piano-roll windows of --max-seq-len slices with the "
                           "sigmoid-BCE reconstruction loss (loss.py:27-81; consumes --label-smoothing and "
                           "--negative-label-downscaling), LSTM decoder")
b200_arg.add_argument('--device-dataset', type=str2bool, default=True,
                      help="build the dataset rows on the GPU and gather every batch there (A2 on the device); false = the "
                           "host NumPy rows of the reference's MelodyDataset")
b200_arg.add_argument('--seed', type=int, default=0)
b200_arg.add_argument('--max-steps', type=int, default=-1, help="stop fit() after this many batches (-1: no limit)")
b200_arg.add_argument('--log-dir', type=str, default='/tmp/out')


def get_config(argv=None):
    config, unparsed = parser.parse_known_args(argv)
    return config


class _ConfigLoader(yaml.SafeLoader):
    """SafeLoader that additionally constructs the registered Config subclasses from their ``!ClassName`` tags."""


class _ConfigDumper(yaml.SafeDumper):
    """SafeDumper that writes registered Config subclasses as ``!ClassName`` mappings."""


_STATE_KEY = "_frozen"          # the only per-instance attribute that is not configuration


def _public_items(obj):
    return [(k, v) for k, v in sorted(vars(obj).items()) if k != _STATE_KEY]


def _represent_config(dumper, obj):
    return dumper.represent_mapping("!" + type(obj).__name__, _public_items(obj))


class Config:
    """Base class of the model / transformer / LSTM configuration objects."""

    registry = {}

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        Config.registry[cls.__name__] = cls
        _ConfigDumper.add_representer(cls, _represent_config)
        _ConfigLoader.add_constructor("!" + cls.__name__, lambda loader, node, _cls=cls: _cls._from_yaml(loader, node))

    def __init__(self):
        object.__setattr__(self, _STATE_KEY, False)

    # ---------------------------------------------------------------- construction from a saved mapping
    @classmethod
    def _from_yaml(cls, loader, node):
        obj = cls.__new__(cls)
        object.__setattr__(obj, _STATE_KEY, False)
        yield obj                                   # two-step construction: nested configs / anchors resolve first
        obj._restore(loader.construct_mapping(node, deep=True))

    def _restore(self, fields):
        """Fills the instance from a saved mapping; constructor arguments that a file written by an older version does
        not carry take their declared defaults, so old model folders keep loading."""
        for k, v in fields.items():
            object.__setattr__(self, k, v)
        for name, param in inspect.signature(type(self).__init__).parameters.items():
            if name != "self" and param.default is not inspect.Parameter.empty and name not in fields:
                object.__setattr__(self, name, param.default)

    def __getstate__(self):
        return dict(_public_items(self))

    def __setstate__(self, state):                  # pickle / deepcopy
        object.__setattr__(self, _STATE_KEY, False)
        self._restore(state)

    # ---------------------------------------------------------------- attribute protocol
    def __setattr__(self, key, value):
        if vars(self).get(_STATE_KEY, False):
            raise AttributeError("Cannot set '%s' in frozen config" % key)
        if value is self:
            raise AttributeError("Cannot set self as attribute")
        object.__setattr__(self, key, value)

    def _children(self):
        return [v for _, v in _public_items(self) if isinstance(v, Config)]

    def freeze(self):
        """Disallows any further modification of this object and of the configs nested in it."""
        object.__setattr__(self, _STATE_KEY, True)
        for child in self._children():
            child.freeze()

    @property
    def frozen(self):
        return bool(vars(self).get(_STATE_KEY, False))

    def __repr__(self):
        return "Config[%s]" % ", ".join("%s=%s" % (k, v) for k, v in _public_items(self))

    def __eq__(self, other):
        return type(other) is type(self) and dict(_public_items(self)) == dict(_public_items(other))

    __hash__ = None

    # ---------------------------------------------------------------- (de)serialisation
    def output_to_stream(self, stream):
        yaml.dump(self, stream, Dumper=_ConfigDumper, default_flow_style=False)

    def save(self, fname: str):
        """Writes the configuration (never its frozen state) to `fname`."""
        with open(fname, "w") as out:
            self.output_to_stream(out)

    @staticmethod
    def load(fname: str) -> "Config":
        """Reads a configuration written by save(); the result is not frozen.  Only registered Config subclasses and
        plain YAML scalars / lists / mappings are constructed."""
        with open(fname) as inp:
            obj = yaml.load(inp, Loader=_ConfigLoader)
        if not isinstance(obj, Config):
            raise ValueError("%s does not hold a Config object" % fname)
        return obj

    def copy(self, **kwargs):
        """Unfrozen deep copy, optionally with some attributes replaced: ``cfg.copy(num_layers=3)``."""
        dup = copy.deepcopy(self)
        for name, value in kwargs.items():
            object.__setattr__(dup, name, value)
        return dup

    def set_attrs(self, attrs):
        """Constructor helper, ``self.set_attrs(locals())``: adds every entry that is not already an attribute."""
        for k, v in attrs.items():
            if k == "self" and v is self:
                continue
            if hasattr(self, k):
                print("Not automatically over writing setting %s, %s. %s is already defined for Object %s" % (k, v, k, self))
            else:
                setattr(self, k, v)
