"""Command-line flags and YAML-serialisable config objects — same flag table as the reference's
VarAutoEncoder/config.py:19-75 (names, types, defaults; unknown flags are ignored via parse_known_args) plus
additive flags for this implementation, and a ``Config`` base class with the reference's save / load / copy /
freeze surface (config.py:90-222)."""
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
b200_arg.add_argument('--precision', choices=["fp32", "tf32", "bf16"], default="tf32",
                      help="GEMM path: exact fp32 FFMA, tcgen05 TF32 tensor cores (fp32 storage / accumulation), or bf16 = "
                           "the TF32 path with the Transformer layers' GEMM operands stored as bfloat16 (fp32 accumulation, "
                           "fp32 master weights / LayerNorm / softmax / losses / Adam)")
b200_arg.add_argument('--cuda-graph', type=str2bool, default=True,
                      help="replay the train step from a CUDA graph per batch shape (one launch instead of ~70)")
b200_arg.add_argument('--seed', type=int, default=0)
b200_arg.add_argument('--max-steps', type=int, default=-1, help="stop fit() after this many batches (-1: no limit)")
b200_arg.add_argument('--log-dir', type=str, default='/tmp/out')


def get_config(argv=None):
    config, unparsed = parser.parse_known_args(argv)
    return config


class _TaggedMeta(yaml.YAMLObjectMetaclass):
    """Every subclass gets the YAML tag ``!<ClassName>`` so configs round-trip as typed objects."""

    def __init__(cls, name, bases, kwds):
        cls.yaml_tag = "!" + name
        cls.yaml_loader = yaml.UnsafeLoader
        super().__init__(name, bases, dict(kwds, yaml_tag="!" + name))


class Config(yaml.YAMLObject, metaclass=_TaggedMeta):
    """Freezable, YAML-(de)serialisable configuration object (interface of config.py:90-222)."""

    def __init__(self):
        self.__add_frozen()

    def __setattr__(self, key, value):
        if getattr(self, '_frozen', False):
            raise AttributeError("Cannot set '%s' in frozen config" % key)
        if value is self:
            raise AttributeError("Cannot set self as attribute")
        object.__setattr__(self, key, value)

    def __setstate__(self, state):
        self.__dict__.update(state)
        # constructor defaults for arguments an older saved config does not carry
        for pname, param in inspect.signature(self.__init__).parameters.items():
            if param.default is not param.empty and not hasattr(self, pname):
                object.__setattr__(self, pname, param.default)

    def freeze(self):
        if getattr(self, '_frozen', False):
            return
        object.__setattr__(self, "_frozen", True)
        for k, v in self.__dict__.items():
            if isinstance(v, Config) and k != "self":
                v.freeze()

    def __repr__(self):
        return "Config[%s]" % ", ".join("%s=%s" % (str(k), str(v)) for k, v in sorted(self.__dict__.items()))

    def __eq__(self, other):
        if type(other) is not type(self):
            return False
        return all(k in other.__dict__ and other.__dict__[k] == v for k, v in self.__dict__.items() if k != "self")

    __hash__ = None

    def __del_frozen(self):
        if '_frozen' in self.__dict__:
            object.__delattr__(self, '_frozen')
        for val in self.__dict__.values():
            if isinstance(val, Config):
                val.__del_frozen()

    def __add_frozen(self):
        object.__setattr__(self, "_frozen", False)
        for val in self.__dict__.values():
            if isinstance(val, Config):
                val.__add_frozen()

    def save(self, fname: str):
        obj = copy.deepcopy(self)
        obj.__del_frozen()
        with open(fname, 'w') as out:
            yaml.dump(obj, out, default_flow_style=False)

    @staticmethod
    def load(fname: str) -> 'Config':
        with open(fname) as inp:
            obj = yaml.load(inp, Loader=yaml.UnsafeLoader)
            obj.__add_frozen()
            return obj

    def copy(self, **kwargs):
        copy_obj = copy.deepcopy(self)
        for name, value in kwargs.items():
            object.__setattr__(copy_obj, name, value)
        return copy_obj

    def set_attrs(self, attrs):
        for k, v in attrs.items():
            if k == 'self' and self is v:
                continue
            if hasattr(self, k):
                print('Not automatically over writing setting %s, %s. %s is already defined for Object %s' % (k, str(v), k, self))
            else:
                setattr(self, k, v)

    def output_to_stream(self, stream):
        obj = copy.deepcopy(self)
        obj.__del_frozen()
        yaml.dump(obj, stream)
