"""Checkpoint / pickle / directory helpers with the reference's names (VarAutoEncoder/utils.py:15-71).
Parameters are stored as a torch state-dict keyed by the Gluon parameter paths."""
import os
import pickle
import pprint
import re

import torch


def get_latest_checkpoint_index(model_folder: str):
    """utils.py:15-25 with the evident intent: the reference's regex ``params.(\\d)+`` keeps only the last
    digit of the index; here the whole number is used."""
    checkpoint = -1
    for file in os.listdir(model_folder):
        match = re.fullmatch(r"params\.(\d+)", file)
        if match is not None:
            checkpoint = max(int(match.group(1)), checkpoint)
    if checkpoint == -1:
        raise ValueError("No checkpoints found in {}".format(model_folder))
    return checkpoint


def save_model(model, output_path: str):
    model.save_parameters(output_path)


def save_object(object, output_path: str):
    with open(output_path, "wb") as file:
        pickle.dump(object, file)


class _TrainStateUnpickler(pickle.Unpickler):
    """train_state.pkl sits in a model folder that may come from somewhere else: only the training-state class of this
    package (or the reference's module path for it) and plain scalars may be constructed, nothing else is importable."""

    _ALLOWED = {("musicstyletransfer_b200.VarAutoEncoder.trainer", "TrainingState"),
                ("music_style_transfer.VarAutoEncoder.trainer", "TrainingState"),
                ("VarAutoEncoder.trainer", "TrainingState")}

    def find_class(self, module, name):
        if (module, name) in self._ALLOWED:
            from . import trainer
            return trainer.TrainingState
        if module == "numpy" and name in ("float64", "float32", "dtype"):
            import numpy
            return getattr(numpy, name)
        if (module, name) in (("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar")):
            import numpy
            return numpy.core.multiarray.scalar if hasattr(numpy, "core") else numpy._core.multiarray.scalar
        raise pickle.UnpicklingError("train state files may only hold a TrainingState (found %s.%s)" % (module, name))


def load_object(path: str):
    with open(path, "rb") as file:
        return _TrainStateUnpickler(file).load()


def load_model_parameters(model, path: str, context=None):
    model.load_parameters(path, ctx=context)


def create_directory_if_not_present(directory: str):
    if not os.path.exists(directory):
        os.makedirs(directory)


def log_config(config):
    pprint.pprint("Using configuration: ")
    pprint.pprint(config)


def log_model_variables(model):
    print("Model variables: ")
    pprint.pprint({k: tuple(v.shape) for k, v in model.collect_params().items()})


def to_device_i32(x, device):
    """Batches follow the reference and carry float32 ids (data.py:161-169); kernels take int32."""
    t = torch.as_tensor(x)
    return t.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()
