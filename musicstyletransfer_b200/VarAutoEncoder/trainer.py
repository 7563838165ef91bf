"""Training driver with the reference's interface (VarAutoEncoder/trainer.py): ``Trainer(config, context, model,
sampler).fit(dataset, model_folder, epochs, validation_dataset)`` and ``._step(batch, is_train)``.

The step is one engine call sequence: forward (+ fused CE/KL), hand-scheduled backward into the flat gradient
arena, optional NCCL all-reduce of that arena (data parallelism, one process per GPU) and the fused Adam pass.
File layout of checkpoints is the reference's: ``params.<n>``, ``train_state.pkl`` (+ ``opt_state.<n>`` with the
Adam moments and the step count of the random streams, which the reference does not save).

Provenance: ``OptimizerConfig`` / ``TrainConfig`` / ``TrainingState`` and the control flow of ``fit`` / ``_checkpoint`` /
``_periodic_log`` (what is printed, when checkpoints are written, the early-stopping rule) deliberately follow the
reference's trainer.py:14-65,122-153,202-255 — they are host control flow that a drop-in must reproduce; ``_step``, the
optimiser, the metrics, data parallelism, CUDA-graph replay and the checkpoint contents are new."""
import os
from time import time

import numpy as np
import torch

from . import utils
from .metrics import DeviceMetrics
from .utils import to_device_i32


class OptimizerConfig:
    def __init__(self, optimizer: str, optimizer_params: str, learning_rate: float):
        self.optimizer = optimizer
        self.optimizer_params = optimizer_params
        self.learning_rate = learning_rate

    def params_to_dict(self):
        """'key1:value1,key2:value2' -> dict of floats (trainer.py:23-35)."""
        out = {}
        for key_val in self.optimizer_params.strip().split(','):
            key_val = key_val.split(':')
            if len(key_val) != 2:
                continue
            out[str(key_val[0])] = float(key_val[1])
        return out


class TrainConfig:
    def __init__(self, batch_size: int, sampling_frequency: int, checkpoint_frequency: int,
                 num_checkpoints_not_improved: int, optimizer: OptimizerConfig, kl_loss: float, label_smoothing: float,
                 negative_label_downscaling: bool, verbose: bool):
        self.batch_size = batch_size
        self.sampling_frequency = sampling_frequency
        self.checkpoint_frequency = checkpoint_frequency
        self.num_checkpoints_not_improved = num_checkpoints_not_improved
        self.optimizer = optimizer
        self.kl_loss_weight = kl_loss
        self.label_smoothing = label_smoothing
        self.negative_label_downscaling = negative_label_downscaling
        self.verbose = verbose


class TrainingState:
    def __init__(self):
        self.n_checkpoints = 0
        self.n_batches = 0
        self.num_checkpoints_not_improved = 0
        self.best_resconstruction_loss = np.inf


class _LazyLoss:
    """Per-sample total loss of a step (what the reference's _step returns, trainer.py:172,179), formed only when the
    caller actually looks at it: the training loop itself does not, and forming it eagerly would put torch elementwise
    kernels on every step."""

    def __init__(self, ce, kl, kl_weight):
        self.ce, self.kl, self.kl_weight = ce, kl, kl_weight

    def value(self):
        return self.ce + self.kl_weight * self.kl

    def __getattr__(self, name):                # tensor protocol by delegation (mean(), cpu(), shape, ...)
        return getattr(self.value(), name)

    def __array__(self, dtype=None):
        a = self.value().detach().cpu().numpy()
        return a if dtype is None else a.astype(dtype)


class Trainer:
    def __init__(self, config: TrainConfig, context, model, sampler=None, log_dir='/tmp/out', max_steps=-1, cuda_graph=True):
        self.config = config
        # replay the train step from a CUDA graph per batch shape (engine.train_step_graphed): at the reference's own
        # batch size (32, scripts/train-vae.sh) the ~70 launches of an eagerly issued step are host-bound
        self.cuda_graph = cuda_graph
        self.context = context
        self.model = model
        self.engine = model.engine
        self.sampler = sampler
        self.max_steps = max_steps
        self.world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        self.rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        self._initialize_model()
        self._initialize_optimizers()
        self._initialize_metrics()
        self.summary_writer = None
        if self.rank == 0:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.summary_writer = SummaryWriter(log_dir=log_dir, flush_secs=5)
            except Exception:                      # tensorboard is optional
                self.summary_writer = None
        self.train_state = TrainingState()

    def _initialize_optimizers(self):
        assert self.config.optimizer.optimizer == 'adam', "the fused optimizer kernel implements MXNet's Adam"
        p = {'learning_rate': self.config.optimizer.learning_rate}
        p.update(self.config.optimizer.params_to_dict())
        self.opt = dict(lr=p['learning_rate'], beta1=p.get('beta1', 0.9), beta2=p.get('beta2', 0.999),
                        eps=p.get('epsilon', 1e-8), wd=p.get('wd', 0.0), clip_gradient=p.get('clip_gradient', None))

    def _initialize_model(self):
        self.model.initialize()                    # mx.init.Xavier() (trainer.py:103-105)
        if self.world > 1:                         # every rank starts from rank 0's parameters
            torch.distributed.broadcast(self.engine.arena.w, src=0)

    def _initialize_metrics(self):
        self.metrics = DeviceMetrics(self.engine)

    def fit(self, dataset, model_folder: str, epochs: int, validation_dataset=None):
        start_time = time()
        self.train_state = TrainingState()
        self._load_latest_checkpoint(model_folder)
        for epoch in range(epochs):
            for batch in dataset:
                self._step(batch)
                self.train_state.n_batches += 1
                if self.train_state.n_batches % 50 == 0:
                    self._periodic_log(epoch, start_time)
                if self.train_state.n_batches % self.config.checkpoint_frequency == 0:
                    self._checkpoint(model_folder, validation_dataset)
                    if self.train_state.num_checkpoints_not_improved == self.config.num_checkpoints_not_improved:
                        print("Maximum checkpoints not improved reached. Stopping training.")
                        return
                if self.sampler is not None and self.config.sampling_frequency > 0 and \
                        self.train_state.n_batches % self.config.sampling_frequency == 0 and self.rank == 0:
                    self.sampler.update_parameters(self.model)
                    self.sampler.process_batch(batch, os.path.join(model_folder, 'samples/step-{}'.format(
                        self.train_state.n_batches)), dataset.num_classes())
                if 0 < self.max_steps <= self.train_state.n_batches:
                    return

    def _shard(self, t):
        """Data parallelism: rank r takes rows r::world of the global batch."""
        return t[self.rank::self.world] if self.world > 1 else t

    def _step_roll(self, batch, is_train=True):
        """Piano-roll batch (data = [roll uint8 [B, S, 128], classes]): BCE reconstruction + KL (--featurisation roll)."""
        dev = self.engine.device
        roll = self._shard(batch.data[0]).to(dev).contiguous()
        classes = to_device_i32(self._shard(batch.data[1]), dev)
        global_batch = batch.data[0].shape[0]
        o = self.opt
        kw = dict(kl_weight=self.config.kl_loss_weight, label_smoothing=self.config.label_smoothing,
                  downweight=bool(self.config.negative_label_downscaling))
        if is_train and self.cuda_graph and self.world == 1 and o['beta1'] == 0.9 and o['beta2'] == 0.999 and \
                o['eps'] == 1e-8 and o['wd'] == 0.0:
            out = self.engine.train_step_roll_graphed(roll, classes, global_batch=global_batch, lr=o['lr'],
                                                      clip_gradient=o['clip_gradient'], **kw)
        else:
            out = self.engine.forward_roll(roll, classes, train=True, label_smoothing=kw["label_smoothing"],
                                           downweight=kw["downweight"], want_grad=is_train)
            if is_train:
                self.engine.backward_roll(kl_weight=self.config.kl_loss_weight)
                if self.world > 1:
                    torch.distributed.all_reduce(self.engine.arena.g)
                self.engine.adam_step(global_batch, **self.opt)
        self.metrics.update(out["bce"], out["kl"], self.config.kl_loss_weight)
        return _LazyLoss(out["bce"], out["kl"], self.config.kl_loss_weight)

    def _step(self, batch, is_train=True):
        if torch.is_tensor(batch.data[0]) and batch.data[0].dtype == torch.uint8 and batch.data[0].dim() == 3:
            return self._step_roll(batch, is_train)
        dev = self.engine.device
        tokens, seq_lens, classes = [to_device_i32(self._shard(x), dev) for x in batch.data]
        labels = to_device_i32(self._shard(batch.label[0]), dev)
        global_batch = batch.data[0].shape[0]
        if self.config.verbose:
            print("Step {}".format(self.train_state.n_batches))
            print("tokens:  {}, {}".format(tuple(tokens.shape), tokens))
            print("classes: {}, {}".format(tuple(classes.shape), classes))
            print("labels:  {}, {}".format(tuple(labels.shape), labels))
        # the reference keeps autograd.record() (train mode, dropout on) for validation too (trainer.py:166-168)
        o = self.opt
        graph_ok = (is_train and self.cuda_graph and not self.config.verbose and o['beta1'] == 0.9 and o['beta2'] == 0.999
                    and o['eps'] == 1e-8 and o['wd'] == 0.0)
        if graph_ok:
            # forward + backward (+ all-reduce) + Adam as one graph replay per (batch, length) shape; the optimiser
            # hyper-parameters baked into the graph are the reference's defaults checked above
            ar = (lambda g: torch.distributed.all_reduce(g)) if self.world > 1 else None
            if ar is not None and not hasattr(self, "_ar"):
                self._ar = ar                      # one callable object: it is part of the graph cache key
            out = self.engine.train_step_graphed(tokens, seq_lens, classes, labels, kl_weight=self.config.kl_loss_weight,
                                                 global_batch=global_batch, lr=o['lr'], clip_gradient=o['clip_gradient'],
                                                 allreduce=getattr(self, "_ar", None))
        else:
            out = self.engine.forward(tokens, seq_lens, classes, labels, train=True)
            if is_train:
                self.engine.backward(kl_weight=self.config.kl_loss_weight)
                if self.world > 1:
                    torch.distributed.all_reduce(self.engine.arena.g)
                self.engine.adam_step(global_batch, **self.opt)
        self.metrics.update(out["ce"], out["kl"], self.config.kl_loss_weight)
        return _LazyLoss(out["ce"], out["kl"], self.config.kl_loss_weight)

    def _load_latest_checkpoint(self, model_folder):
        print("Looking into folder {} for a valid training.".format(model_folder))
        try:
            latest_checkpoint = utils.get_latest_checkpoint_index(model_folder)
        except Exception:
            print("No checkpoint was found. Starting training from scratch")
            return
        print("Checkpoint {} found. Resuming training.".format(latest_checkpoint))
        utils.load_model_parameters(self.model, os.path.join(model_folder, "params.{}".format(latest_checkpoint)), self.context)
        self.train_state = utils.load_object(os.path.join(model_folder, "train_state.pkl"))
        opt_path = os.path.join(model_folder, "opt_state.{}".format(latest_checkpoint))
        if os.path.exists(opt_path):
            st = torch.load(opt_path, map_location="cpu")
            a = self.engine.arena
            a.m.copy_(st["m"]); a.v.copy_(st["v"]); a.adam_state.copy_(st["state"])
            # continue the dropout / eps seed sequence where the saved run stopped (older files: the Adam step count)
            self.engine.set_step_count(int(st.get("step_count", int(st["state"][0]))))

    def _checkpoint(self, model_folder, validation_dataset):
        self.train_state.n_checkpoints += 1
        print("\nCheckpoint {} reached.".format(self.train_state.n_checkpoints))
        if self.rank == 0:
            utils.create_directory_if_not_present(model_folder)
            utils.save_model(self.model, os.path.join(model_folder, 'params.{}'.format(self.train_state.n_checkpoints)))
            utils.save_object(self.train_state, os.path.join(model_folder, "train_state.pkl"))
            a = self.engine.arena
            torch.save({"m": a.m.cpu(), "v": a.v.cpu(), "state": a.adam_state.cpu(), "step_count": self.engine.step_count},
                       os.path.join(model_folder, "opt_state.{}".format(self.train_state.n_checkpoints)))
        self._reset_metrics()
        if validation_dataset is None:
            return
        for batch in validation_dataset:
            self._step(batch, is_train=False)
        # every rank validated its own shard of each batch: sum the counters over ranks so that the improved / stop
        # decision below (and best_resconstruction_loss) is the same everywhere
        self.metrics.all_reduce()
        reconstruction_loss = dict(self.metrics.get_name_value())["total_loss"]
        if reconstruction_loss < self.train_state.best_resconstruction_loss:
            print("Loss improved from {} to {}.".format(self.train_state.best_resconstruction_loss, reconstruction_loss))
            self.train_state.best_resconstruction_loss = reconstruction_loss
        else:
            self.train_state.num_checkpoints_not_improved += 1
            print("Loss did not improve. {} out {} unsucessful checkpoints".format(
                self.train_state.num_checkpoints_not_improved, self.config.num_checkpoints_not_improved))
            print("Best loss thus far: {}".format(self.train_state.best_resconstruction_loss))
        print("Checkpoint [{}]  {}\n".format(self.train_state.n_checkpoints,
                                             self._metric_to_string_output(self.train_state.n_batches)))
        self._reset_metrics()

    def _reset_metrics(self):
        self.metrics.reset()

    def _metric_to_string_output(self, n_batches):
        out = ''
        for metric_name, val in self.metrics.get_name_value():
            if self.summary_writer is not None:
                self.summary_writer.add_scalar(metric_name, val, global_step=n_batches)
            out += '{}={:.3f} '.format(metric_name, val)
        self.metrics.reset()
        return out

    def _periodic_log(self, epoch, start_time):
        if self.rank != 0:
            self.metrics.reset()
            return
        print("Epoch [{}] Batch [{}] updates/sec: {:.2f} {}".format(
            epoch, self.train_state.n_batches, self.train_state.n_batches / (time() - start_time),
            self._metric_to_string_output(self.train_state.n_batches)))
