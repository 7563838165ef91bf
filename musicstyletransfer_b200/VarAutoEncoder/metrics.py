"""Step metrics (trainer.py:107-120, metrics.py): Perplexity(ignore_label=0), masked Accuracy, TopKAccuracy(k=5)
and the running means of KL / total loss.  The reference pulls ``probs`` to the host three times per step
(metrics.py:32-33,60-61); here the CE kernel accumulates the four counters on the device and they are read
only when a value is requested (every 50 batches / at checkpoints)."""
import math

import torch

from .. import ops


class DeviceMetrics:
    names = ["ppl", "acc", "topk", "kl_loss", "total_loss"]

    def __init__(self, engine):
        self.engine = engine
        self.loss_sums = torch.zeros(3, dtype=torch.float32, device=engine.device)   # [sum kl, sum total, count]
        self.reset()

    def reset(self):
        self.engine.metrics.zero_()
        self.loss_sums.zero_()

    def update(self, ce, kl, kl_weight):
        """Adds this step's per-sample losses to the running sums: one libmsx launch, nothing leaves the device."""
        ops.loss_sums(ce, kl, kl_weight, self.loss_sums)

    def all_reduce(self):
        """Data parallel: every rank sees the sums over ALL shards, so checkpoint decisions (early stopping, best loss)
        are identical on every rank and no rank leaves fit() while the others wait in the next step's collective."""
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            torch.distributed.all_reduce(self.engine.metrics)
            torch.distributed.all_reduce(self.loss_sums)

    def get_name_value(self):
        m = self.engine.metrics.tolist()           # the only device->host read
        s = self.loss_sums.tolist()
        ntok = max(m[1], 1.0)
        n = max(s[2], 1.0)
        return [("ppl", math.exp(m[0] / ntok)), ("acc", m[2] / ntok), ("topk", m[3] / ntok),
                ("kl_loss", s[0] / n), ("total_loss", s[1] / n)]
