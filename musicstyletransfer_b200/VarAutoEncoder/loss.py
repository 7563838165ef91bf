"""Loss classes with the reference's names and call signatures (VarAutoEncoder/loss.py), backed by the K3
kernels.  Each is callable on CUDA tensors and differentiable w.r.t. its first argument(s) through a
torch.autograd.Function whose backward is again a kernel launch.  The training step itself uses the fused
logits path of the engine (probabilities are never materialised there)."""
import torch

from .. import ops


def _f32(x, dev=None):
    t = torch.as_tensor(x)
    return t.to(device=dev if dev is not None else t.device, dtype=torch.float32).contiguous()


class _KL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, stds):
        B, Z = means.shape
        lat = torch.cat([means, stds], dim=1).contiguous()
        zeros = torch.zeros_like(means)
        z = torch.empty_like(means)
        kl = torch.empty(B, dtype=torch.float32, device=means.device)
        ops.reparam_kl_fwd(lat, zeros, z, kl, B, Z)
        ctx.save_for_backward(lat, zeros)
        return kl

    @staticmethod
    def backward(ctx, g):
        lat, zeros = ctx.saved_tensors
        B, Z2 = lat.shape
        dlat = torch.empty_like(lat)
        ops.reparam_kl_bwd(lat, zeros, None, g.contiguous(), 1.0, dlat, B, Z2 // 2)
        return dlat[:, :Z2 // 2], dlat[:, Z2 // 2:]


class VariationalKLLoss:
    """loss.py:4-12: kl_b = sum_z 0.5 (s^2 + m^2 - 1 - log s^2), no batch mean."""

    def __call__(self, z_means, z_vars):
        return _KL.apply(_f32(z_means), _f32(z_vars))


class SoftmaxCrossEntropy:
    """loss.py:15-23: on PROBABILITIES; ce_b = mean over all T columns of -log p[label] * [label != 0]."""

    def __init__(self, axis=-1, batch_axis=0, **kwargs):
        self._axis, self._batch_axis = axis, batch_axis

    def __call__(self, pred, label, sample_weight=None):
        pred = _f32(pred)
        B, T, V = pred.shape
        lab = torch.as_tensor(label).to(pred.device, torch.int32).contiguous()
        ce = torch.empty(B, dtype=torch.float32, device=pred.device)
        ops.ce_from_probs(pred, lab, ce, B, T, V)
        return ce


class _BCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, label_u8, from_sigmoid, smoothing, downweight):
        B = pred.shape[0]
        n = pred[0].numel()
        out = torch.empty(B, dtype=torch.float32, device=pred.device)
        ops.bce(pred, label_u8, out, None, None, B, n, from_sigmoid, smoothing, downweight)
        ctx.save_for_backward(pred, label_u8)
        ctx.cfg = (from_sigmoid, smoothing, downweight)
        return out

    @staticmethod
    def backward(ctx, g):
        pred, label_u8 = ctx.saved_tensors
        B = pred.shape[0]
        n = pred[0].numel()
        dpred = torch.empty_like(pred)
        ops.bce(pred, label_u8, None, g.contiguous(), dpred, B, n, *ctx.cfg)
        return dpred, None, None, None, None


class BinaryCrossEntropy:
    """loss.py:27-81 (sigmoid, label smoothing, eps 1e-12, per-sample negative-label down-weighting with the
    (w*bce)*bce quirk, mean over T*P).  ``label`` is a 0/1 piano roll (uint8 from the K1 rasteriser, or float)."""

    def __init__(self, from_sigmoid=False, label_smoothing=0.0, negative_label_downweighting=True):
        self._from_sigmoid = from_sigmoid
        self.label_smoothing = label_smoothing
        self.negative_label_downweighting = negative_label_downweighting

    def __call__(self, pred, label):
        pred = _f32(pred)
        lab = torch.as_tensor(label).to(pred.device)
        if lab.dtype != torch.uint8:
            lab = lab.to(torch.uint8)
        return _BCE.apply(pred, lab.contiguous(), self._from_sigmoid, self.label_smoothing,
                          self.negative_label_downweighting)
