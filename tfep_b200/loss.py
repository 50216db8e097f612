"""KL-divergence training loss between Boltzmann distributions (reference tfep/loss.py:26-140).

One streaming reduction kernel (tfepb_kl_loss) replaces the reference's chain of elementwise / softmax / sum
launches over ``(batch,)`` vectors; its backward is one elementwise kernel (tfepb_kl_loss_backward).
"""

from typing import Optional

import torch

from . import _lib
from ._lib import check, dtype_code, ptr, stream_ptr


class _KLDivLossFunction(torch.autograd.Function):
    """tfepb_kl_loss / tfepb_kl_loss_backward as an autograd node over the four optional vectors."""

    @staticmethod
    def forward(ctx, ignore_nan, u_b, ld, u_a, lw):
        _lib.require_cuda(u_b, ld, u_a, lw)
        tensors = [None if t is None else t.detach().to(u_b.dtype).contiguous() for t in (u_b, ld, u_a, lw)]
        n = tensors[0].numel()
        for t in tensors[1:]:
            if t is not None and t.numel() != n:
                raise ValueError('BoltzmannKLDivLoss takes vectors of equal length')
        lib = _lib.load()
        ws = torch.empty(lib.tfepb_kl_loss_workspace_bytes() // 8, dtype=torch.float64, device=u_b.device)
        stats = torch.empty(5, dtype=torch.float64, device=u_b.device)
        with torch.cuda.device(u_b.device):
            check(lib.tfepb_kl_loss(dtype_code(tensors[0]), *(ptr(t) for t in tensors), n, int(ignore_nan), ptr(ws),
                                    ptr(stats), stream_ptr(u_b)))
        ctx.ignore_nan = bool(ignore_nan)
        ctx.save_for_backward(stats, *[t for t in tensors if t is not None])
        ctx.present = [t is not None for t in tensors]
        return stats[0].to(u_b.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        stats, *saved = ctx.saved_tensors
        it = iter(saved)
        tensors = [next(it) if present else None for present in ctx.present]
        needs = ctx.needs_input_grad[1:]
        grads = [torch.empty_like(t) if (t is not None and need) else None for t, need in zip(tensors, needs)]
        go = grad_out.detach().to(tensors[0].dtype).reshape(1).contiguous()
        with torch.cuda.device(go.device):
            check(_lib.load().tfepb_kl_loss_backward(dtype_code(tensors[0]), *(ptr(t) for t in tensors), tensors[0].numel(),
                                                     int(ctx.ignore_nan), ptr(stats), ptr(go), *(ptr(g) for g in grads),
                                                     stream_ptr(go)))
        return (None, *grads)


class BoltzmannKLDivLoss(torch.nn.Module):
    """KL divergence between the mapped distribution A and the target Boltzmann distribution B, up to a constant.

    ``forward(target_potentials, log_det_J=None, log_weights=None, ref_potentials=None)`` with ``(batch,)`` CUDA
    vectors in units of kT returns the 0-dim loss: the mean reduced work ``u_B - log|det J| - u_A`` or, given
    ``log_weights``, its ``softmax(log_weights)``-weighted sum; ``ignore_nan`` drops NaN terms the way
    ``torch.nanmean`` / ``torch.nansum`` do (constructor argument of the reference, tfep/loss.py:60-74).

    Gradients equal the reference's autograd results, with one deliberate difference: with ``ignore_nan`` and a NaN
    term, the reference's gradient with respect to ``log_weights`` is NaN everywhere (``0 * NaN`` inside the softmax
    backward); here it is the gradient of the sum over the kept terms.
    """

    def __init__(self, ignore_nan: bool = False):
        super().__init__()
        self.ignore_nan = ignore_nan

    def forward(self, target_potentials: torch.Tensor, log_det_J: Optional[torch.Tensor] = None,
                log_weights: Optional[torch.Tensor] = None, ref_potentials: Optional[torch.Tensor] = None):
        return _KLDivLossFunction.apply(self.ignore_nan, target_potentials, log_det_J, ref_potentials, log_weights)
