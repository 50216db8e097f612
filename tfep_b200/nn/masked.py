"""Masked linear layers and their NaN-free weight normalisation.

Same public names, arguments and ``state_dict`` keys as the reference's ``tfep/nn/masked.py``
(``create_autoregressive_mask`` :36-108, ``MaskedLinear`` :115-213, ``masked_linear`` :220-305,
``masked_weight_norm`` :312-330).  The arithmetic runs in the sm_100a kernels behind the C ABI
(tfepb_masked_linear_forward / _backward_input / _backward_weight); there is no CPU path.
"""

import torch
from torch.nn.parameter import Parameter

from .. import _ops


def create_autoregressive_mask(degrees_in, degrees_out, strictly_less=True, transpose=False, dtype=None):
    """Connectivity mask between two layers of a MADE network (reference nn/masked.py:36-108).

    ``mask[i][j]`` is 1 if input ``i`` feeds output ``j`` (``(out, in)`` indexing with ``transpose=True``):
    outputs see inputs of strictly smaller degree, or of smaller-or-equal degree if ``strictly_less=False``.
    """
    degrees_in = torch.as_tensor(degrees_in)
    degrees_out = torch.as_tensor(degrees_out)
    lhs, rhs = (degrees_out[:, None], degrees_in[None, :]) if transpose else (degrees_out[None, :], degrees_in[:, None])
    mask = (lhs > rhs) if strictly_less else (lhs >= rhs)
    return mask.to(torch.get_default_dtype() if dtype is None else dtype)


class MaskedLinearFunc(torch.autograd.Function):
    """``y = x (M o A)^T + b`` with the reference's hand-written backward (nn/masked.py:266-302)."""

    @staticmethod
    def forward(ctx, input, weight, bias=None, mask=None):
        if mask is not None:
            weight = weight * mask
        lead = input.shape[:-1]
        x2 = input.reshape(-1, input.shape[-1])
        ctx.save_for_backward(x2, weight, mask)
        ctx.has_bias = bias is not None
        ctx.lead = lead
        y = _ops.linear_forward(x2, weight, bias)
        return y.reshape(*lead, weight.shape[0])

    @staticmethod
    def backward(ctx, grad_output):
        x2, masked_weight, mask = ctx.saved_tensors
        g2 = grad_output.reshape(-1, grad_output.shape[-1]).contiguous()
        grad_input = grad_weight = grad_bias = None
        if ctx.needs_input_grad[0]:
            grad_input = _ops.linear_backward_input(g2, masked_weight).reshape(*ctx.lead, masked_weight.shape[1])
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            grad_weight, grad_bias = _ops.linear_backward_weight(g2, x2, need_bias=ctx.has_bias)
            if mask is not None:
                grad_weight = grad_weight * mask
        return grad_input, grad_weight, grad_bias, None


masked_linear = MaskedLinearFunc.apply


class MaskedLinear(torch.nn.Linear):
    r"""Masked linear transformation :math:`y = x \cdot (M \circ A)^T + b` (reference nn/masked.py:115-213)."""

    def __init__(self, in_features, out_features, bias=True, mask=None):
        super().__init__(in_features, out_features, bias=bias)
        self.register_buffer('mask', mask)
        if self.mask is not None:
            # masked entries start (and, with zero gradients, stay) at zero
            self.weight.data = self.weight.data * self.mask

    def n_parameters(self):
        """The total number of (unmasked) parameters."""
        n = self.weight.numel() if self.mask is None else (self.mask != 0).sum()
        if self.bias is not None:
            n = n + self.bias.numel()
        return n

    def forward(self, input):
        return masked_linear(input, self.weight, self.bias, self.mask)

    def extra_repr(self):
        return 'in_features={}, out_features={}, bias={}, mask={}'.format(
            self.in_features, self.out_features, self.bias is not None, self.mask)


def effective_weight(weight_v, weight_g, mask):
    """``M o (g v / |v|_row)`` with rows of zero norm mapped to zero instead of NaN.

    Value of the reference's ``MaskedWeightNorm.compute_weight`` + ``_ApplyMask`` + the mask multiply of
    ``MaskedLinearFunc.forward`` (nn/masked.py:369-371, 433-439, 270).  Differentiable w.r.t. ``g`` and ``v``
    with exactly the gradient masking the reference implements with hooks (nn/masked.py:400-402): masked
    entries of ``v`` and the ``g`` of fully masked rows receive zero gradient, never NaN.
    """
    norm = torch.linalg.vector_norm(weight_v, dim=1, keepdim=True)
    safe = torch.where(norm > 0, norm, torch.ones_like(norm))
    w = weight_v * (weight_g / safe)
    if mask is not None:
        w = w * mask
    return w


class MaskedWeightNorm:
    """Forward pre-hook recomputing ``module.<name>`` from ``<name>_g`` and ``<name>_v`` (nn/masked.py:351-404)."""

    def __init__(self, name, dim, mask):
        if dim not in (0, None):
            raise NotImplementedError('masked weight normalisation is implemented along dim 0 (per output row)')
        self.name = name
        self.dim = 0

    def compute_weight(self, module):
        g = getattr(module, self.name + '_g')
        v = getattr(module, self.name + '_v')
        return effective_weight(v, g, getattr(module, 'mask', None))

    @staticmethod
    def apply(module, name, dim, mask):
        for hook in module._forward_pre_hooks.values():
            if isinstance(hook, MaskedWeightNorm) and hook.name == name:
                raise RuntimeError('Cannot register two weight_norm hooks on the same parameter {}'.format(name))
        fn = MaskedWeightNorm(name, dim, mask)
        weight = getattr(module, name)
        del module._parameters[name]
        g = Parameter(torch.linalg.vector_norm(weight.data, dim=1, keepdim=True))
        v = Parameter(weight.data)
        module.register_parameter(name + '_g', g)
        module.register_parameter(name + '_v', v)
        setattr(module, name, fn.compute_weight(module))
        module.register_forward_pre_hook(fn)
        return fn

    def remove(self, module):
        weight = self.compute_weight(module).detach()
        delattr(module, self.name)
        del module._parameters[self.name + '_g']
        del module._parameters[self.name + '_v']
        module.register_parameter(self.name, Parameter(weight))

    def __call__(self, module, inputs):
        setattr(module, self.name, self.compute_weight(module))


def masked_weight_norm(module, name='weight', dim=0):
    """NaN-free weight normalisation of a (masked) linear module (reference nn/masked.py:312-330)."""
    MaskedWeightNorm.apply(module, name, dim, getattr(module, 'mask', None))
    return module


def remove_masked_weight_norm(module, name='weight'):
    """Undo :func:`masked_weight_norm` (reference nn/masked.py:333-348)."""
    for k, hook in module._forward_pre_hooks.items():
        if isinstance(hook, MaskedWeightNorm) and hook.name == name:
            hook.remove(module)
            del module._forward_pre_hooks[k]
            return module
    raise ValueError("weight_norm of '{}' not found in {}".format(name, module))
