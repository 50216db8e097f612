"""Transformer "programs": how a (possibly mixed) transformer is lowered onto the elementwise kernels.

Every transformer module describes itself as a list of :class:`Part` -- one kernel launch each -- giving
the kernel kind, its constants, which columns of ``x`` / ``y`` it maps and where its parameters live in
the conditioner output.  An elementary transformer is one part; ``MixedTransformer`` concatenates the
parts of its children (reference nn/transformers/mixed.py:168-186).  The same program is executed with
the reference's parameter-major layout (public ``transformer.forward(x, parameters)`` API) or with the
degree-sorted feature-major layout produced by the packed conditioner inside ``MAF``.
"""

import torch

from . import _ops
from ._ops import ParamLayout


class Part:
    def __init__(self, kind, spec, n_features, n_params, offset=0, cols=None, param_major=True):
        self.kind, self.spec = kind, spec
        self.n_features, self.n_params = int(n_features), int(n_params)
        self.offset = int(offset)
        self.cols = None if cols is None else torch.as_tensor(cols).long().cpu()
        self.param_major = param_major          # column = offset + p * F + f ; else offset + f * P + p
        self._dev = {}

    def ref_columns(self):
        """(n_features, n_params) LongTensor: column of parameter p of feature f in the reference layout."""
        f = torch.arange(self.n_features)[:, None]
        p = torch.arange(self.n_params)[None, :]
        if self.param_major:
            return self.offset + p * self.n_features + f
        return self.offset + f * self.n_params + p

    def ref_layout(self):
        if self.param_major:
            return ParamLayout(self.offset, self.n_features, 1)
        return ParamLayout(self.offset, 1, self.n_params)

    def x_columns(self):
        return torch.arange(self.n_features) if self.cols is None else self.cols

    def moved(self, offset, outer_cols):
        """Same part with its parameter block shifted by ``offset`` and its columns mapped through ``outer_cols``."""
        cols = self.cols
        if outer_cols is not None:
            outer_cols = torch.as_tensor(outer_cols).long().cpu()
            cols = outer_cols[self.x_columns()]
        return Part(self.kind, self.spec, self.n_features, self.n_params, self.offset + offset, cols, self.param_major)

    def cols_on(self, device):
        if self.cols is None:
            return None
        key = str(device)
        if key not in self._dev:
            self._dev[key] = self.cols.to(device=device, dtype=torch.int32)
        return self._dev[key]


def n_parameters(parts):
    return sum(p.n_features * p.n_params for p in parts)


class ProgramFunction(torch.autograd.Function):
    """(y, logdet) = program(x, par) with the hand-written VJP kernels as backward."""

    @staticmethod
    def forward(ctx, x, par, parts, layouts, inverse, passthrough):
        dev = x.device
        y = x.clone() if passthrough else torch.empty_like(x)
        ld = torch.empty(x.shape[0], dtype=x.dtype, device=dev)
        any_ld_grad = False
        for i, (part, layout) in enumerate(zip(parts, layouts)):
            _ops.transformer_apply(part.kind, part.spec, x, par, layout, part.n_features, inverse=inverse,
                                   cols=part.cols_on(dev), y=y, logdet=ld, accumulate=i > 0)
            any_ld_grad = any_ld_grad or part.kind not in ('sos', 'shift')
        ctx.save_for_backward(x, par)
        ctx.meta = (parts, layouts, inverse, passthrough)
        if not any_ld_grad:
            ctx.mark_non_differentiable(ld)         # the reference's SOS log-det carries no gradient (sos.py:233);
                                                    # the volume-preserving shift returns constant zeros
        return y, ld

    @staticmethod
    def backward(ctx, grad_y, grad_ld):
        x, par = ctx.saved_tensors
        parts, layouts, inverse, passthrough = ctx.meta
        if inverse:
            raise NotImplementedError('tfep_b200: gradients through the inverse direction are not implemented')
        dev = x.device
        grad_y = torch.zeros_like(x) if grad_y is None else grad_y.contiguous()
        # every kernel writes all the entries it owns: zero-fill only what no part covers (without passthrough the parts
        # map every column of x, as in forward where y starts uninitialised)
        gx = grad_y.clone() if passthrough else torch.empty_like(x)
        covered = n_parameters(parts) == par.shape[1]
        gpar = (torch.empty if covered else torch.zeros)(
            (par.shape[0], par.stride(0) if par.shape[0] > 1 else par.shape[1]), dtype=par.dtype,
            device=par.device)[:, :par.shape[1]]    # same leading dimension as par (possibly a padded view)
        for part, layout in zip(parts, layouts):
            _ops.transformer_vjp(part.kind, part.spec, x, par, layout, part.n_features, grad_y, grad_ld,
                                 cols=part.cols_on(dev), grad_x=gx, grad_par=gpar)
        return gx, gpar, None, None, None, None


def run(parts, x, par, layouts=None, inverse=False, passthrough=False):
    """Execute a program.  ``layouts`` defaults to the reference layout of every part."""
    if layouts is None:
        layouts = [p.ref_layout() for p in parts]
    return ProgramFunction.apply(x, par, parts, layouts, inverse, passthrough)


def run_group(parts, layouts, group_ids, src, dst, par, logdet, inverse, first):
    """Apply the program to a subset of features (one degree group of the inverse sweep), no autograd.

    group_ids[i] is an int32 device tensor of local feature ids of part i (or None if the part has no
    feature in the group).  Reads ``src`` and writes ``dst`` through the parts' column maps; the log-det is
    accumulated into ``logdet`` (overwritten by the first launch when ``first``)."""
    dev = src.device
    launched = False
    for part, layout, ids in zip(parts, layouts, group_ids):
        if ids is None or len(ids) == 0:
            continue
        _ops.transformer_apply(part.kind, part.spec, src, par, layout, len(ids), inverse=inverse,
                               cols=part.cols_on(dev), feat_ids=ids, y=dst, logdet=logdet,
                               accumulate=not (first and not launched))
        launched = True
    return launched
