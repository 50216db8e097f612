"""Generic autoregressive flow: conditioner -> transformer (reference tfep/nn/flows/autoregressive.py:29-247).

Same constructor, buffers and semantics as the reference.  When the transformer is one of this package's
transformers the conditioning features are handled inside the kernels (column maps) instead of by
gather / scatter; any other object implementing the ``Transformer`` API goes through the reference's
indexing recipe.  :class:`tfep_b200.nn.flows.MAF` overrides forward / inverse with the packed fast path.
"""

from collections.abc import Sequence
from typing import Optional

import torch

from ... import _program
from ...utils.misc import ensure_tensor_sequence
from ..transformers.transformer import Transformer


def _needs_grad(flow, y):
    return torch.is_grad_enabled() and (y.requires_grad or any(p.requires_grad for p in flow.parameters()))


class _InverseFunction(torch.autograd.Function):
    """``x, log_det_J = flow.inverse(y)`` (AutoregressiveFlow / MAF) with gradients by implicit differentiation of ``F(x; theta) = y``.

    The reference differentiates through its ``n_degrees`` conditioner passes (autoregressive.py:199-229).  Here the
    inverse itself runs without autograd (the sweep kernels) and the backward pass solves ``J^T v = gx - g_ld grad_x ld``
    with ``J = dF/dx`` by back-substitution over the degree groups, highest degree first -- ``J`` is block lower
    triangular in degree order, its diagonal blocks couple only features of one degree -- using vector-Jacobian
    products of ONE forward graph (the hand-written backward kernels).  Then ``grad_y = v`` and
    ``grad_theta = -(dF/dtheta)^T v - g_ld d ld / d theta``.  Cost: a few backward passes per degree.
    """

    @staticmethod
    def forward(ctx, maf, y, *params):
        with torch.no_grad():
            x, ld = maf.inverse(y)
        ctx.maf, ctx.n_params = maf, len(params)
        ctx.save_for_backward(x)
        return x, ld

    @staticmethod
    def backward(ctx, gx, gld):
        maf = ctx.maf
        (x,) = ctx.saved_tensors
        params = [p for p in maf.parameters() if p.requires_grad]
        assert len(params) == ctx.n_params
        with torch.enable_grad():
            xg = x.detach().requires_grad_(True)
            yf, ldf = maf.forward(xg)
            w = torch.zeros_like(x) if gx is None else gx.clone()
            ld_grad = gld is not None and ldf.requires_grad
            if ld_grad:
                w = w - torch.autograd.grad(ldf, xg, gld, retain_graph=True)[0]
            v = torch.zeros_like(w)
            groups = [g.to(x.device) for g in maf._groups_host]
            fixed = maf._fixed_indices.to(x.device) if maf.has_fixed_indices else None
            for cols in reversed(groups):
                # (J^T v) restricted to the group, from the degrees already solved
                r = w[:, cols]
                if bool(v.any()):
                    r = r - torch.autograd.grad(yf, xg, v, retain_graph=True)[0][:, cols]
                # diagonal block J_gg (per sample, G x G): one vector-Jacobian product per feature of the group
                G = len(cols)
                rows = []
                for j in range(G):
                    e = torch.zeros_like(w)
                    e[:, cols[j]] = 1.0
                    rows.append(torch.autograd.grad(yf, xg, e, retain_graph=True)[0][:, cols])   # row j of J_gg
                Jgg = torch.stack(rows, dim=1)                                   # (B, G, G): [b, j, k] = dy_j / dx_k
                if G == 1:
                    v[:, cols] = r / Jgg[:, 0]
                else:
                    v[:, cols] = torch.linalg.solve(Jgg.transpose(1, 2), r.unsqueeze(-1)).squeeze(-1)
            if fixed is not None and len(fixed) > 0:
                # conditioning features pass through (identity block) and feed every conditioner
                v[:, fixed] = w[:, fixed] - torch.autograd.grad(yf, xg, v, retain_graph=True)[0][:, fixed]
            gparams = [None] * len(params)
            if params:
                outs, cots = [yf], [-v]
                if ld_grad:
                    outs.append(ldf)
                    cots.append(-gld)
                gparams = torch.autograd.grad(outs, params, cots, allow_unused=True)
        return (None, v, *gparams)


class AutoregressiveFlow(torch.nn.Module):
    """Autoregressive flow: ``y_i = T(x_i; conditioner(x_<i))`` for the features in ``transformer_indices``;
    all other features are propagated unchanged.

    Parameters
    ----------
    n_features_in : int
    transformer_indices : Sequence[Sequence[int]]
        Feature indices passed to the transformer, grouped by their order in the autoregressive model
        (needed by the inverse), e.g. ``[[0, 2], [3], [1, 4]]``.
    conditioner, transformer : torch.nn.Module
    conditioner_indices : Sequence[int], optional
        Subset of features passed to the conditioner (default: all).
    initialize_identity : bool
        Initialise the flow to the identity map.
    """

    def __init__(
            self,
            n_features_in: int,
            transformer_indices: Sequence[Sequence[int]],
            conditioner: torch.nn.Module,
            transformer: torch.nn.Module,
            conditioner_indices: Optional[Sequence[int]] = None,
            initialize_identity: bool = True,
    ):
        super().__init__()
        transformer_indices = [ensure_tensor_sequence(x) for x in transformer_indices]
        if conditioner_indices is None:
            conditioner_indices = torch.tensor([], dtype=int)
        else:
            conditioner_indices = ensure_tensor_sequence(conditioner_indices)
        for indices in (conditioner_indices, *transformer_indices):
            if (indices is not None) and torch.any((indices < 0) | (n_features_in <= indices)):
                raise ValueError("All indices must be 0 <= i < n_features_in.")

        inverse_masks = torch.full((len(transformer_indices), n_features_in), False)
        for idx, indices in enumerate(transformer_indices):
            inverse_masks[idx, indices] = True
        self._groups_host = [g.long().cpu().clone() for g in transformer_indices]

        transformer_indices = torch.cat(transformer_indices).sort().values
        fixed_indices = torch.arange(n_features_in)
        fixed_indices = fixed_indices[~torch.isin(fixed_indices, transformer_indices)]
        n_transformer_indices = len(transformer_indices)
        self._mapped_host = transformer_indices.long().cpu().clone()
        self._n_fixed = len(fixed_indices)
        if len(fixed_indices) == 0:
            transformer_indices = torch.empty_like(fixed_indices)

        self._conditioner = conditioner
        self._transformer = transformer
        self.register_buffer('_transformer_indices', transformer_indices)
        self.register_buffer('_inverse_masks', inverse_masks)
        self.register_buffer('_fixed_indices', fixed_indices)
        self.register_buffer('_conditioner_indices', conditioner_indices)
        self._n_conditioner_indices = len(conditioner_indices)

        if initialize_identity:
            identity_parameters = self._transformer.get_identity_parameters(n_transformer_indices)
            self._conditioner.set_output(identity_parameters)

    @property
    def has_fixed_indices(self):
        """True if some of the features are not transformed by the flow."""
        return self._n_fixed > 0

    def _native_parts(self):
        """Kernel program of the transformer with the conditioning-column map folded in (cached)."""
        parts = getattr(self, '_parts_cache', None)
        if parts is None:
            parts = self._transformer._parts(len(self._mapped_host))
            if self.has_fixed_indices:
                parts = [p.moved(0, self._mapped_host) for p in parts]
            self._parts_cache = parts
        return parts

    def get_transformer_parameters(self, x: torch.Tensor) -> torch.Tensor:
        """Conditioner output ``(batch, n_parameters)`` in the reference's layout."""
        if self._n_conditioner_indices > 0:
            x = x[:, self._conditioner_indices]
        return self._conditioner(x)

    def forward(self, x: torch.Tensor):
        """Returns ``(y, log_det_J)`` with shapes ``(batch, n_features)`` and ``(batch,)``."""
        parameters = self.get_transformer_parameters(x)
        if isinstance(self._transformer, Transformer):
            return _program.run(self._native_parts(), x, parameters, passthrough=self.has_fixed_indices)
        if self.has_fixed_indices:
            y = torch.empty_like(x)
            y[:, self._fixed_indices] = x[:, self._fixed_indices]
            y[:, self._transformer_indices], log_det_J = self._transformer(x[:, self._transformer_indices], parameters)
            return y, log_det_J
        return self._transformer(x, parameters)

    def inverse(self, y: torch.Tensor):
        """Inverse pass: one conditioner evaluation per degree group (reference :179-229); the transformer is
        inverted only on the features of the current group.  Returns ``(x, log_det_J)``."""
        x = torch.zeros_like(y)
        if self.has_fixed_indices:
            x[:, self._fixed_indices] = y[:, self._fixed_indices]
        native = isinstance(self._transformer, Transformer)
        if native and _needs_grad(self, y):
            # the kernels below run without autograd: differentiate implicitly instead
            return _InverseFunction.apply(self, y, *[p for p in self.parameters() if p.requires_grad])
        if native:
            parts = self._native_parts()
            layouts = [p.ref_layout() for p in parts]
            group_ids = self._group_ids(y.device)
            log_det_J = torch.zeros(y.shape[0], dtype=y.dtype, device=y.device)
            with torch.no_grad():
                for gi in range(len(self._groups_host)):
                    parameters = self.get_transformer_parameters(x)
                    _program.run_group(parts, layouts, group_ids[gi], y, x, parameters, log_det_J, inverse=True,
                                       first=(gi == 0))
            return x, log_det_J
        y_t = y[:, self._transformer_indices] if self.has_fixed_indices else y
        masks_t = self._inverse_masks[:, self._transformer_indices] if self.has_fixed_indices else self._inverse_masks
        log_det_J = None
        for mask, mask_t in zip(self._inverse_masks, masks_t):
            parameters = self.get_transformer_parameters(x.clone())
            x_temp, log_det_J = self._transformer.inverse(y_t, parameters)
            x[:, mask] = x_temp[:, mask_t]
        return x, log_det_J

    def _group_ids(self, device):
        """For every degree group and every part: int32 device tensor of the part-local feature ids in the group."""
        cache = getattr(self, '_group_cache', None)
        if cache is None:
            cache = self._group_cache = {}
        key = str(device)
        if key not in cache:
            parts = self._native_parts()
            out = []
            for g in self._groups_host:
                gset = set(g.tolist())
                per_part = []
                for p in parts:
                    cols = p.x_columns().tolist()
                    ids = [f for f, c in enumerate(cols) if c in gset]
                    per_part.append(torch.tensor(ids, dtype=torch.int32, device=device) if ids else None)
                out.append(per_part)
            cache[key] = out
        return cache[key]
