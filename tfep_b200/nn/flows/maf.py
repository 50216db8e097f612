"""Masked autoregressive flow layer (reference tfep/nn/flows/maf.py:33-194).

Same constructor arguments, parameter / buffer names and return values as the reference ``MAF``.  The
evaluation is re-designed:

* forward: the MADE conditioner runs in degree-sorted packed order (``tfep_b200._pack.MadePlan``) with
  its output layer packed feature-major, so every GEMM skips the all-zero part of its masked weights,
  ELU is fused in the GEMM epilogue and the transformer kernels read each feature's parameters as one
  contiguous run;
* inverse: ONE degree-ordered sweep through the packed network (each hidden unit and each output row
  is computed exactly once, when everything it depends on is known) instead of the reference's
  ``n_degrees`` full conditioner + transformer passes (autoregressive.py:216-227).  The result is the
  same map; the log-det is accumulated per degree group and equals the reference's last-pass value.
"""

from collections.abc import Sequence
from typing import Optional, Union

import torch

from ... import _ops, _program
from ..._ops import ParamLayout
from ..._pack import MadePlan
from ...utils.misc import ensure_tensor_sequence
from ..conditioners.made import MADE
from ..transformers.affine import AffineTransformer
from ..transformers.transformer import Transformer
from .autoregressive import AutoregressiveFlow, _InverseFunction, _needs_grad


PRECISIONS = ('fp32', 'bf16', 'bf16x3', 'bf16x6')


class MAF(AutoregressiveFlow):
    """Masked Autoregressive Flow: an autoregressive flow with a MADE conditioner and any MAF transformer.

    Parameters
    ----------
    degrees_in : Sequence[int]
        Degree of each input; consecutive values starting from 0, or -1 for conditioning features (which
        affect every output but are not mapped).
    transformer : MAFTransformer, optional
        Default: :class:`AffineTransformer`.
    hidden_layers : int, Sequence[int] or Sequence[Sequence[int]]
        See :class:`tfep_b200.nn.conditioners.MADE`.
    embedding : torch.nn.Module, optional
        Applied to the conditioner input; must implement ``get_degrees_out(degrees_in)``.
    weight_norm : bool
    initialize_identity : bool
    precision : str
        The only argument the reference does not have (keyword only).  ``'fp32'`` (default): exact FFMA kernels, parity
        <= 1e-5 with the reference.  ``'bf16'``: tensor cores with bf16 operands and fp32 accumulation (the fused
        one-launch kernels where they cover the layer, else the general tcgen05 GEMM), stated tolerance in DESIGN.md.
        ``'bf16x3'`` / ``'bf16x6'``: tensor cores with every operand split into two / three bf16 terms (3 / 6 products
        per reduction step, fp32 accumulation): the conditioner at fp32-class accuracy on the tensor cores.
        Also an attribute: ``maf.precision = ...`` switches an existing module.  With ``'bf16'`` and an affine / SOS (two
        polynomials) / Moebius (3-vectors) / neural-spline (8 bins) transformer over all features, the transformer and its VJP run in the epilogue of
        the output-layer product, forward and backward (``maf.fuse_transformer = False`` keeps the separate kernels).
    """

    def __init__(
            self,
            degrees_in: Sequence[int],
            transformer: Optional[torch.nn.Module] = None,
            hidden_layers: Union[int, Sequence[int], Sequence[Sequence[int]]] = 2,
            embedding: Optional[torch.nn.Module] = None,
            weight_norm: bool = True,
            initialize_identity: bool = True,
            *,
            precision: str = 'fp32',
    ):
        if precision not in PRECISIONS:
            raise ValueError(f'precision must be one of {PRECISIONS}')
        if transformer is None:
            transformer = AffineTransformer()
        degrees_in = ensure_tensor_sequence(degrees_in)
        min_degree_in = degrees_in.min().tolist()
        max_degree_in = degrees_in.max().tolist()
        if ((set(degrees_in.tolist()) != set(range(min_degree_in, max_degree_in + 1))) or
                (min_degree_in not in {-1, 0})):
            raise ValueError('degrees_in must assume consecutive values starting '
                             'from 0 (or -1 for conditioning input features).')
        degrees_in_embedded = degrees_in if embedding is None else embedding.get_degrees_out(degrees_in)
        transformer_indices = [(degrees_in == degree).nonzero().flatten() for degree in range(max_degree_in + 1)]
        degrees_out = transformer.get_degrees_out(degrees_in[degrees_in != -1])

        super().__init__(
            n_features_in=len(degrees_in),
            transformer_indices=transformer_indices,
            conditioner=_EmbeddedMADE(embedding=embedding, degrees_in=degrees_in_embedded, degrees_out=degrees_out,
                                      hidden_layers=hidden_layers, weight_norm=weight_norm),
            transformer=transformer,
            initialize_identity=initialize_identity,
        )
        self._embedding = embedding
        self._degrees_in_host = degrees_in.long().cpu().clone()
        self._packing = None
        self._fused = None
        self._sweep = None
        self._blocked = None
        self._tctx = None
        #: False: keep the transformer kernels separate from the tensor-core conditioner (precision='bf16')
        self.fuse_transformer = True
        #: degrees per block of the blocked inverse sweep of wide conditioners (tfep_b200/_blocked.py)
        self.inverse_block_degrees = 64
        #: see the class docstring; one of PRECISIONS
        self.precision = precision

    def invalidate_packed(self):
        """Forget every packed copy of the conditioner weights; required after writing parameters through ``.data``
        (see :meth:`tfep_b200.nn.conditioners.MADE.invalidate_packed`)."""
        self._conditioner.invalidate_packed()

    def n_parameters(self) -> int:
        """The total number of (unmasked) parameters."""
        return self._conditioner.n_parameters()

    # -- packed fast path -------------------------------------------------------------------------
    def _pack(self):
        """Build (once) the packed plan: output rows grouped per feature, features sorted by degree."""
        if self._packing is not None:
            return self._packing
        if not isinstance(self._transformer, Transformer):
            self._packing = False
            return False
        parts = self._native_parts()
        deg = self._degrees_in_host
        feats = []                                   # (degree, x column, part index, local feature)
        for pi, p in enumerate(parts):
            for f, c in enumerate(p.x_columns().tolist()):
                feats.append((int(deg[c]), c, pi, f))
        feats.sort()
        out_order, bases = [], [torch.zeros(p.n_features, dtype=torch.int32) for p in parts]
        ref_cols = [p.ref_columns() for p in parts]
        for _, _, pi, f in feats:
            bases[pi][f] = len(out_order)
            out_order.extend(ref_cols[pi][f].tolist())
        made = self._conditioner
        assert sorted(out_order) == list(range(made.dimension_out)), 'transformer program does not cover the conditioner output'
        plan = MadePlan(made._degree_chain, out_order=torch.tensor(out_order))
        # row range of the packed output layer owned by each degree group, and its features per part
        groups = []
        pos = 0
        for g, cols in enumerate(self._groups_host):
            gset = set(cols.tolist())
            sel = [(pi, f) for d, c, pi, f in feats if c in gset]
            width = sum(parts[pi].n_params for pi, _ in sel)
            groups.append(dict(degree=g, rows=(pos, pos + width),
                               ids=[[f for pi2, f in sel if pi2 == pi] for pi in range(len(parts))]))
            pos += width
        self._packing = dict(parts=parts, plan=plan, bases=bases, groups=groups, dev={})
        return self._packing

    def _packed_tables(self, device):
        pk = self._pack()
        key = str(device)
        if key not in pk['dev']:
            layouts = [ParamLayout(0, 1, 0, base=b.to(device)) for b in pk['bases']]
            gids = [[torch.tensor(ids, dtype=torch.int32, device=device) if ids else None for ids in g['ids']]
                    for g in pk['groups']]
            pk['dev'][key] = (layouts, gids)
        return pk['dev'][key]

    def forward(self, x: torch.Tensor):
        """Returns ``(y, log_det_J)`` with shapes ``(batch, n_features)`` and ``(batch,)``."""
        pk = self._pack()
        if self.precision not in PRECISIONS:
            raise ValueError(f'precision must be one of {PRECISIONS}')
        if self.precision == 'bf16' and self._use_fused(x):
            return self._forward_fused(x)
        if pk is False or self._n_conditioner_indices > 0:
            return super().forward(x)
        if self.precision == 'bf16' and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[0] > 0:
            # tensor-core conditioner with the transformer applied by the epilogue of its output layer, forward and
            # backward (affine / SOS / Moebius; tfep_b200/_txfused.py)
            txp = self._tc_tx_plan() if self.fuse_transformer else None
            if txp is not None:
                return txp.forward(self, pk, x.contiguous())
        layouts, _ = self._packed_tables(x.device)
        xc = x if self._embedding is None else self._embedding(x)
        # precision='bf16' outside the fused kernel (other transformers, training): the MADE conditioner runs on the
        # general tensor-core GEMM, forward and backward; the transformer kernels stay exact
        par = self._conditioner.run_plan(xc.contiguous(), pk['plan'], precision=self.precision)
        return _program.run(pk['parts'], x.contiguous(), par, layouts, passthrough=self.has_fixed_indices)

    def _tc_tx_plan(self):
        """Plan of the fused transformer epilogue of the general tensor-core path, or None (``_tctx_why`` says why)."""
        if self._tctx is None:
            from ... import _txfused
            pk = self._pack()
            why = _txfused.eligibility(self, pk)
            self._tctx = _txfused.TcTxPlan(self, pk) if why is None else False
            self._tctx_why = why
        return self._tctx or None

    def _use_fused(self, x):
        """The one-launch fused kernel serves inference of the splines it covers; everything else that asks for
        precision='bf16' goes through the general tensor-core GEMM."""
        from ... import _fused
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return False
        return self._fused_plan() is not None

    def _fused_plan(self):
        """The plan of the fused kernels, or None if they do not cover this layer (``_fused_why`` says why: transformer
        kind, widths beyond the tensor-memory plan, ...)."""
        from ... import _fused, _lib
        if self._fused is None:
            why = _fused.eligibility(self)
            if why is None:
                try:
                    self._fused = _fused.FusedSplinePlan(self)
                except _lib.TfepB200Error as e:
                    why = str(e)
            if why is not None:
                self._fused, self._fused_why = False, why
        return self._fused or None

    def _forward_fused(self, x):
        """Tensor-core path: the whole layer in one kernel launch (no autograd, no silent fallback)."""
        self._check_fused_inference(x)
        return self._fused_plan().forward(self, x)

    def _check_fused_inference(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError("tfep_b200: precision='bf16' is an inference path; use torch.no_grad() "
                                      "or precision='fp32' for training")

    def inverse(self, y: torch.Tensor):
        """Returns ``(x, log_det_J)``: degree-ordered sweep (see the module docstring)."""
        if _needs_grad(self, y):
            return _InverseFunction.apply(self, y, *[p for p in self.parameters() if p.requires_grad])
        pk = self._pack()
        from ... import _sweep
        if pk is False or self._n_conditioner_indices > 0 or \
                (self._embedding is not None and _sweep.eligibility(self, pk) is not None):
            return super().inverse(y)
        if any(p.kind == 'sos' for p in pk['parts']):
            raise NotImplementedError('Inversion of SOS polynomial transformer has not been implemented yet.')
        if self.precision == 'bf16':
            from ... import _fused
            plan = self._fused_plan()
            if plan is not None and plan.inverse_eligibility(self) is None:
                return _fused.run_inverse_chain([(plan, self)], y)
            # (flows the tensor-core sweep does not cover are inverted by the exact sweep below)
        from ... import _blocked, _sweep
        if _sweep.eligibility(self, pk) is not None:
            return self._inverse_host_sweep(y)
        layouts, _ = self._packed_tables(y.device)
        if _blocked.needs_blocking(self, pk) and self._embedding is None and not self.has_fixed_indices:
            # wide conditioner (the activations of a sample tile do not fit an SM): degree blocks, GEMM panels between
            # them, the persistent sweep inside them (tfep_b200/_blocked.py)
            if self._blocked is None:
                self._blocked = _blocked.BlockedSweepPlan(self, pk, block_degrees=self.inverse_block_degrees)
            panels = self.precision if (self.precision != 'fp32' and y.dtype == torch.float32) else 'fp32'
            with torch.no_grad():
                return self._blocked.inverse(self, y, layouts, panel_precision=panels)
        if self._sweep is None:
            self._sweep = _sweep.SweepPlan(self, pk)
        with torch.no_grad():
            return self._sweep.inverse(self, y, layouts)

    def _inverse_host_sweep(self, y: torch.Tensor):
        """The same degree-ordered sweep driven from the host (four launches per degree); kept as the
        cross-check of the persistent kernel (tests/test_gpu_maf.py)."""
        pk = self._pack()
        y = y.contiguous()
        plan, parts, groups = pk['plan'], pk['parts'], pk['groups']
        layouts, gids = self._packed_tables(y.device)
        made = self._conditioner
        with torch.no_grad():
            pw, pb = made.packed_weights(plan)
            L = plan.n_layers
            B = y.shape[0]
            x = torch.zeros_like(y)
            if self.has_fixed_indices:
                x[:, self._fixed_indices] = y[:, self._fixed_indices]
            hidden = [None] + [torch.empty(B, pw[l].shape[0], dtype=y.dtype, device=y.device) for l in range(L - 1)]
            max_width = max(g['rows'][1] - g['rows'][0] for g in groups)
            par = torch.empty(B, max_width, dtype=y.dtype, device=y.device)
            log_det_J = torch.zeros(B, dtype=y.dtype, device=y.device)

            def update_hidden(degree):
                for l in range(1, L):
                    a, b = plan.degree_rows(l, degree)
                    if a == b:
                        continue
                    if l == 1:
                        src = x
                    else:
                        src = hidden[l - 1][:, :plan.degree_prefix(l - 1, degree, strict=False)]
                    _ops.linear_forward(src, pw[l - 1][a:b, :src.shape[1]], pb[l - 1][a:b], _ops.ACT_ELU,
                                        out=hidden[l][:, a:b])

            if int(self._degrees_in_host.min()) == -1:
                update_hidden(-1)
            first = True
            for g, grp in enumerate(groups):
                r0, r1 = grp['rows']
                src = x if L == 1 else hidden[L - 1][:, :plan.degree_prefix(L - 1, grp['degree'], strict=True)]
                out = par[:, :r1 - r0]
                _ops.linear_forward(src, pw[L - 1][r0:r1, :src.shape[1]], pb[L - 1][r0:r1], _ops.ACT_NONE, out=out)
                shifted = [ParamLayout(-r0, 1, 0, base=lay.base) for lay in layouts]
                launched = _program.run_group(parts, shifted, gids[g], y, x, par, log_det_J, inverse=True, first=first)
                first = first and not launched
                if g != len(groups) - 1:
                    update_hidden(grp['degree'])
        return x, log_det_J


class _EmbeddedMADE(MADE):
    """A MADE conditioner whose input optionally goes through an embedding layer first (reference maf.py:184-194)."""

    def __init__(self, embedding, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.embedding = embedding

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.embedding is not None:
            x = self.embedding(x)
        return super().forward(x)
