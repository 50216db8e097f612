"""Chain of normalizing flows (reference tfep/nn/flows/sequential.py:24-68)."""

import torch


class SequentialFlow(torch.nn.Sequential):
    """Runs the wrapped flows in order (reverse order for ``inverse``) and sums their log-det Jacobians."""

    def n_parameters(self):
        """The total number of parameters that can be optimized."""
        return sum(flow.n_parameters() for flow in self)

    def forward(self, x):
        return self._pass(x, inverse=False)

    def inverse(self, y):
        return self._pass(y, inverse=True)

    def _pass(self, x, inverse):
        if len(self) > 1 and self._fused_chain_ok(x, inverse):
            return self._inverse_fused_chain(x) if inverse else self._forward_fused_chain(x)
        cumulative_log_det_J = None
        for flow in (reversed(self) if inverse else self):
            x, log_det_J = flow.inverse(x) if inverse else flow(x)
            cumulative_log_det_J = log_det_J if cumulative_log_det_J is None else cumulative_log_det_J + log_det_J
        if cumulative_log_det_J is None:
            cumulative_log_det_J = torch.zeros(x.size(0), dtype=x.dtype, device=x.device)
        return x, cumulative_log_det_J

    # -- tensor-core fast path: the whole chain of precision='bf16' MAF layers in ONE kernel launch --------
    def _fused_chain_ok(self, x, inverse=False):
        from .maf import MAF
        if not all(isinstance(f, MAF) and f.precision == 'bf16' for f in self):
            return False
        if not all(f._use_fused(x) for f in self):
            return False
        return not inverse or all(f._fused_plan().inverse_eligibility(f) is None for f in self)

    def _forward_fused_chain(self, x):
        from ... import _fused
        pairs = []
        for f in self:
            f._check_fused_inference(x)
            pairs.append((f._fused_plan(), f))
        if not _fused.chain_compatible([pl for pl, _ in pairs]) or sum(len(pl.ops_host) for pl, _ in pairs) > _fused.MAX_OPS:
            y, ld = x, None
            for pl, f in pairs:
                y, l = pl.forward(f, y)
                ld = l if ld is None else ld + l
            return y, ld
        return _fused.run_chain(pairs, x)

    def _inverse_fused_chain(self, y):
        from ... import _fused
        pairs = []
        for f in reversed(self):
            f._check_fused_inference(y)
            pairs.append((f._fused_plan(), f))
        if not _fused.chain_compatible([pl for pl, _ in pairs]):
            x, ld = y, None
            for pl, f in pairs:
                x, l = _fused.run_inverse_chain([(pl, f)], x)
                ld = l if ld is None else ld + l
            return x, ld
        return _fused.run_inverse_chain(pairs, y)
