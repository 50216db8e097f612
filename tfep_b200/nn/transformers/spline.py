"""Rational-quadratic neural spline transformer, incl. the circular variant.

Same constructor, buffers (``state_dict`` keys) and parameter layout as the reference's
``NeuralSplineTransformer`` (tfep/nn/transformers/spline.py:29-417); the arithmetic (softmax / softplus
normalisation of the raw parameters, circular shift and wrap, knot construction, bin search, rational
quadratic map and log-det; spline.py:319-650) runs in tfepb_spline / tfepb_spline_backward.
"""

from typing import Optional

import torch

from ... import _program
from .transformer import MAFTransformer


class NeuralSplineTransformer(MAFTransformer):
    r"""Neural spline transformer (Durkan et al. 2019) with K bins on ``[x0, xf] -> [y0, yf]``.

    Outside the domain the map is linear with the boundary slopes.  ``circular=True`` implements circular
    splines (Rezende et al. 2020): the boundary slopes are tied and an extra shift parameter is learned,
    :math:`y = \mathrm{spline}((x - x_0 + \phi) \bmod (x_f - x_0) + x_0)`.  With ``identity_boundary_slopes`` the
    boundary slopes are fixed to 1; with ``learn_lower_bound`` / ``learn_upper_bound`` the domain is scaled
    (and shifted) by extra parameters.  See the reference docstring for the parameter layout
    (spline.py:184-241): ``parameters`` reshapes to ``(batch, n_parameters_per_feature, n_features)`` =
    widths (K), heights (K), slopes (K+1, K or K-1), [domain shift], [domain scale | circular shift].
    """

    def __init__(
            self,
            x0: torch.Tensor,
            xf: torch.Tensor,
            n_bins: int,
            y0: Optional[torch.Tensor] = None,
            yf: Optional[torch.Tensor] = None,
            circular: bool = False,
            identity_boundary_slopes: bool = False,
            learn_lower_bound: bool = False,
            learn_upper_bound: bool = False,
            min_bin_size: float = 1e-4,
            min_slope: float = 1e-4,
    ):
        super().__init__()
        if y0 is None:
            y0 = x0.detach()
        if yf is None:
            yf = xf.detach()
        if circular and (learn_lower_bound or learn_upper_bound):
            raise ValueError('Cannot instantiate a circular spline with learnable limits.')
        if circular and not (torch.allclose(x0[circular], y0[circular]) and torch.allclose(xf[circular], yf[circular])):
            raise ValueError('x0==y0 and xf==yf must hold for all periodic degrees of freedom.')
        if min_bin_size <= 0.:
            raise ValueError('The minimum bin size should be positive.')
        if (min_slope <= 0.) or (min_slope >= 1.):
            raise ValueError('The minimum slope should be between 0 and 1.')

        self.register_buffer('x0', x0)
        self.register_buffer('xf', xf)
        self.register_buffer('n_bins', torch.as_tensor(n_bins))
        self.register_buffer('_y0', y0)
        self.register_buffer('_yf', yf)
        self.register_buffer('_circular', torch.as_tensor(circular))
        self.register_buffer('_identity_boundary_slopes', torch.as_tensor(identity_boundary_slopes))
        self.register_buffer('_learn_lower_bound', torch.as_tensor(learn_lower_bound))
        self.register_buffer('_learn_upper_bound', torch.as_tensor(learn_upper_bound))
        self.register_buffer('_min_bin_size', torch.as_tensor(min_bin_size))
        self.register_buffer('_min_slope', torch.as_tensor(min_slope))
        self._sync_host_config()

    # Scalar configuration is mirrored on the host so that no forward pass has to read it back from the
    # device; the buffers stay the source of truth for ``state_dict``.
    def _sync_host_config(self):
        self.n_bins_int = int(self.n_bins)
        self.circular = bool(self._circular)
        self.identity_slopes = bool(self._identity_boundary_slopes)
        self.learn_lower = bool(self._learn_lower_bound)
        self.learn_upper = bool(self._learn_upper_bound)
        self.min_bin_size = float(self._min_bin_size)
        self.min_slope = float(self._min_slope)
        self._domain_cache = {}

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._sync_host_config()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._domain_cache = {}
        return out

    def domain_tensors(self, dtype, device):
        """(x0, xf, y0, yf) as contiguous tensors of the kernel's dtype on ``device``."""
        key = (dtype, str(device), self.x0.data_ptr(), self.x0._version, self.xf._version)
        hit = self._domain_cache.get('dom')
        if hit is None or hit[0] != key:
            hit = (key, tuple(t.to(device=device, dtype=dtype).contiguous() for t in (self.x0, self.xf, self._y0, self._yf)))
            self._domain_cache['dom'] = hit
        return hit[1]

    @property
    def n_parameters_per_feature(self) -> int:
        n = 3 * self.n_bins_int + 1 + int(self.learn_lower) + int(self.learn_upper)
        if self.identity_slopes:
            n -= 1 if self.circular else 2
        return n

    def get_identity_parameters(self, n_features: int) -> torch.Tensor:
        if not (torch.allclose(self.x0, self._y0) and torch.allclose(self.xf, self._yf)):
            raise ValueError('The identity neural spline transformer can be implemented only if x0=y0 and xf=yf.')
        return torch.zeros(size=(self.n_parameters_per_feature, n_features)).to(self.x0).reshape(-1)

    def get_degrees_out(self, degrees_in: torch.Tensor) -> torch.Tensor:
        return degrees_in.tile((self.n_parameters_per_feature,))

    def _parts(self, n_features):
        return [_program.Part('spline', self, n_features, self.n_parameters_per_feature)]

    def bin_indices(self, x, parameters, inverse=False):
        """Bin index of every input (0 = left tail, 1..K = spline bins, K+1 = right tail), as computed by the
        kernel; debugging / parity aid (reference semantics: spline.py:622-625)."""
        from ... import _ops
        bins = torch.empty(x.shape, dtype=torch.int32, device=x.device)
        part = self._parts(x.shape[1])[0]
        _ops.transformer_apply('spline', self, x, parameters, part.ref_layout(), part.n_features, inverse=inverse,
                               bins=bins)
        return bins


def neural_spline_transformer(x, x0, y0, widths, heights, slopes):
    """Functional rational-quadratic spline on already-normalised widths / heights / slopes.

    Reference: spline.py:424-501.  Implemented by inverting the parameter normalisation (log of the
    normalised bins, inverse softplus of the slopes) and calling the fused kernel with minimum sizes of
    (almost) zero, so it shares its arithmetic; intended for tests and small inputs.
    """
    batch, K, F = widths.shape
    tiny = torch.finfo(widths.dtype).tiny
    W = widths.sum(dim=1)
    H = heights.sum(dim=1)
    x0b = x0.expand(batch, F) if x0.dim() < 2 else x0
    y0b = y0.expand(batch, F) if y0.dim() < 2 else y0
    if not (torch.allclose(W, W[:1].expand_as(W)) and torch.allclose(H, H[:1].expand_as(H))
            and torch.allclose(x0b, x0b[:1].expand_as(x0b)) and torch.allclose(y0b, y0b[:1].expand_as(y0b))):
        raise NotImplementedError('tfep_b200.neural_spline_transformer supports a per-feature (not per-sample) domain')
    min_slope = 1e-12
    t = NeuralSplineTransformer(x0=x0b[0].clone(), xf=(x0b[0] + W[0]), n_bins=K, y0=y0b[0].clone(), yf=(y0b[0] + H[0]),
                                min_bin_size=float(tiny), min_slope=min_slope)
    offset = torch.log(torch.expm1(torch.tensor(1. - min_slope, dtype=widths.dtype)))
    s = slopes - min_slope
    raw = torch.where(s > 20, s, torch.log(torch.expm1(s))) - offset.to(slopes.device)
    par = torch.cat([torch.log(widths), torch.log(heights), raw], dim=1).reshape(batch, -1)
    return t(x, par)
