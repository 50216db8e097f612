"""CPU restatement of the reference MAF forward / inverse / log|det J| path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity pinned: yes --
bit-for-bit against the real reference in fp32 and fp64
(oracle/check_against_reference.py) and against tests/golden/*.npz.

Every function names the reference lines it follows (paths relative to
/root/reference/tfep).  The arithmetic is deliberately issued as the same
sequence of ATen CPU operations as the reference so that fp32 results are
bit-identical to it; the code is organised as stateless functions and small
spec objects instead of the reference's nn.Module hierarchy.
"""

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Union

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------
# degrees and masks (integer, construction-time)
# ---------------------------------------------------------------------------

def _tile_to_length(motif: torch.Tensor, length: int, err: Optional[str] = None) -> torch.Tensor:
    """nn/conditioners/made.py:441-461 (_round_robin)."""
    reps, rest = divmod(length, len(motif))
    if reps == 0:
        raise ValueError(err or f'Length {length} is smaller than the array (len={len(motif)}).')
    out = motif.repeat(reps)
    if rest:
        out = torch.cat([out, motif[:rest]])
    return out


def gen_degrees(n_features: int, order: str = 'ascending', max_value: Optional[int] = None,
                conditioning_indices: Optional[Sequence[int]] = None,
                repeats: Union[int, Sequence[int]] = 1) -> torch.Tensor:
    """nn/conditioners/made.py:32-145 (generate_degrees)."""
    n_free = n_features - (0 if conditioning_indices is None else len(conditioning_indices))
    if max_value is None:
        if isinstance(repeats, int):
            max_value = int(np.ceil(n_free / repeats)) - 1
        else:
            max_value = len(repeats) - 1
    if order == 'ascending':
        base = torch.arange(max_value + 1)
    elif order == 'descending':
        base = torch.arange(max_value, -1, -1)
    elif order == 'random':
        base = torch.randperm(max_value + 1)
    else:
        raise ValueError("Accepted string values for 'order' are 'ascending', 'descending', and 'random'.")
    rep = torch.as_tensor(repeats, dtype=torch.long)
    deg = _tile_to_length(torch.repeat_interleave(base, rep)[:n_free], n_free)
    if conditioning_indices is None:
        return deg
    cond = [int(i) for i in (conditioning_indices.tolist() if hasattr(conditioning_indices, 'tolist')
                             else conditioning_indices)]
    free = [i for i in range(n_features) if i not in set(cond)]
    out = torch.empty(n_features, dtype=deg.dtype)
    out[cond] = -1
    out[free] = deg
    return out


def hidden_degrees(degrees_in: torch.Tensor, degrees_out: torch.Tensor, hidden_layers) -> List[torch.Tensor]:
    """nn/conditioners/made.py:366-434 (MADE._get_degrees_hidden)."""
    if hasattr(hidden_layers, 'tolist'):
        hidden_layers = hidden_layers.tolist()
    top = degrees_out.max()
    relevant = degrees_in < top
    if isinstance(hidden_layers, int):
        n_rel = int(relevant.sum())
        width = max(int(np.ceil((n_rel * len(degrees_out)) ** 0.5)), n_rel)
        hidden_layers = [width] * hidden_layers
    if isinstance(hidden_layers[0], int):
        motif = degrees_in[relevant]
        return [_tile_to_length(motif, w, err=(f'Hidden layer {i} is too small for the number of input features. '
                                               'Increase the size of the layer or explicitly pass the degrees '
                                               'for the hidden layers.'))
                for i, w in enumerate(hidden_layers)]
    out = [torch.as_tensor(h) for h in hidden_layers]
    for i, h in enumerate(out):
        if torch.any(h >= top):
            raise ValueError(f'The {i}-th hidden layer contain nodes with degrees that will be ignored '
                             'by the output layer.')
    return out


def ar_mask(deg_prev: torch.Tensor, deg_cur: torch.Tensor, strict: bool, dtype=None) -> torch.Tensor:
    """(out, in) connectivity mask.  nn/masked.py:90-108 with transpose=True; made.py:308-309."""
    m = (deg_cur[:, None] > deg_prev[None, :]) if strict else (deg_cur[:, None] >= deg_prev[None, :])
    return m.to(dtype or torch.get_default_dtype())


def made_masks(degrees_in, degrees_out, hidden_layers, dtype=None):
    """All layer masks of a MADE, input to output.  nn/conditioners/made.py:294-329."""
    hid = hidden_degrees(degrees_in, degrees_out, hidden_layers)
    chain = [degrees_in] + hid + [degrees_out]
    return [ar_mask(chain[i], chain[i + 1], strict=(i == len(chain) - 2), dtype=dtype)
            for i in range(len(chain) - 1)], chain


# ---------------------------------------------------------------------------
# conditioner arithmetic
# ---------------------------------------------------------------------------

def effective_weight(weight_v, weight_g, mask):
    """Masked weight-norm weight.  nn/masked.py:369-371 (compute_weight) + :433-439 (_ApplyMask)
    followed by the multiply in MaskedLinearFunc.forward, nn/masked.py:268-270."""
    w = torch._weight_norm(weight_v, weight_g, 0)
    w = w.clone()
    w[mask == 0.0] = 0.0          # also clears 0/0 NaNs of fully masked rows
    return w * mask


def made_forward(x, layers):
    """layers: list of (W_eff, bias).  ELU after every layer but the last.
    nn/conditioners/made.py:320-329,355-356; nn/masked.py:277."""
    h = x
    for i, (w, b) in enumerate(layers):
        h = F.linear(h, w, b)
        if i != len(layers) - 1:
            h = F.elu(h)
    return h


# ---------------------------------------------------------------------------
# transformers
# ---------------------------------------------------------------------------

class _Spec:
    def degrees_out(self, degrees_in):
        """affine.py:134, spline.py:317, sos.py:156: tile of the input degrees P times."""
        return degrees_in.tile((self.n_params_per_feature,))


@dataclass
class Affine(_Spec):
    """nn/transformers/affine.py:28-141, 281-363."""
    n_params_per_feature: int = 2

    def identity_params(self, n):
        return torch.zeros(2 * n)

    def _split(self, par):
        par = par.reshape(par.shape[0], 2, -1)
        return par[:, 0], par[:, 1]

    def forward(self, x, par):
        shift, log_scale = self._split(par)
        return x * torch.exp(log_scale) + shift, torch.sum(log_scale, dim=1)

    def inverse(self, y, par):
        shift, log_scale = self._split(par)
        return (y - shift) * torch.exp(-log_scale), -torch.sum(log_scale, dim=1)


@dataclass
class Shift(_Spec):
    """VolumePreservingShiftTransformer: nn/transformers/affine.py:148-275, 366-456."""
    periodic_indices: Optional[torch.Tensor] = None
    periodic_limits: Optional[torch.Tensor] = None
    n_params_per_feature: int = 1

    def identity_params(self, n):
        return torch.zeros(n)

    def _wrap(self, v):
        if self.periodic_indices is not None:
            lim = self.periodic_limits
            v[:, self.periodic_indices] = v[:, self.periodic_indices] % (lim[1] - lim[0]) + lim[0]
        return v

    def forward(self, x, par):
        y = self._wrap(x + par)
        return y, torch.zeros(x.shape[0], dtype=x.dtype, device=x.device)

    def inverse(self, y, par):
        x = self._wrap(y - par)
        return x, torch.zeros(y.shape[0], dtype=y.dtype, device=y.device)


@dataclass
class Spline(_Spec):
    """nn/transformers/spline.py:29-417 (module) and :424-650 (functional)."""
    x0: torch.Tensor = None
    xf: torch.Tensor = None
    n_bins: int = 8
    y0: Optional[torch.Tensor] = None
    yf: Optional[torch.Tensor] = None
    circular: bool = False
    identity_boundary_slopes: bool = False
    learn_lower_bound: bool = False
    learn_upper_bound: bool = False
    min_bin_size: float = 1e-4
    min_slope: float = 1e-4

    def __post_init__(self):
        if self.y0 is None:
            self.y0 = self.x0.detach()
        if self.yf is None:
            self.yf = self.xf.detach()
        if self.circular and (self.learn_lower_bound or self.learn_upper_bound):
            raise ValueError('Cannot instantiate a circular spline with learnable limits.')
        # The reference keeps these two as 0-dim tensors of the default dtype
        # (spline.py:160-161); they enter the arithmetic as such.
        self._min_bin = torch.as_tensor(self.min_bin_size)
        self._min_slope = torch.as_tensor(self.min_slope)

    @property
    def n_params_per_feature(self):
        """spline.py:166-182."""
        p = 3 * self.n_bins + 1 + int(self.learn_lower_bound) + int(self.learn_upper_bound)
        if self.identity_boundary_slopes:
            p -= 1 if self.circular else 2
        return p

    def identity_params(self, n):
        return torch.zeros(self.n_params_per_feature * n).to(self.x0)

    def unpack(self, par):
        """spline.py:319-417 (_get_parameters)."""
        K = self.n_bins
        par = par.reshape(par.shape[0], self.n_params_per_feature, -1)
        widths, heights = par[:, :K], par[:, K:2 * K]
        if self.identity_boundary_slopes:
            n_slopes = K - 1
        elif self.circular:
            n_slopes = K
        else:
            n_slopes = K + 1
        slopes = par[:, 2 * K:2 * K + n_slopes]
        shifts = None
        if self.circular:
            shifts = par[:, -1]
            if not self.identity_boundary_slopes:
                slopes = torch.cat([slopes, slopes[:, :1]], dim=1)
        if self.identity_boundary_slopes:
            z = torch.zeros_like(widths[:, :1])
            slopes = torch.cat([z, slopes, z], dim=1)
        min_interval = K * self._min_bin
        resc_w = self.xf - self.x0 - min_interval
        resc_h = self.yf - self.y0 - min_interval
        if self.learn_lower_bound or self.learn_upper_bound:
            scale = torch.exp(par[:, -1:])
            resc_w = resc_w * scale
            resc_h = resc_h * scale
        widths = F.softmax(widths, dim=1) * resc_w + self._min_bin
        heights = F.softmax(heights, dim=1) * resc_h + self._min_bin
        x0, y0 = self.x0, self.y0
        if self.learn_lower_bound and self.learn_upper_bound:
            x0 = x0 + par[:, -2]
            y0 = y0 + par[:, -2]
        elif self.learn_lower_bound:
            x0 = self.xf - resc_w.squeeze(1) - min_interval
            y0 = self.yf - resc_h.squeeze(1) - min_interval
        offset = torch.log(torch.exp(1. - self._min_slope) - 1.)
        slopes = F.softplus(slopes + offset) + self._min_slope
        return x0, y0, widths, heights, slopes, shifts

    def forward(self, x, par):
        """spline.py:184-241."""
        x0, y0, w, h, d, shifts = self.unpack(par)
        if self.circular:
            x = (x - x0 + shifts) % (self.xf - x0) + x0
        return rq_spline(x, x0, y0, w, h, d)

    def inverse(self, y, par):
        """spline.py:243-261."""
        x0, y0, w, h, d, shifts = self.unpack(par)
        x, ld = rq_spline_inverse(y, x0, y0, w, h, d)
        if shifts is not None:
            x = (x - x0 - shifts) % (self.xf - x0) + x0
        return x, ld


def _bin_lookup(t, x0, y0, widths, heights, slopes, inverse, return_bins=False):
    """spline.py:567-650 (_assign_bins).  Bin 0 / K+1 are the linear tails."""
    B, K, Fe = widths.shape
    cw = torch.cumsum(widths, dim=1)
    ch = torch.cumsum(heights, dim=1)
    if x0.dim() == 0:
        x0 = x0.unsqueeze(0)
    if y0.dim() == 0:
        y0 = y0.unsqueeze(0)
    kx = torch.empty(B, K + 3, Fe).to(x0)
    kx[:, 1] = x0
    kx[:, 2:-1] = x0.unsqueeze(-2) + cw
    ky = torch.empty(B, K + 3, Fe).to(x0)
    ky[:, 1] = y0
    ky[:, 2:-1] = y0.unsqueeze(-2) + ch
    dx = cw[:, -1] * 1000.
    kx[:, 0] = x0 - dx
    kx[:, -1] = kx[:, -2] + dx
    dy0 = slopes[:, 0] * dx
    ky[:, 0] = y0 - dy0
    dyf = slopes[:, -1] * dx
    ky[:, -1] = ky[:, -2] + dyf
    slopes = torch.cat([slopes[:, 0:1], slopes, slopes[:, -1:]], dim=1)
    dx = dx.unsqueeze(1)
    widths = torch.cat([dx, widths, dx], dim=1)
    heights = torch.cat([dy0.unsqueeze(1), heights, dyf.unsqueeze(1)], dim=1)
    bi = torch.arange(B).unsqueeze(-1)
    fi = torch.arange(Fe).repeat(B, 1)
    bins = torch.sum(t.unsqueeze(1) > (ky if inverse else kx), dim=1) - 1
    w = widths[bi, bins, fi]
    h = heights[bi, bins, fi]
    out = (w, h, kx[bi, bins, fi], ky[bi, bins, fi], slopes[bi, bins, fi], slopes[bi, bins + 1, fi], h / w)
    return out + (bins,) if return_bins else out


def _rq_logdet(dk, dk1, s, e, e1me, e2, inverse):
    """spline.py:546-564 (_compute_log_det_J)."""
    num = s**2 * (dk1 * e2 + 2 * s * e1me + dk * (1 - e)**2)
    den = (s + (dk1 + dk - 2 * s) * e1me)**2
    ld = torch.sum(torch.log(num / den), dim=1)
    return -ld if inverse else ld


def rq_spline(x, x0, y0, widths, heights, slopes):
    """spline.py:424-501 (neural_spline_transformer)."""
    w, h, xk, yk, dk, dk1, s = _bin_lookup(x, x0, y0, widths, heights, slopes, inverse=False)
    e = (x - xk) / w
    e1me = e * (1 - e)
    e2 = e**2
    num = h * (s * e2 + dk * e1me)
    den = s + (dk1 + dk - 2 * s) * e1me
    return yk + num / den, _rq_logdet(dk, dk1, s, e, e1me, e2, inverse=False)


def rq_spline_inverse(y, x0, y0, widths, heights, slopes):
    """spline.py:504-543 (neural_spline_transformer_inverse)."""
    w, h, xk, yk, dk, dk1, s = _bin_lookup(y, x0, y0, widths, heights, slopes, inverse=True)
    yr = y - yk
    q = dk1 + dk - 2 * s
    a = h * (s - dk) + yr * q
    b = h * dk - yr * q
    c = -s * yr
    e = 2 * c.div(-b - torch.sqrt(b**2 - 4 * a * c))
    x = e * w + xk
    return x, _rq_logdet(dk, dk1, s, e, e * (1 - e), e**2, inverse=True)


def spline_bins(spec: 'Spline', x, par, inverse=False):
    """Bin index (0 = left tail, 1..K bins, K+1 = right tail) the reference assigns; spline.py:622-625."""
    x0, y0, w, h, d, shifts = spec.unpack(par)
    if spec.circular and not inverse:
        x = (x - x0 + shifts) % (spec.xf - x0) + x0
    return _bin_lookup(x, x0, y0, w, h, d, inverse=inverse, return_bins=True)[-1]


@dataclass
class SOS(_Spec):
    """nn/transformers/sos.py:28-160, 207-306.  No inverse (sos.py:111-114)."""
    n_polynomials: int = 2

    def __post_init__(self):
        if self.n_polynomials < 2:
            raise ValueError('n_polynomials must be strictly greater than 1.')

    @property
    def n_params_per_feature(self):
        return 2 * self.n_polynomials + 1

    def identity_params(self, n):
        p = torch.zeros(self.n_params_per_feature, n)
        p[1::2].fill_(np.sqrt(1 / self.n_polynomials))
        return p.flatten()

    @staticmethod
    def coefficients(par):
        """sos.py:271-306."""
        a0, k0, k1 = par[:, 0], par[:, 1::2], par[:, 2::2]
        return [a0, torch.sum(k0**2, dim=1), torch.sum(k0 * k1, dim=1), torch.sum(k1**2, dim=1) / 3]

    def forward(self, x, par):
        """sos.py:207-235."""
        par = par.reshape(par.shape[0], self.n_params_per_feature, -1)
        c = self.coefficients(par)
        xp = [x, x * x]
        y = c[1].clone()
        g = c[1].clone()
        for deg, coef in enumerate(c[2:]):
            term = coef * xp[deg]
            y += term
            g += (deg + 2) * term
        y *= x
        y += c[0]
        return y, torch.sum(torch.log(g), dim=1)

    def inverse(self, y, par):
        raise NotImplementedError('Inversion of SOS polynomial transformer has not been implemented yet.')

    def vjp(self, x, par, grad_y):
        """Hand-written backward of the reference; the log-det cotangent is dropped.  sos.py:237-268."""
        par = par.reshape(par.shape[0], self.n_params_per_feature, -1)
        c = self.coefficients(par)
        dydx = c[1] + 2 * c[2] * x + 3 * c[3] * x * x
        gp = torch.empty_like(par)
        gp[:, 0] = 1.0
        k0, k1 = par[:, 1::2], par[:, 2::2]
        x1, x2 = x.unsqueeze(1), (x * x).unsqueeze(1)
        gp[:, 1::2] = k1 * x2 + 2 * k0 * x1
        gp[:, 2::2] = 2 / 3 * k1 * (x2 * x1) + k0 * x2
        return dydx * grad_y, (gp * grad_y.unsqueeze(1)).reshape(par.shape[0], -1)


@dataclass
class Moebius(_Spec):
    """nn/transformers/moebius.py:27-190, 374-478."""
    dimension: int = 3
    max_radius: float = 0.99
    unit_sphere: bool = False
    n_params_per_feature: int = 1

    def degrees_out(self, degrees_in):
        return degrees_in.detach().clone()

    def identity_params(self, n):
        return torch.zeros(n)

    def _map(self, x, w):
        B, n = x.shape
        x = x.reshape(B, -1, self.dimension)
        w = w.reshape(B, -1, self.dimension)
        d = self.dimension
        wn = torch.linalg.norm(w, dim=-1, keepdim=True)
        resc = self.max_radius / (1 + wn)
        if not self.unit_sphere:
            xn = torch.linalg.norm(x, dim=-1, keepdim=True)
            resc = xn * resc
        w = resc * w
        wn = resc * wn
        num = (1 - wn**2) if self.unit_sphere else (xn**2 - wn**2)
        diff = x - w
        dn = torch.linalg.norm(diff, dim=-1, keepdim=True)
        y = num / dn.pow(2) * diff - w
        num, dn = num.unsqueeze(-1), dn.unsqueeze(-1)
        outer = lambda a, b: torch.einsum('...i, ...j -> ...ij', a, b)   # utils/math.py:47-66
        dd = outer(diff, diff)
        eye = torch.eye(d).expand_as(dd)
        jac = num * (eye / dn.pow(2) - 2 / dn.pow(4) * dd)
        if not self.unit_sphere:
            xe = xn.unsqueeze(-1)
            jac2 = eye - outer(x, x) / xe**2
            jac = torch.einsum('...ij, ...jk -> ...ik', jac, jac2)
            jac = outer(y, x) / xe**2 + jac
        ld = torch.linalg.slogdet(jac)[1].sum(dim=-1)
        return y.reshape(B, n), ld

    def _map_only(self, x, w):
        """The map without its Jacobian (return_log_det_J=False, moebius.py:455-457), radius |x|."""
        B, n = x.shape
        x = x.reshape(B, -1, self.dimension)
        w = w.reshape(B, -1, self.dimension)
        wn = torch.linalg.norm(w, dim=-1, keepdim=True)
        xn = torch.linalg.norm(x, dim=-1, keepdim=True)
        resc = self.max_radius / (1 + wn)
        resc = xn * resc
        w, wn = resc * w, resc * wn
        diff = x - w
        return ((xn**2 - wn**2) / torch.linalg.norm(diff, dim=-1, keepdim=True).pow(2) * diff - w).reshape(B, n)

    def forward(self, x, par):
        return self._map(x, par)

    def inverse(self, y, par):
        """moebius.py:119-151: same map with the sign of the parameters flipped."""
        return self._map(y, -par)


@dataclass
class SymMoebius(_Spec):
    """SymmetrizedMoebiusTransformer: nn/transformers/moebius.py:193-372 (class), 481-550 (forward), 553-600
    (analytic inverse), 607-629 (closed-form log-det)."""
    dimension: int = 3
    max_radius: float = 0.99
    identity_eps: float = 1e-9
    n_params_per_feature: int = 1

    def degrees_out(self, degrees_in):
        return degrees_in.detach().clone()

    def identity_params(self, n):
        return (2 * torch.rand(n) - 1) * self.identity_eps

    def _unit_w(self, w):
        wn = torch.linalg.norm(w, dim=-1, keepdim=True)
        resc = self.max_radius / (1 + wn)
        return resc * w, (resc * wn)**2

    def _logdet(self, x_unit, w_unit, r2):
        d = self.dimension
        qy2 = r2 - (x_unit * w_unit).sum(dim=-1, keepdim=True)**2
        dV = (1 - r2) * (1 + r2)**(d - 1) / (4 * qy2 + (1 - r2)**2)**(d / 2)
        return torch.log(dV).squeeze(-1).sum(dim=1)

    def forward(self, x, par):
        B, n = x.shape
        plain = Moebius(self.dimension, self.max_radius, unit_sphere=False)
        f = plain._map_only(x, par) + plain._map_only(x, -par)
        x3, f3, w3 = (t.reshape(B, -1, self.dimension) for t in (x, f, par))
        xn = torch.linalg.norm(x3, dim=-1, keepdim=True)
        y = xn / torch.linalg.norm(f3, dim=-1, keepdim=True) * f3
        w_unit, r2 = self._unit_w(w3)
        return y.reshape(B, n), self._logdet(x3 / xn, w_unit, r2)

    def inverse(self, y, par):
        B, n = y.shape
        y3, w3 = y.reshape(B, -1, self.dimension), par.reshape(B, -1, self.dimension)
        yn = torch.linalg.norm(y3, dim=-1, keepdim=True)
        y_unit = y3 / yn
        w_unit, r2 = self._unit_w(w3)
        da = w_unit / torch.sqrt(r2)
        a = (y_unit * da).sum(dim=-1, keepdim=True)
        db = y_unit - a * da
        db = db / torch.linalg.norm(db, dim=-1, keepdim=True)
        a_inv = -a * (r2 + 1.0) / torch.sqrt(1 + r2**2 + r2 * (4 * a**2 - 2))
        b_inv = -torch.sqrt(1 - a_inv**2)
        x_unit = -(a_inv * da + b_inv * db)
        return (yn * x_unit).reshape(B, n), -self._logdet(x_unit, w_unit, r2)


@dataclass
class Mixed(_Spec):
    """nn/transformers/mixed.py:29-186."""
    transformers: list = field(default_factory=list)
    indices: list = field(default_factory=list)

    def __post_init__(self):
        if len(self.transformers) < 2:
            raise ValueError('The number of transformers must be greater than 1.')
        if len(self.transformers) != len(self.indices):
            raise ValueError('The number of elements in indices must equal that in transformers.')
        self.indices = [torch.as_tensor(i) for i in self.indices]
        lens = [len(t.identity_params(len(i))) for t, i in zip(self.transformers, self.indices)]
        self._splits = torch.cumsum(torch.tensor(lens[:-1]), dim=0)

    def identity_params(self, n):
        return torch.cat([t.identity_params(len(i)) for t, i in zip(self.transformers, self.indices)], dim=-1)

    def degrees_out(self, degrees_in):
        return torch.cat([t.degrees_out(degrees_in[i]) for t, i in zip(self.transformers, self.indices)], dim=-1)

    def _run(self, x, par, inverse):
        y = torch.empty_like(x)
        total = 0.0
        for t, idx, p in zip(self.transformers, self.indices, torch.tensor_split(par, self._splits, dim=1)):
            y[:, idx], ld = (t.inverse if inverse else t.forward)(x[:, idx], p)
            total = total + ld
        return y, total

    def forward(self, x, par):
        return self._run(x, par, False)

    def inverse(self, y, par):
        return self._run(y, par, True)


# ---------------------------------------------------------------------------
# autoregressive flow
# ---------------------------------------------------------------------------

class PeriodicEmbed:
    """PeriodicEmbedding: nn/embeddings/mafembed.py:65-172 (non-periodic features first, then (cos, sin) pairs)."""

    def __init__(self, n_features_in, limits, periodic_indices=None):
        self.n_features_in = n_features_in
        self.limits = torch.as_tensor(limits)
        self.periodic_indices = torch.arange(n_features_in) if periodic_indices is None else torch.as_tensor(periodic_indices)
        allidx = torch.arange(n_features_in)
        self.nonperiodic_indices = allidx[~torch.isin(allidx, self.periodic_indices)]

    def __call__(self, x):
        scale = 2 * torch.pi / (self.limits[1] - self.limits[0])
        xp = (x[:, self.periodic_indices] - self.limits[0]) * scale
        return torch.cat([x[:, self.nonperiodic_indices],
                          torch.stack([torch.cos(xp), torch.sin(xp)], dim=2).reshape(x.shape[0], -1)], dim=1)

    def get_degrees_out(self, degrees_in):
        return torch.cat([degrees_in[self.nonperiodic_indices], degrees_in[self.periodic_indices].repeat_interleave(2)])


class MafOracle:
    """One MAF layer evaluated from a reference-compatible ``state_dict``.

    Follows nn/flows/maf.py:82-173 (construction), nn/flows/autoregressive.py:144-177
    (forward) and :179-229 (inverse: one full conditioner + transformer pass per
    distinct degree, log-det of the last pass).
    """

    def __init__(self, degrees_in, transformer, hidden_layers=2, weight_norm=True, embedding=None):
        degrees_in = torch.as_tensor(degrees_in)
        lo, hi = int(degrees_in.min()), int(degrees_in.max())
        if set(degrees_in.tolist()) != set(range(lo, hi + 1)) or lo not in (-1, 0):
            raise ValueError('degrees_in must assume consecutive values starting '
                             'from 0 (or -1 for conditioning input features).')
        self.degrees_in = degrees_in
        self.transformer = transformer
        self.weight_norm = weight_norm
        self.embedding = embedding
        deg_embedded = degrees_in if embedding is None else embedding.get_degrees_out(degrees_in)
        self.degrees_out = transformer.degrees_out(degrees_in[degrees_in != -1])
        self.masks, self.degree_chain = made_masks(deg_embedded, self.degrees_out, hidden_layers)
        self.groups = [(degrees_in == d).nonzero().flatten() for d in range(hi + 1)]
        self.mapped = torch.cat(self.groups).sort().values
        allidx = torch.arange(len(degrees_in))
        self.fixed = allidx[~torch.isin(allidx, self.mapped)]
        self.layers = None

    def load(self, state_dict, prefix=''):
        """Pick ``_conditioner.layers.{0,2,4,..}.{weight_g,weight_v|weight,bias}`` out of a state dict."""
        self.layers = []
        for li in range(len(self.masks)):
            key = f'{prefix}_conditioner.layers.{2 * li}.'
            mask = self.masks[li].to(state_dict[key + 'bias'].dtype)
            if self.weight_norm:
                w = effective_weight(state_dict[key + 'weight_v'], state_dict[key + 'weight_g'], mask)
            else:
                w = state_dict[key + 'weight'] * mask
            self.layers.append((w, state_dict[key + 'bias']))
        return self

    def parameters_of(self, x):
        if self.embedding is not None:
            x = self.embedding(x)
        return made_forward(x, self.layers)

    def forward(self, x):
        par = self.parameters_of(x)
        if len(self.fixed) == 0:
            return self.transformer.forward(x, par)
        y = torch.empty_like(x)
        y[:, self.fixed] = x[:, self.fixed]
        y[:, self.mapped], ld = self.transformer.forward(x[:, self.mapped], par)
        return y, ld

    def inverse(self, y):
        x = torch.zeros_like(y)
        yt = y
        if len(self.fixed) > 0:
            x[:, self.fixed] = y[:, self.fixed]
            yt = y[:, self.mapped]
        ld = None
        for idx in self.groups:
            sel = torch.zeros(len(self.degrees_in), dtype=torch.bool)
            sel[idx] = True
            sel_t = sel[self.mapped] if len(self.fixed) > 0 else sel
            xt, ld = self.transformer.inverse(yt, self.parameters_of(x.clone()))
            x[:, sel] = xt[:, sel_t]
        return x, ld


def sequential(flows, x, inverse=False):
    """nn/flows/sequential.py:50-68."""
    total = torch.zeros(x.size(0)).to(x)
    for f in (reversed(flows) if inverse else flows):
        x, ld = f.inverse(x) if inverse else f.forward(x)
        total += ld
    return x, total


def kl_loss(potentials_b, log_det_J=None, potentials_a=None, log_weights=None, ignore_nan=False):
    """loss.py:76-140 (BoltzmannKLDivLoss.forward): reduced work loss.py:125-129, weighted sum loss.py:132-136,
    mean loss.py:138-140."""
    r = potentials_b
    if log_det_J is not None:
        r = r - log_det_J
    if potentials_a is not None:
        r = r - potentials_a
    if log_weights is None:
        return torch.nanmean(r) if ignore_nan else torch.mean(r)
    w = F.softmax(log_weights, dim=0)
    return torch.nansum(w * r) if ignore_nan else torch.sum(w * r)
