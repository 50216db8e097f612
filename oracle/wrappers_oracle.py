"""CPU restatement of the reference's wrapper flows (TEST INFRASTRUCTURE ONLY: imported by tests/, never by the
product).  Pinned to the reference by oracle/check_against_reference.py (bitwise, fp32 and fp64).

PartialFlow            tfep/nn/flows/partial.py:88-121
CenteredCentroidFlow   tfep/nn/flows/centroid.py:107-263
OrientedFlow           tfep/nn/flows/oriented.py:66-225, frame: tfep/utils/geometry.py:71-124, 185-276, 296-411

A wrapped flow is any object with ``forward(x) -> (y, ld)`` and ``inverse(y) -> (x, ld)`` (e.g. flow_oracle.MafOracle).
"""

import torch


class Partial:
    """partial.py:88-121: the wrapped flow sees the features that are not fixed; the fixed ones are copied through."""

    def __init__(self, flow, fixed_indices, return_partial=False):
        self.flow, self.return_partial = flow, return_partial
        self.fixed = torch.as_tensor(fixed_indices)

    def _pass(self, x, inverse):
        if len(self.fixed) > 0:
            fixed = set(self.fixed.tolist())
            prop = torch.tensor([i for i in range(x.shape[1]) if i not in fixed])
            y = torch.empty_like(x)
            y[:, self.fixed] = x[:, self.fixed]
            x = x[:, prop]
        out = self.flow.inverse(x) if inverse else self.flow.forward(x)
        if self.return_partial:
            return out
        if len(self.fixed) > 0:
            y[:, prop] = out[0]
        else:
            y = out[0]
        return y, out[1]

    def forward(self, x):
        return self._pass(x, False)

    def inverse(self, y):
        return self._pass(y, True)


class Centroid(Partial):
    """centroid.py:107-263."""

    def __init__(self, flow, space_dimension, subset_point_indices=None, weights=None, fixed_point_idx=0, origin=None,
                 translate_back=True, return_partial=False):
        self.dim = space_dimension
        self.origin = torch.zeros(space_dimension) if origin is None else torch.as_tensor(origin)
        self.subset = None if subset_point_indices is None else torch.as_tensor(subset_point_indices)
        point = fixed_point_idx if self.subset is None else int(self.subset[fixed_point_idx])
        super().__init__(flow, [point * space_dimension + i for i in range(space_dimension)], return_partial)
        self.weights = None
        if weights is not None:
            w = torch.as_tensor(weights)
            self.weights = (w / torch.sum(w)).unsqueeze(1)
        self.fixed_point_idx, self.translate_back = fixed_point_idx, translate_back

    def _centroid(self, pts, exclude_fixed=False):
        if self.subset is not None:
            pts = pts[:, self.subset]
        if self.weights is None:
            c, fw = torch.mean(pts, dim=1), 1 / pts.shape[1]
        else:
            c, fw = torch.sum(pts * self.weights.to(pts), dim=1), self.weights.to(pts)[self.fixed_point_idx]
        if exclude_fixed:
            return c - pts[:, self.fixed_point_idx] * fw, fw
        return c

    def _transform(self, x, inverse):
        B = x.shape[0]
        pts = x.reshape(B, -1, self.dim)
        t = (self.origin.to(x) - self._centroid(pts)).unsqueeze(1)
        xt = (pts + t).reshape(B, -1)
        y, ld = super()._pass(xt, inverse)
        if self.return_partial:
            return y, ld
        if self.subset is None or len(self.subset) > 1:
            c, fw = self._centroid(y.reshape(B, -1, self.dim), exclude_fixed=True)
            y[:, self.fixed] = (self.origin.to(x) - c) / fw
        if self.translate_back:
            y = (y.reshape(B, -1, self.dim) - t).reshape(B, -1)
        return y, ld

    def forward(self, x):
        return self._transform(x, False)

    def inverse(self, y):
        return self._transform(y, True)


_AXES = {'x': [1.0, 0.0, 0.0], 'y': [0.0, 1.0, 0.0], 'z': [0.0, 0.0, 1.0]}


def _dot(a, b, keepdim=False):
    return (a * b).sum(dim=-1, keepdim=keepdim)


def _angle_cos(a, b):
    c = _dot(a, b) / (torch.linalg.vector_norm(a, dim=-1) * torch.linalg.vector_norm(b, dim=-1))
    return torch.clamp(c, min=-1, max=1)


def rotation_matrix(angles, directions):
    """geometry.py:185-236 (Rodrigues; same operation order as the reference)."""
    B = len(angles)
    sina, cosa = torch.sin(angles), torch.cos(angles)
    k = torch.nn.functional.normalize(directions, dim=-1)
    if k.dim() < 2:
        k = k.unsqueeze(0)
    cosa = cosa.unsqueeze(-1).unsqueeze(-1)
    R = cosa * torch.eye(3).expand(B, 3, 3).to(cosa)
    R = R + (1 - cosa) * torch.einsum('...i, ...j -> ...ij', k, k)
    sk = sina.unsqueeze(-1) * k
    z = torch.zeros_like(angles)
    cross = torch.stack([torch.stack([z, -sk[:, 2], sk[:, 1]]), torch.stack([sk[:, 2], z, -sk[:, 0]]),
                         torch.stack([-sk[:, 1], sk[:, 0], z])])
    return R + cross.permute(2, 0, 1)


def frame_rotation(axis_pos, plane_pos, axis, plane_axis, plane_normal):
    """geometry.py:296-411 with project_on_positive_axis=False."""
    rv = torch.cross(axis_pos, axis.unsqueeze(0), dim=1)
    par = torch.isclose(rv, torch.zeros(1, dtype=rv.dtype)).all(dim=1)
    rv[par] = torch.cross(plane_axis, axis, dim=0)
    a1 = torch.acos(_angle_cos(axis_pos, axis))
    a1 = a1 - torch.pi * (a1 > torch.pi / 2).to(a1.dtype)
    r1 = rotation_matrix(a1, rv)
    p = torch.bmm(plane_pos.unsqueeze(1), r1.permute(0, 2, 1)).squeeze(1)
    p = p - axis * _dot(p, axis, keepdim=True)
    a2 = torch.asin(_angle_cos(p, plane_normal))
    sign = -torch.sign(_dot(p, plane_axis))
    r2 = rotation_matrix(sign * a2, axis)
    return torch.bmm(r2, r1)


class Oriented(Partial):
    """oriented.py:66-225."""

    def __init__(self, flow, axis_point_idx=None, plane_point_idx=None, axis='x', plane='xy', round_off_imprecisions=True,
                 rotate_back=True, return_partial=False):
        if axis_point_idx is None:
            axis_point_idx = 0 if plane_point_idx != 0 else 1
        if plane_point_idx is None:
            plane_point_idx = 0 if axis_point_idx != 0 else 1
        self.axis = torch.tensor(_AXES[axis])
        self.plane_axis = torch.tensor([_AXES[n] for n in 'xyz' if n not in axis and n in plane][0])
        self.normal = torch.cross(self.axis, self.plane_axis, dim=0)
        fixed = [3 * axis_point_idx + i for i in range(3) if self.axis[i] == 0.0] + \
                [3 * plane_point_idx + i for i in range(3) if self.normal[i] != 0.0]
        super().__init__(flow, fixed, return_partial)
        self.ap, self.pp = axis_point_idx, plane_point_idx
        self.round_off, self.rotate_back = round_off_imprecisions, rotate_back

    def _transform(self, x, inverse):
        B = x.shape[0]
        pts = x.reshape(B, -1, 3)
        R = frame_rotation(pts[:, self.ap], pts[:, self.pp], self.axis.to(x), self.plane_axis.to(x), self.normal.to(x))
        xr = torch.bmm(pts, R.permute(0, 2, 1)).reshape(B, -1)
        if self.round_off:
            xr[:, self.fixed] = 0.0
        y, ld = super()._pass(xr, inverse)
        if self.return_partial:
            return y, ld
        if self.rotate_back:
            y = torch.bmm(y.reshape(B, -1, 3), R).reshape(B, -1)
        return y, ld

    def forward(self, x):
        return self._transform(x, False)

    def inverse(self, y):
        return self._transform(y, True)
