"""Rigid-frame helpers for the wrapper flows (reference: tfep/utils/geometry.py:71-124, 185-276, 279-411 and the
point <-> feature reshapes of tfep/utils/misc.py).

Plain differentiable tensor algebra on ``(batch, n_points, 3)`` coordinates: these run once per flow evaluation on a
few points per sample, either side of the MAF kernels, and must carry autograd through the rotation (the frame
depends on the sample).
"""

import torch

_AXES = {'x': (1.0, 0.0, 0.0), 'y': (0.0, 1.0, 0.0), 'z': (0.0, 0.0, 1.0)}


def get_axis_from_name(name):
    """Unit vector of the axis 'x', 'y' or 'z' (reference geometry.py:279-293)."""
    return torch.tensor(_AXES[name])


def flattened_to_atom(x, space_dimension=3):
    """(batch, n_points * dim) -> (batch, n_points, dim)."""
    return x.reshape(x.shape[0], -1, space_dimension)


def atom_to_flattened(x):
    """(batch, n_points, dim) -> (batch, n_points * dim)."""
    return x.reshape(x.shape[0], -1)


def atom_to_flattened_indices(point_indices, space_dimension=3):
    """Feature indices of the coordinates of the given points."""
    point_indices = torch.as_tensor(point_indices)
    offsets = torch.arange(space_dimension)
    return (point_indices[:, None] * space_dimension + offsets[None, :]).reshape(-1)


def _unit_dot(a, b):
    """cos of the angle between the rows of a and the (broadcast) vector(s) b, clamped to [-1, 1]."""
    cos = (a * b).sum(dim=-1) / (torch.linalg.vector_norm(a, dim=-1) * torch.linalg.vector_norm(b, dim=-1))
    return torch.clamp(cos, min=-1, max=1)


def vector_vector_angle(x1, x2):
    """Angle in [0, pi] between vectors (reference geometry.py:71-99)."""
    return torch.acos(_unit_dot(x1, x2))


def vector_plane_angle(x, plane_normal):
    """Angle in [-pi/2, pi/2] between vectors and the plane with the given normal (reference geometry.py:102-124)."""
    return torch.asin(_unit_dot(x, plane_normal))


def rotation_matrix_3d(angles, directions):
    """Rodrigues matrices ``cos a I + (1 - cos a) k k^T + sin a [k]_x`` for a batch of angles about ``directions``
    (batch, 3) or (3,) (reference geometry.py:185-236).  Returns (batch, 3, 3)."""
    k = torch.nn.functional.normalize(directions, dim=-1)
    if k.dim() < 2:
        k = k.unsqueeze(0).expand(len(angles), 3)
    c, s = torch.cos(angles), torch.sin(angles)
    t = 1 - c
    kx, ky, kz = k.unbind(dim=-1)
    rows = [c + t * kx * kx, t * kx * ky - s * kz, t * kx * kz + s * ky,
            t * ky * kx + s * kz, c + t * ky * ky, t * ky * kz - s * kx,
            t * kz * kx - s * ky, t * kz * ky + s * kx, c + t * kz * kz]
    return torch.stack(rows, dim=-1).reshape(-1, 3, 3)


def batchwise_rotate(x, rotation_matrices, inverse=False):
    """Rotate the points (batch, n_points, 3) of every sample by its matrix (or its transpose if ``inverse``)."""
    return torch.bmm(x, rotation_matrices if inverse else rotation_matrices.transpose(1, 2))


def reference_frame_rotation_matrix(axis_atom_positions, plane_atom_positions, axis, plane_axis, plane_normal=None,
                                    project_on_positive_axis=False):
    """Rotation that brings one point on ``axis`` and a second one on the plane spanned by ``axis`` and
    ``plane_axis`` (reference geometry.py:296-411): first the rotation about ``p x axis`` by the angle between
    the point and the axis (to the nearest of the two half-axes unless ``project_on_positive_axis``), then the
    rotation about ``axis`` that zeroes the out-of-plane component of the second point."""
    if plane_normal is None:
        plane_normal = torch.linalg.cross(axis, plane_axis, dim=0)
    rot_vec = torch.linalg.cross(axis_atom_positions, axis.unsqueeze(0).expand_as(axis_atom_positions), dim=1)
    parallel = torch.isclose(rot_vec, torch.zeros(1, dtype=rot_vec.dtype, device=rot_vec.device)).all(dim=1)
    fallback = torch.linalg.cross(plane_axis, axis, dim=0)          # any direction orthogonal to the axis will do
    rot_vec = torch.where(parallel.unsqueeze(1), fallback.unsqueeze(0).to(rot_vec), rot_vec)
    a1 = vector_vector_angle(axis_atom_positions, axis)
    if not project_on_positive_axis:
        a1 = a1 - torch.pi * (a1 > torch.pi / 2).to(a1.dtype)
    r1 = rotation_matrix_3d(a1, rot_vec)
    p = batchwise_rotate(plane_atom_positions.unsqueeze(1), r1).squeeze(1)
    p = p - axis * (p * axis).sum(dim=-1, keepdim=True)              # component orthogonal to the axis
    a2 = vector_plane_angle(p, plane_normal)
    sign = -torch.sign((p * plane_axis).sum(dim=-1))
    r2 = rotation_matrix_3d(sign * a2, axis)
    return torch.bmm(r2, r1)
