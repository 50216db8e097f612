"""Host side of the fused pre / post kernels of the wrapper flows (tfep_b200/csrc/frames.cu): inference only.

``CentroidFrame`` / ``OrientedFrame`` hold the integer tables of one wrapper (propagated columns, their inverse map,
the points defining the centroid) on the device and run ``pre`` (input row -> contiguous propagated features in the
new frame + the per-sample frame) and ``post`` (flow output -> full row in the original frame).
"""

import ctypes
import os

import torch

from . import _lib
from ._lib import check, dtype_code, ptr, stream_ptr


def usable(x):
    """The kernels serve CUDA fp32 / fp64 rows outside autograd (TFEPB_NO_FRAME_KERNELS=1 forces the tensor-algebra
    path: development A/B timing, scripts/bench_wrapped.py)."""
    if os.environ.get('TFEPB_NO_FRAME_KERNELS'):
        return False
    return x.is_cuda and x.dtype in (torch.float32, torch.float64) and x.dim() == 2


class _Frame:
    def __init__(self, n_features, fixed_indices):
        fixed = set(int(i) for i in fixed_indices)
        prop = [c for c in range(n_features) if c not in fixed]
        inverse = [-1] * n_features
        for j, c in enumerate(prop):
            inverse[c] = j
        self.n_features, self.n_prop = n_features, len(prop)
        self._host = dict(prop=torch.tensor(prop, dtype=torch.int32), inverse=torch.tensor(inverse, dtype=torch.int32))
        self._dev = {}

    def tables(self, device, dtype):
        key = (str(device), dtype)
        if key not in self._dev:
            self._dev[key] = {k: (v.to(device=device, dtype=dtype) if v.is_floating_point() else v.to(device))
                              for k, v in self._host.items()}
        return self._dev[key]


class CentroidFrame(_Frame):
    def __init__(self, n_features, space_dimension, fixed_indices, subset_points, weights, fixed_slot, fixed_point, origin,
                 restore, translate_back):
        super().__init__(n_features, fixed_indices)
        self.dim = int(space_dimension)
        self.n_points = n_features // self.dim if subset_points is None else len(subset_points)
        if subset_points is not None:
            self._host['points'] = torch.as_tensor(subset_points).to(torch.int32)
        if weights is not None:
            self._host['weights'] = torch.as_tensor(weights).flatten().double()
        self.origin = [float(v) for v in origin]
        self.fixed_slot, self.fixed_point = int(fixed_slot), int(fixed_point)
        self.restore, self.translate_back = bool(restore), bool(translate_back)

    def _args(self, x, y, out, shift):
        tb = self.tables(x.device, x.dtype)
        return _lib.CentroidArgs(
            dtype=dtype_code(x), batch=x.shape[0], n_features=self.n_features, space_dimension=self.dim,
            x=x.data_ptr(), ldx=x.stride(0), y_propagated=None if y is None else y.data_ptr(),
            ldy=0 if y is None else y.stride(0), out=out.data_ptr(), ldout=out.stride(0), shift=shift.data_ptr(),
            n_propagated=self.n_prop, propagated_columns=tb['prop'].data_ptr(), column_to_propagated=tb['inverse'].data_ptr(),
            centroid_points=tb['points'].data_ptr() if 'points' in tb else None, n_centroid_points=self.n_points,
            weights=tb['weights'].data_ptr() if 'weights' in tb else None,
            origin=(ctypes.c_double * 4)(*(self.origin + [0.0] * (4 - len(self.origin)))),
            fixed_point=self.fixed_point, fixed_slot=self.fixed_slot, restore_fixed_point=int(self.restore),
            translate_back=int(self.translate_back))

    def pre(self, x):
        x = x.contiguous()
        out = torch.empty(x.shape[0], self.n_prop, dtype=x.dtype, device=x.device)
        shift = torch.empty(x.shape[0], self.dim, dtype=x.dtype, device=x.device)
        if x.shape[0] > 0:
            a = self._args(x, None, out, shift)
            with torch.cuda.device(x.device):
                check(_lib.load().tfepb_centroid_pre(ctypes.byref(a), stream_ptr(x)))
        return x, out, shift

    def post(self, x, y_prop, shift):
        y_prop = y_prop.contiguous()
        out = torch.empty_like(x)
        if x.shape[0] > 0:
            a = self._args(x, y_prop, out, shift)
            with torch.cuda.device(x.device):
                check(_lib.load().tfepb_centroid_post(ctypes.byref(a), stream_ptr(x)))
        return out


class OrientedFrame(_Frame):
    def __init__(self, n_features, fixed_indices, axis_point, plane_point, axis, plane_axis, round_off, rotate_back):
        super().__init__(n_features, fixed_indices)
        self.axis_point, self.plane_point = int(axis_point), int(plane_point)
        self.axis, self.plane_axis = int(axis), int(plane_axis)
        self.round_off, self.rotate_back = bool(round_off), bool(rotate_back)

    def _args(self, x, y, out, rot):
        tb = self.tables(x.device, x.dtype)
        return _lib.OrientedArgs(
            dtype=dtype_code(x), batch=x.shape[0], n_features=self.n_features, n_propagated=self.n_prop,
            x=x.data_ptr(), ldx=x.stride(0), y_propagated=None if y is None else y.data_ptr(),
            ldy=0 if y is None else y.stride(0), out=out.data_ptr(), ldout=out.stride(0), rotation=rot.data_ptr(),
            propagated_columns=tb['prop'].data_ptr(), column_to_propagated=tb['inverse'].data_ptr(),
            axis_point=self.axis_point, plane_point=self.plane_point, axis=self.axis, plane_axis=self.plane_axis,
            round_off_imprecisions=int(self.round_off), rotate_back=int(self.rotate_back))

    def pre(self, x):
        x = x.contiguous()
        out = torch.empty(x.shape[0], self.n_prop, dtype=x.dtype, device=x.device)
        rot = torch.empty(x.shape[0], 9, dtype=x.dtype, device=x.device)
        if x.shape[0] > 0:
            a = self._args(x, None, out, rot)
            with torch.cuda.device(x.device):
                check(_lib.load().tfepb_oriented_pre(ctypes.byref(a), stream_ptr(x)))
        return x, out, rot

    def post(self, x, y_prop, rot):
        y_prop = y_prop.contiguous()
        out = torch.empty_like(x)
        if x.shape[0] > 0:
            a = self._args(x, y_prop, out, rot)
            with torch.cuda.device(x.device):
                check(_lib.load().tfepb_oriented_post(ctypes.byref(a), stream_ptr(x)))
        return out
