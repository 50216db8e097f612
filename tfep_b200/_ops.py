"""Tensor-level wrappers around the C ABI (raw pointers, current CUDA stream) and the autograd
functions built from them.  PyTorch is used for memory, streams and autograd bookkeeping only."""

import ctypes

import torch

from . import _lib
from ._lib import (ACT_ELU, ACT_NONE, LinearBwdInputArgs, LinearBwdWeightArgs, LinearFwdArgs, SplineCfg, TxGrads, TxIo,
                   check, dtype_code, ptr, require_cuda, stream_ptr)


def _rows(t):
    """2-D tensor with unit column stride (row-major, arbitrary leading dimension)."""
    if t.dim() != 2:
        raise _lib.TfepB200Error(f'expected a 2-D tensor, got shape {tuple(t.shape)}')
    if t.shape[1] > 1 and t.stride(1) != 1:
        t = t.contiguous()
    if t.shape[0] > 1 and t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


def _ld(t):
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)


# ---------------------------------------------------------------------------------------------
# masked linear layers
# ---------------------------------------------------------------------------------------------

def linear_forward(x, w, bias=None, activation=ACT_NONE, k_ranges=None, out=None):
    """y = act(x w^T + bias); see tfepb_masked_linear_forward."""
    require_cuda(x, w, bias)
    x, w = _rows(x), _rows(w)
    B, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K, (w.shape, x.shape)
    if out is not None:
        y = out
    else:
        # leading dimension padded to 16 bytes: the 128 x 128 GEMM kernel needs 16-byte aligned rows of the next layer's input
        pad = 16 // x.element_size()
        y = torch.empty((B, (N + pad - 1) // pad * pad), dtype=x.dtype, device=x.device)[:, :N]
    if B == 0:
        return y
    a = LinearFwdArgs(dtype=dtype_code(x), batch=B, in_features=K, out_features=N,
                      x=x.data_ptr(), ldx=_ld(x), w=w.data_ptr(), ldw=_ld(w),
                      bias=None if bias is None else bias.data_ptr(), y=y.data_ptr(), ldy=_ld(y),
                      activation=activation, reserved=0,
                      k_ranges=None if k_ranges is None else k_ranges.data_ptr())
    with torch.cuda.device(x.device):
        check(_lib.load().tfepb_masked_linear_forward(ctypes.byref(a), stream_ptr(x)))
    return y


def linear_backward_input(grad_y, w, act_out=None, n_ranges=None, out=None, accumulate=False):
    """grad_x = (grad_y w) [* ELU'(act_out)]; see tfepb_masked_linear_backward_input."""
    require_cuda(grad_y, w, act_out)
    grad_y, w = _rows(grad_y), _rows(w)
    B, N = grad_y.shape
    K = w.shape[1]
    if out is not None:
        gx = out
    else:
        pad = 16 // grad_y.element_size()
        gx = torch.empty((B, (K + pad - 1) // pad * pad), dtype=grad_y.dtype, device=grad_y.device)[:, :K]
    if B == 0:
        return gx
    if act_out is not None:
        act_out = _rows(act_out)
    a = LinearBwdInputArgs(dtype=dtype_code(grad_y), batch=B, in_features=K, out_features=N,
                           grad_y=grad_y.data_ptr(), ldgy=_ld(grad_y), w=w.data_ptr(), ldw=_ld(w),
                           grad_x=gx.data_ptr(), ldgx=_ld(gx),
                           act_out=None if act_out is None else act_out.data_ptr(),
                           ldact=0 if act_out is None else _ld(act_out),
                           accumulate=int(accumulate), reserved=0,
                           n_ranges=None if n_ranges is None else n_ranges.data_ptr())
    with torch.cuda.device(grad_y.device):
        check(_lib.load().tfepb_masked_linear_backward_input(ctypes.byref(a), stream_ptr(grad_y)))
    return gx


def linear_backward_weight(grad_y, x, need_bias=True, n_ranges=None):
    """(grad_w, grad_bias) = (grad_y^T x, sum_b grad_y); see tfepb_masked_linear_backward_weight.  With ``n_ranges``
    (the per-column-tile row ranges of a staircase mask) tiles the mask zeroes anyway are skipped."""
    require_cuda(grad_y, x)
    grad_y, x = _rows(grad_y), _rows(x)
    B, N = grad_y.shape
    K = x.shape[1]
    gw = torch.zeros((N, K), dtype=x.dtype, device=x.device)
    gb = torch.zeros((N,), dtype=x.dtype, device=x.device) if need_bias else None
    if B == 0:
        return gw, gb
    a = LinearBwdWeightArgs(dtype=dtype_code(x), batch=B, in_features=K, out_features=N,
                            grad_y=grad_y.data_ptr(), ldgy=_ld(grad_y), x=x.data_ptr(), ldx=_ld(x),
                            grad_w=gw.data_ptr(), ldgw=K, grad_bias=None if gb is None else gb.data_ptr(),
                            n_ranges=None if n_ranges is None else n_ranges.data_ptr())
    with torch.cuda.device(x.device):
        check(_lib.load().tfepb_masked_linear_backward_weight(ctypes.byref(a), stream_ptr(x)))
    return gw, gb


class MadeFunction(torch.autograd.Function):
    """All layers of a MADE conditioner as one autograd node.

    forward : h_0 = x, h_l = ELU(h_{l-1} W_l^T + b_l), out = h_{L-1} W_L^T + b_L     (made.py:294-329)
    backward: nn/masked.py:280-302 per layer, ELU' fused into the dX GEMM epilogue.
    Inputs are the PACKED effective weights (mask folded in); gradients w.r.t. them are dense and
    flow back to (g, v) through the PyTorch graph of the weight normalisation.
    """

    @staticmethod
    def forward(ctx, x, n_layers, k_ranges, n_ranges, *wb):
        ws, bs = wb[:n_layers], wb[n_layers:]
        acts = [x]
        h = x
        for l in range(n_layers):
            last = l == n_layers - 1
            h = linear_forward(h, ws[l], bs[l], ACT_NONE if last else ACT_ELU,
                               None if k_ranges is None else k_ranges[l])
            if not last:
                acts.append(h)
        ctx.save_for_backward(*acts, *ws)
        ctx.n_layers = n_layers
        ctx.n_ranges = n_ranges
        return h

    @staticmethod
    def backward(ctx, grad_out):
        L = ctx.n_layers
        saved = ctx.saved_tensors
        acts, ws = saved[:L], saved[L:]
        g = grad_out
        gws, gbs = [None] * L, [None] * L
        gx = None
        for l in range(L - 1, -1, -1):
            if ctx.needs_input_grad[4 + l] or ctx.needs_input_grad[4 + L + l]:
                gws[l], gbs[l] = linear_backward_weight(g, acts[l],
                                                        n_ranges=None if ctx.n_ranges is None else ctx.n_ranges[l])
            if l > 0:
                g = linear_backward_input(g, ws[l], act_out=acts[l],
                                          n_ranges=None if ctx.n_ranges is None else ctx.n_ranges[l])
            elif ctx.needs_input_grad[0]:
                gx = linear_backward_input(g, ws[0], n_ranges=None if ctx.n_ranges is None else ctx.n_ranges[0])
        return (gx, None, None, None, *gws, *gbs)


class PackedWeightFunction(torch.autograd.Function):
    """(W_packed, b_packed) of a weight-normalised masked layer in one launch, with its VJP in one launch (tfepb_wn_pack):
    ``W_packed[r, c] = (M o g v / |v|)[row_perm[r], col_perm[c]]`` (row_perm -1 = zero row), leading dimension padded to
    16 bytes.  Same values and gradients as ``masked.effective_weight`` followed by ``MadePlan.pack``."""

    @staticmethod
    def forward(ctx, v, g, bias, mask, row_perm, col_perm, onto):
        R0, C = v.shape
        R = R0 if row_perm is None else row_perm.numel()
        ldo = (C + 3) // 4 * 4
        out = torch.empty((R, ldo), dtype=torch.float32, device=v.device)
        b_out = torch.empty(R, dtype=torch.float32, device=v.device) if bias is not None else None
        with torch.cuda.device(v.device):
            check(_lib.load().tfepb_wn_pack(ptr(v), _ld(v), ptr(g), ptr(mask), 0 if mask is None else _ld(mask), ptr(bias), R0, C,
                                            ptr(row_perm), R, ptr(col_perm), ptr(out), ldo, ldo, ptr(b_out), stream_ptr(v)))
        ctx.save_for_backward(v, g, mask, row_perm, col_perm)
        ctx.onto, ctx.has_bias = onto, bias is not None
        w = out[:, :C]
        return w, b_out

    @staticmethod
    def backward(ctx, gw, gb):
        v, g, mask, row_perm, col_perm = ctx.saved_tensors
        R0, C = v.shape
        R = R0 if row_perm is None else row_perm.numel()
        gw = _rows(gw.contiguous() if gw.stride(1) != 1 else gw)
        alloc = torch.empty if ctx.onto else torch.zeros
        gv = alloc((R0, C), dtype=torch.float32, device=v.device)
        gg = alloc(g.shape, dtype=torch.float32, device=v.device)
        want_b = ctx.has_bias and gb is not None
        gbias = alloc(R0, dtype=torch.float32, device=v.device) if want_b else None
        with torch.cuda.device(v.device):
            check(_lib.load().tfepb_wn_pack_backward(ptr(v), _ld(v), ptr(g), ptr(mask), 0 if mask is None else _ld(mask), R0, C,
                                                     ptr(row_perm), R, ptr(col_perm), ptr(gw), _ld(gw),
                                                     ptr(gb.contiguous()) if want_b else None, ptr(gv), C, ptr(gg), ptr(gbias),
                                                     stream_ptr(v)))
        return gv, gg, gbias, None, None, None, None


def wn_pack(v, g, bias, mask, row_perm=None, col_perm=None, onto=True):
    """See :class:`PackedWeightFunction`.  ``onto``: every row of ``v`` is referred to by ``row_perm`` (else the gradients of
    the other rows are zero-filled)."""
    require_cuda(v, g, bias, mask)
    if v.dtype != torch.float32:
        raise _lib.TfepB200Error('wn_pack takes float32 parameters')
    return PackedWeightFunction.apply(v.contiguous(), g.contiguous(), None if bias is None else bias.contiguous(),
                                      None if mask is None else _rows(mask), row_perm, col_perm, onto)


def made_forward(x, weights, biases, k_ranges=None, n_ranges=None):
    return MadeFunction.apply(x, len(weights), k_ranges, n_ranges, *weights, *biases)


# ---------------------------------------------------------------------------------------------
# transformers
# ---------------------------------------------------------------------------------------------

class ParamLayout:
    """Addressing of transformer parameters inside a (batch, *) parameter matrix (see tfep_b200.h)."""

    def __init__(self, offset=0, stride_p=1, stride_f=1, base=None):
        self.offset, self.stride_p, self.stride_f, self.base = offset, stride_p, stride_f, base

    @staticmethod
    def reference(n_features):
        """Parameter-major layout of the reference: column p * n_features + f."""
        return ParamLayout(0, n_features, 1, None)


def _tx_io(x, y, par, layout, n_features, inverse, cols, feat_ids, logdet, accumulate):
    return TxIo(dtype=dtype_code(x), batch=x.shape[0], n_features=n_features, inverse=int(inverse),
                x=x.data_ptr(), ldx=_ld(x), y=None if y is None else y.data_ptr(), ldy=0 if y is None else _ld(y),
                par=par.data_ptr(), ldp=_ld(par),
                par_offset=layout.offset, par_stride_p=layout.stride_p, par_stride_f=layout.stride_f,
                par_base=None if layout.base is None else layout.base.data_ptr(),
                cols=None if cols is None else cols.data_ptr(),
                feat_ids=None if feat_ids is None else feat_ids.data_ptr(),
                logdet=None if logdet is None else logdet.data_ptr(),
                accumulate_logdet=int(accumulate), reserved=0)


def _spline_cfg(spec, dtype, device, bins=None):
    dom = spec.domain_tensors(dtype, device)
    cfg = SplineCfg(n_bins=spec.n_bins_int, circular=int(spec.circular), identity_boundary_slopes=int(spec.identity_slopes),
                    learn_lower_bound=int(spec.learn_lower), learn_upper_bound=int(spec.learn_upper), reserved=0,
                    x0=dom[0].data_ptr(), xf=dom[1].data_ptr(), y0=dom[2].data_ptr(), yf=dom[3].data_ptr(),
                    min_bin_size=spec.min_bin_size, min_slope=spec.min_slope,
                    bins_out=None if bins is None else bins.data_ptr(), ldbins=0 if bins is None else bins.stride(0))
    return cfg, dom


def transformer_apply(kind, spec, x, par, layout, n_features, *, inverse=False, cols=None, feat_ids=None,
                      y=None, logdet=None, accumulate=False, bins=None):
    """Run one transformer kernel.  ``kind`` in {'affine','shift','spline','sos','moebius'}; ``spec`` carries its
    constants.  x / y are (batch, *) matrices addressed through ``cols``; returns (y, logdet)."""
    require_cuda(x, par)
    x, par = _rows(x), _rows(par)
    B = x.shape[0]
    if y is None:
        y = torch.empty_like(x)
    if logdet is None:
        logdet = torch.zeros(B, dtype=x.dtype, device=x.device) if accumulate else \
            torch.empty(B, dtype=x.dtype, device=x.device)
    if B == 0 or n_features == 0:
        if not accumulate:
            logdet.zero_()
        return y, logdet
    io = _tx_io(x, y, par, layout, n_features, inverse, cols, feat_ids, logdet, accumulate)
    lib, s = _lib.load(), stream_ptr(x)
    with torch.cuda.device(x.device):
        if kind == 'affine':
            check(lib.tfepb_affine(ctypes.byref(io), s))
        elif kind == 'shift':
            period, lower = spec.tables(x.dtype, x.device)
            check(lib.tfepb_shift(ctypes.byref(io), ptr(period), ptr(lower), s))
        elif kind == 'spline':
            cfg, keep = _spline_cfg(spec, x.dtype, x.device, bins)
            check(lib.tfepb_spline(ctypes.byref(io), ctypes.byref(cfg), s))
        elif kind == 'sos':
            check(lib.tfepb_sos(ctypes.byref(io), spec.n_polynomials, s))
        elif kind == 'moebius':
            check(lib.tfepb_moebius(ctypes.byref(io), spec.dimension, float(spec.max_radius), int(spec.unit_sphere), s))
        else:
            raise ValueError(kind)
    return y, logdet


def transformer_vjp(kind, spec, x, par, layout, n_features, grad_y, grad_logdet, *, cols=None, feat_ids=None,
                    grad_x=None, grad_par=None):
    """Vector-Jacobian product of the forward map; returns (grad_x, grad_par) (grad_par laid out like par)."""
    require_cuda(x, par, grad_y, grad_logdet)
    x, par, grad_y = _rows(x), _rows(par), _rows(grad_y)
    if grad_x is None:
        grad_x = torch.zeros_like(x)
    if grad_par is None:          # same leading dimension as par (which may be a view of a padded buffer)
        grad_par = torch.zeros((par.shape[0], _ld(par)), dtype=par.dtype, device=par.device)[:, :par.shape[1]]
    assert _ld(grad_par) == _ld(par)
    if x.shape[0] == 0 or n_features == 0:
        return grad_x, grad_par
    if grad_logdet is not None:
        grad_logdet = grad_logdet.contiguous()
    io = _tx_io(x, None, par, layout, n_features, False, cols, feat_ids, None, False)
    g = TxGrads(grad_y=grad_y.data_ptr(), ldgy=_ld(grad_y),
                grad_logdet=None if grad_logdet is None else grad_logdet.data_ptr(),
                grad_x=grad_x.data_ptr(), ldgx=_ld(grad_x), grad_par=grad_par.data_ptr())
    lib, s = _lib.load(), stream_ptr(x)
    with torch.cuda.device(x.device):
        if kind == 'affine':
            check(lib.tfepb_affine_backward(ctypes.byref(io), ctypes.byref(g), s))
        elif kind == 'shift':
            period, lower = spec.tables(x.dtype, x.device)
            check(lib.tfepb_shift_backward(ctypes.byref(io), ptr(period), ptr(lower), ctypes.byref(g), s))
        elif kind == 'spline':
            cfg, keep = _spline_cfg(spec, x.dtype, x.device)
            check(lib.tfepb_spline_backward(ctypes.byref(io), ctypes.byref(cfg), ctypes.byref(g), s))
        elif kind == 'sos':
            check(lib.tfepb_sos_backward(ctypes.byref(io), spec.n_polynomials, ctypes.byref(g), s))
        elif kind == 'moebius':
            check(lib.tfepb_moebius_backward(ctypes.byref(io), spec.dimension, float(spec.max_radius),
                                             int(spec.unit_sphere), ctypes.byref(g), s))
        else:
            raise ValueError(kind)
    return grad_x, grad_par


# ---------------------------------------------------------------------------------------------
# estimator / bootstrap
# ---------------------------------------------------------------------------------------------

def lse(w, scale, logw=None):
    """(max, sum exp(v - max)) of v = scale * w (+ logw) as a (2,) float64 device tensor."""
    require_cuda(w, logw)
    w = w.contiguous()
    if logw is not None:
        logw = logw.contiguous().to(w.dtype)
    lib = _lib.load()
    ws = torch.empty(lib.tfepb_lse_workspace_bytes() // 8, dtype=torch.float64, device=w.device)
    out = torch.empty(2, dtype=torch.float64, device=w.device)
    with torch.cuda.device(w.device):
        check(lib.tfepb_lse(dtype_code(w), ptr(w), ptr(logw), w.numel(), float(scale), ptr(ws), ptr(out),
                            stream_ptr(w)))
    return out


def fep_estimate(w, kT, log_norm, logw=None):
    """``-kT (logsumexp(-w / kT (+ logw)) - log_norm)`` as a 0-dim tensor of ``w``'s dtype, and the ``(max, sum)`` pair of the
    pass as a (2,) float64 tensor: one streaming kernel and one final reduction that also evaluates the estimate."""
    require_cuda(w, logw)
    w = w.contiguous()
    if logw is not None:
        logw = logw.contiguous().to(w.dtype)
    lib = _lib.load()
    ws = torch.empty(lib.tfepb_lse_workspace_bytes() // 8, dtype=torch.float64, device=w.device)
    out = torch.empty(2, dtype=torch.float64, device=w.device)
    res = torch.empty((), dtype=w.dtype, device=w.device)
    with torch.cuda.device(w.device):
        check(lib.tfepb_fep_estimate(dtype_code(w), ptr(w), ptr(logw), w.numel(), float(kT), float(log_norm), ptr(ws), ptr(out),
                                     ptr(res), stream_ptr(w)))
    return res, out


def exp_table(w, scale, max_dev):
    require_cuda(w, max_dev)
    w = w.contiguous()
    e = torch.empty(w.numel(), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        check(_lib.load().tfepb_exp_table(dtype_code(w), ptr(w), w.numel(), float(scale), ptr(max_dev), ptr(e),
                                          stream_ptr(w)))
    return e


#: resamples one tfepb_bootstrap_sums launch takes (they map to grid.y)
MAX_RESAMPLES_PER_CALL = 65535


def bootstrap_sums(e, max_idx, n_resamples, sample_size, idx=None, philox_seed=0, philox_offset=0, shard_lo=0,
                   sample_sizes=None):
    """Per-resample sums of the exp table ``e`` (the shard starting at global index ``shard_lo``).  ``sample_sizes``:
    optional int64 device tensor with the number of draws of every resample (Philox stream; at most ``sample_size``)."""
    require_cuda(e, idx, sample_sizes)
    out = torch.empty(n_resamples, dtype=torch.float64, device=e.device)
    with torch.cuda.device(e.device):
        check(_lib.load().tfepb_bootstrap_sums(ptr(e), e.numel(), int(shard_lo), int(max_idx), ptr(idx),
                                               0 if idx is None else idx.stride(0), int(n_resamples), int(sample_size),
                                               int(philox_seed), int(philox_offset), ptr(sample_sizes), ptr(out),
                                               stream_ptr(e)))
    return out


def bayesian_bootstrap_sums(e, n_resamples, philox_seed=0, philox_offset=0):
    """Per-resample ``(sum_i e_i g_ri, sum_i g_ri)`` with Exp(1) variates g (Dirichlet(1..1) weights, unnormalised)."""
    require_cuda(e)
    s = torch.empty(n_resamples, dtype=torch.float64, device=e.device)
    g = torch.empty(n_resamples, dtype=torch.float64, device=e.device)
    with torch.cuda.device(e.device):
        check(_lib.load().tfepb_bayesian_bootstrap_sums(ptr(e), e.numel(), int(n_resamples), int(philox_seed),
                                                        int(philox_offset), ptr(s), ptr(g), stream_ptr(e)))
    return s, g


def mt19937_seed(seed):
    st = torch.empty(625, dtype=torch.int32)
    check(_lib.load().tfepb_mt19937_seed(ctypes.c_uint32(seed & 0xffffffff), ptr(st)))
    return st


#: requests of at least this many draws are generated by all SMs (jump-ahead); below it one CTA walks the stream
MT_PARALLEL_MIN = 1 << 22
MT_DRAWS_PER_STREAM_MIN = 64 * 624
_MT_WORKSPACES = {}


def mt19937_jump_polynomial(steps):
    """t^steps mod the characteristic polynomial of MT19937: 624 uint32 words (host computation, no GPU needed)."""
    import numpy as np
    out = np.zeros(624, dtype=np.uint32)
    check(_lib.load().tfepb_mt19937_jump_polynomial(ctypes.c_uint64(int(steps)), out.ctypes.data_as(ctypes.c_void_p)))
    return out


def mt19937_indices(state_dev, count, max_idx, out=None, n_streams=None):
    """Next ``count`` draws ``u32 % max_idx`` of the MT19937 stream held in ``state_dev`` (625 int32 on the
    device, updated in place).  Large requests are split into sub-streams by jump-ahead and generated by all SMs
    (tfepb_mt19937_indices_parallel), bit-identical to the sequential walk; ``n_streams`` forces the split."""
    require_cuda(state_dev)
    idx = out if out is not None else torch.empty(count, dtype=torch.int32, device=state_dev.device)
    lib = _lib.load()
    if n_streams is None:
        n_streams = 1
        if count >= MT_PARALLEL_MIN:
            sms = torch.cuda.get_device_properties(state_dev.device).multi_processor_count
            n_streams = max(1, min(4 * sms, int(count) // MT_DRAWS_PER_STREAM_MIN))
    with torch.cuda.device(state_dev.device):
        if n_streams <= 1:
            check(lib.tfepb_mt19937_indices(ptr(state_dev), int(count), int(max_idx), ptr(idx), stream_ptr(state_dev)))
        else:
            key = (str(state_dev.device), torch.cuda.current_stream(state_dev.device).cuda_stream)
            ws = _MT_WORKSPACES.get(key)
            need = lib.tfepb_mt19937_parallel_workspace_bytes(int(n_streams))
            if ws is None or ws.numel() < need:
                ws = _MT_WORKSPACES[key] = torch.empty(need, dtype=torch.uint8, device=state_dev.device)
            check(lib.tfepb_mt19937_indices_parallel(ptr(state_dev), int(count), int(max_idx), ptr(idx), int(n_streams),
                                                     ptr(ws), stream_ptr(state_dev)))
    return idx


# ---------------------------------------------------------------------------------------------
# tensor-core GEMM (bf16 operand images, fp32 accumulation)
# ---------------------------------------------------------------------------------------------

def tc_pack(src, block_rows, transpose=False, out=None, n_split=1):
    """bf16 operand image of a 2-D fp32 CUDA tensor (rows x k, or its transpose if ``transpose``); see tfepb_tc_pack.
    ``n_split`` = 2 / 3: the consecutive images of the split-precision terms (tfepb_tc_pack_split)."""
    require_cuda(src)
    if src.dtype != torch.float32:
        raise _lib.TfepB200Error('tc_pack takes float32 tensors')
    src = _rows(src)
    rows, k = (src.shape[1], src.shape[0]) if transpose else (src.shape[0], src.shape[1])
    lib = _lib.load()
    nbytes = lib.tfepb_tc_image_bytes(rows, k, block_rows) * n_split
    img = out if out is not None else torch.empty(nbytes, dtype=torch.uint8, device=src.device)
    assert img.numel() >= nbytes
    with torch.cuda.device(src.device):
        if n_split == 1:
            check(lib.tfepb_tc_pack(ptr(src), _ld(src), rows, k, block_rows, int(transpose), ptr(img), stream_ptr(src)))
        else:
            check(lib.tfepb_tc_pack_split(ptr(src), _ld(src), rows, k, block_rows, int(transpose), int(n_split), ptr(img),
                                          stream_ptr(src)))
    return img


def tc_pack_dual(src, t_block_rows=None, column_sums=False, image=True, block_rows=128):
    """One pass over a 2-D fp32 CUDA tensor: (its image with ``block_rows`` = 128 / 256, the image of its transpose with
    ``t_block_rows`` = 128 / 256, its column sums); see tfepb_tc_pack_dual.  Parts not asked for are None."""
    require_cuda(src)
    if src.dtype != torch.float32:
        raise _lib.TfepB200Error('tc_pack_dual takes float32 tensors')
    src = _rows(src)
    rows, cols = src.shape
    lib = _lib.load()
    img = torch.empty(lib.tfepb_tc_image_bytes(rows, cols, int(block_rows)), dtype=torch.uint8, device=src.device) if image else None
    img_t = None
    if t_block_rows:
        img_t = torch.empty(lib.tfepb_tc_image_bytes(cols, rows, int(t_block_rows)), dtype=torch.uint8, device=src.device)
    sums = torch.zeros(cols, dtype=torch.float32, device=src.device) if column_sums else None
    with torch.cuda.device(src.device):
        check(lib.tfepb_tc_pack_dual(ptr(src), _ld(src), rows, cols, ptr(img), int(block_rows), ptr(img_t), int(t_block_rows or 0),
                                     ptr(sums), stream_ptr(src)))
    return img, img_t, sums


#: launch the tensor-core products as clusters of two CTAs that share the B operand (tfepb_tc_gemm_args.cluster): a third less
#: L2 -> SM traffic, same speed on an otherwise idle B200; off by default
TC_CLUSTER = False

TCTX_KINDS = {'affine': 1, 'sos': 2, 'moebius': 3, 'spline': 4}
TCTX_UNITS_PER_CHUNK = {'affine': 8, 'sos': 3, 'moebius': 5}       # units per 16-column chunk (tfepb_tc_tx)
TCTX_COLUMNS_PER_UNIT = {'affine': 2, 'sos': 5, 'moebius': 3}


class TcTx:
    """The transformer fused into the epilogue of an output-layer product (tfepb_tc_tx).  ``cols``: int32 device tensor,
    the x / y columns of every unit in chunk order.  Forward: ``y`` / ``logdet`` (accumulated into); backward:
    ``grad_y``, ``grad_logdet`` (or None), ``grad_x``."""

    def __init__(self, kind, cols, x, *, y=None, logdet=None, grad_y=None, grad_logdet=None, grad_x=None,
                 max_radius=0.0, unit_sphere=0, spline=None):
        self.kind, self.cols, self.x = kind, cols, x
        #: kind 'spline': float32 device tensor (n_units, 8) = x0, xf, y0, yf, min_bin_size, min_slope, flags (int bits), 0
        self.spline = spline
        self.y, self.logdet, self.grad_y, self.grad_logdet, self.grad_x = y, logdet, grad_y, grad_logdet, grad_x
        self.max_radius, self.unit_sphere = float(max_radius), int(unit_sphere)
        self.backward = grad_x is not None

    def struct(self):
        xcols = 3 if self.kind == 'moebius' else 1
        opt = lambda t: None if t is None else t.data_ptr()
        ld = lambda t: 0 if t is None else _ld(t)
        st = _lib.TcTx(kind=TCTX_KINDS[self.kind], backward=int(self.backward), n_units=self.cols.numel() // xcols,
                       unit_sphere=self.unit_sphere, max_radius=self.max_radius, cols=self.cols.data_ptr(),
                       x=self.x.data_ptr(), ldx=_ld(self.x), y=opt(self.y), ldy=ld(self.y), logdet=opt(self.logdet),
                       grad_y=opt(self.grad_y), ldgy=ld(self.grad_y), grad_logdet=opt(self.grad_logdet),
                       grad_x=opt(self.grad_x), ldgx=ld(self.grad_x))
        if self.spline is not None:
            st.spline_table = self.spline.data_ptr()
        return st


def tc_gemm(a_img, b_img, m, n, k, *, c=None, bias=None, activation=ACT_NONE, aux=None, out_image=False,
            k_block_ranges=None, row_ranges=None, split_k=1, error_flag=None, out_image_t=None, column_sums=False, n_split=1,
            tx=None, aux_image=None, accumulate=False, mn_major=False, cluster=None, tile_list=None):
    """C = act(A B^T + bias) [* ELU'(aux)] from operand images; returns (c, out_img).  ``c``: True to allocate, a
    tensor to write into (zero-filled by the caller when split_k > 1), None for no fp32 output.
    ``out_image_t`` = 128 / 256: also the image of the transposed result with that block_rows; ``column_sums``: also
    the sums over the m rows; with either, returns (c, out_img, out_img_t, column_sums).  ``tx``: a :class:`TcTx`, the
    transformer applied by the epilogue (the columns are then parameter chunks, see tfepb_tc_tx).  ``aux_image``: the
    ELU' operand as the bf16 image a forward product wrote (instead of the fp32 ``aux``).  ``accumulate``: ``c`` (a tensor)
    += the result.  ``mn_major``: ``a_img`` / ``b_img`` are the ROW images of a (k x m) and a (k x n) matrix, C = A^T B
    (the weight gradient straight from the images of grad_y and x).  ``cluster``: see :data:`TC_CLUSTER`."""
    lib = _lib.load()
    dev = a_img.device
    if c is True:
        pad4 = (n + 3) // 4 * 4
        c = (torch.zeros if split_k > 1 or mn_major else torch.empty)((m, pad4), dtype=torch.float32, device=dev)[:, :n]
    img = None
    if out_image:
        img = torch.empty(lib.tfepb_tc_image_bytes(m, n, 128) * n_split, dtype=torch.uint8, device=dev)
    img_t = None
    if out_image_t:
        img_t = torch.empty(lib.tfepb_tc_image_bytes(n, m, int(out_image_t)), dtype=torch.uint8, device=dev)
    sums = torch.zeros(n, dtype=torch.float32, device=dev) if column_sums else None
    a = _lib.TcGemmArgs(a_image=a_img.data_ptr(), b_image=b_img.data_ptr(), m=m, n=n, k=k, activation=activation,
                        c=None if c is None else c.data_ptr(), ldc=0 if c is None else _ld(c),
                        bias=None if bias is None else bias.data_ptr(),
                        aux=None if aux is None else aux.data_ptr(), ldaux=0 if aux is None else _ld(aux),
                        out_image=None if img is None else img.data_ptr(),
                        k_block_ranges=None if k_block_ranges is None else k_block_ranges.data_ptr(),
                        split_k=int(split_k), out_image_t_rows=int(out_image_t or 0),
                        error_flag=None if error_flag is None else error_flag.data_ptr(),
                        row_ranges=None if row_ranges is None else row_ranges.data_ptr(),
                        out_image_t=None if img_t is None else img_t.data_ptr(),
                        column_sums=None if sums is None else sums.data_ptr(), n_split=int(n_split),
                        c_accumulate=int(bool(accumulate)))
    if tx is not None:
        txs = tx.struct()                 # kept alive until the launch returns
        a.tx = ctypes.pointer(txs)
    if aux_image is not None:
        a.aux_image = aux_image.data_ptr()
    a.mn_major = int(bool(mn_major))
    a.cluster = int(TC_CLUSTER if cluster is None else bool(cluster))
    if tile_list is not None:
        a.tile_list, a.n_tile_list = tile_list.data_ptr(), tile_list.shape[0]
    with torch.cuda.device(dev):
        check(lib.tfepb_tc_gemm(ctypes.byref(a), stream_ptr(a_img)))
    if out_image_t or column_sums:
        return c, img, img_t, sums
    return c, img


class MadeFunctionTC(torch.autograd.Function):
    """All layers of a MADE conditioner on the tensor cores (tfepb_tc_gemm), as one autograd node.

    Same contract as :class:`MadeFunction` (packed effective weights in, dense gradients out) with bf16 operands and
    fp32 accumulation: the forward pass chains the bf16 row image of every hidden activation into the next product; the
    backward pass runs dX = (dY W) * ELU'(h) on the same kernel (ELU' read from the image of h) and dW = dY^T X on the
    MN-major weight-gradient kernel, which reduces over the rows of the SAME images of dY and X (split over the batch,
    fp32 atomics).  ``kb_fwd[l]`` / ``kb_bwd[l]`` are the per-256-column-tile ranges of non-zero 64-wide k-blocks of the
    staircase masks (or None).
    """

    @staticmethod
    def forward(ctx, x, n_layers, kb_fwd, kb_bwd, rr_w, *wb):
        ws, bs = wb[:n_layers], wb[n_layers:]
        B = x.shape[0]
        # Every activation exists as ONE bf16 row image: operand of the next product, ELU' operand of the backward pass and
        # (read MN-major) operand of the weight gradient.
        img = tc_pack(x, 128)
        acts = [x]
        imgs = [img]
        h = x
        keep = any(ctx.needs_input_grad)
        wts = []
        for l in range(n_layers):
            last = l == n_layers - 1
            N, K = ws[l].shape
            wimg, wt = _weight_images(ws[l], keep and (l > 0 or ctx.needs_input_grad[0]))
            wts.append(wt)
            h, img = tc_gemm(img, wimg, B, N, K, c=True if last else None, bias=bs[l],
                             activation=ACT_NONE if last else ACT_ELU, out_image=not last,
                             k_block_ranges=None if kb_fwd is None else kb_fwd[l])
            if not last:
                acts.append(img)
                imgs.append(img)
        ctx.save_for_backward(*acts, *ws, imgs[0])
        ctx.wts = wts
        ctx.n_layers = n_layers
        ctx.kb_bwd = kb_bwd
        ctx.rr_w = rr_w
        return h

    @staticmethod
    def backward(ctx, grad_out):
        L = ctx.n_layers
        saved = ctx.saved_tensors
        acts, ws, x_img = saved[:L], saved[L:2 * L], saved[2 * L]
        g = _rows(grad_out.contiguous())
        # image of g, image of g^T (A operand of the weight gradient) and the bias gradient: one pass over grad_out for the
        # top layer, written by the epilogue of the backward-input product for the layers below
        top = ctx.needs_input_grad[5 + L - 1] or ctx.needs_input_grad[5 + 2 * L - 1]
        gimg, _, gb = tc_pack_dual(g, None, column_sums=top)
        need_w = [ctx.needs_input_grad[5 + l] or ctx.needs_input_grad[5 + L + l] for l in range(L)]
        gx, gws, gbs = _made_tc_backward_layers(g.shape[0], [x_img] + list(acts[1:]), ws, ctx.kb_bwd, ctx.rr_w, need_w,
                                                ctx.needs_input_grad[0], gimg, gb, wts=ctx.wts)
        ctx.wts = None
        return (gx, None, None, None, None, *gws, *gbs)


class WgTiles:
    """The (128-row, 256-column) tiles of a weight gradient that its mask leaves non-zero: int32 device tensor (n, 2)."""

    def __init__(self, tiles):
        self.tiles, self.n = tiles, int(tiles.shape[0])


def _weight_gradient_split(n, k, batch, n_sm, block=128, tiles=None):
    """Split of the batch reduction of dW[n x k]: the kernel runs n_sm // split CTAs per slice, each walking
    ceil(tiles / (n_sm // split)) tiles over batch / split samples -- pick the split with the shortest critical path (a split
    that leaves a partial last round of tiles wastes up to half the machine: 39 tiles on 18 CTAs = 3 rounds for 2.2)."""
    if tiles is None:
        tiles = ((n + 127) // 128) * ((k + 255) // 256)
    tiles = max(tiles, 1)
    k_blocks = (batch + block - 1) // block          # reduction blocks (128 image rows per ring stage)
    best, best_cost = 1, None
    for split in range(1, min(k_blocks, n_sm) + 1):
        ctas = n_sm // split
        rounds = (tiles + ctas - 1) // ctas
        chunk = (k_blocks + split - 1) // split
        cost = rounds * (chunk + 4)                  # + per-tile epilogue (atomics) in units of reduction blocks
        if best_cost is None or cost < best_cost:
            best, best_cost = split, cost
    return best


def _weight_images(w, both):
    """(image of w, image of w^T or None): the B operands of the forward and of the backward-input product of a layer, in ONE
    pass over the packed weight when the backward pass will run."""
    if both:
        img, img_t, _ = tc_pack_dual(w, 256, block_rows=256)
        return img, img_t
    return tc_pack(w, 256), None


def _made_tc_backward_layers(B, imgs, ws, kb_bwd, rr_w, need_w, need_x, gimg, gb, gx_into=None, wts=None):
    """Backward pass of the layers of a MADE on the tensor cores.  ``imgs``: the bf16 row images of the input of every layer (x,
    then the hidden activations); ``gimg`` / ``gb``: the row image of the cotangent of the output layer's result and its column
    sums.  Every matrix exists as ONE image: the backward-input product reads ``gimg`` K-major, the weight gradient
    dW = dY^T X reads ``gimg`` and ``imgs[l]`` MN-major (reduction over their rows), ELU' reads ``imgs[l]``.  Returns
    (grad_x or None, grad_ws, grad_bs).  ``B``: the batch size."""
    L = len(ws)
    dev = gimg.device
    gws, gbs = [None] * L, [None] * L
    gx = None
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    for l in range(L - 1, -1, -1):
        N, K = ws[l].shape
        if need_w[l]:
            # dW[N x K] = dY^T X: reduction over the batch, split so that the grid fills the machine
            rr = None if rr_w is None else rr_w[l]
            if isinstance(rr, WgTiles):
                # only the tiles the mask leaves non-zero, as a list: every CTA carries the same number of real tiles
                split = _weight_gradient_split(N, K, B, n_sm, tiles=rr.n)
                gws[l], _ = tc_gemm(gimg, imgs[l], N, K, B, c=True, split_k=split, mn_major=True, tile_list=rr.tiles)
            else:
                split = _weight_gradient_split(N, K, B, n_sm)
                gws[l], _ = tc_gemm(gimg, imgs[l], N, K, B, c=True, split_k=split, row_ranges=rr, mn_major=True)
            gbs[l] = gb
        gb = None
        if l > 0 or need_x:
            # rows = inputs of the layer, k = its outputs (packed by the forward pass together with the forward operand)
            wt = wts[l] if wts is not None and wts[l] is not None else tc_pack(ws[l], 256, transpose=True)
            kb = None if kb_bwd is None else kb_bwd[l]
            if l == 0:
                # ``gx_into``: the cotangent of x is ADDED to that tensor (the direct term of a fused transformer)
                gx, _ = tc_gemm(gimg, wt, B, K, N, c=True if gx_into is None else gx_into,
                                accumulate=gx_into is not None, k_block_ranges=kb)
            else:
                # the cotangent of a hidden activation exists as its row image (and column sums) only
                out = tc_gemm(gimg, wt, B, K, N, aux_image=imgs[l], out_image=True, column_sums=need_w[l - 1], k_block_ranges=kb)
                gimg, gb = out[1], (out[3] if need_w[l - 1] else None)
    return gx, gws, gbs


class MadeTxFunctionTC(torch.autograd.Function):
    """(y, log_det) = transformer(x, MADE(x)) with the transformer fused into the output-layer product (tfepb_tc_tx): the
    conditioner output is never written.  ``ws[-1]`` / ``bs[-1]`` are the output layer in the padded chunk layout of the
    transformer kind (tfep_b200/_txfused.py), the range tables of that layer refer to the same layout.  Backward: the
    output-layer product is recomputed from the saved image of the last hidden activation and its epilogue emits the
    parameter cotangents as the operand image of the products below (no fp32 cotangent matrix either), then the layers
    follow as in :class:`MadeFunctionTC`."""

    @staticmethod
    def forward(ctx, x, xc, n_layers, kb_fwd, kb_bwd, rr_w, spec, *wb):
        # ``xc``: the conditioner input where it differs from x (an embedding layer in front of the MADE), else None
        ws, bs = wb[:n_layers], wb[n_layers:]
        L = n_layers
        B = x.shape[0]
        keep = any(ctx.needs_input_grad)
        need_cx = ctx.needs_input_grad[0] if xc is None else ctx.needs_input_grad[1]
        img = tc_pack(x if xc is None else xc, 128)
        imgs, wts = [img], []
        for l in range(L - 1):
            N, K = ws[l].shape
            wimg, wt = _weight_images(ws[l], keep and (l > 0 or need_cx))
            wts.append(wt)
            _, img = tc_gemm(img, wimg, B, N, K, bias=bs[l], activation=ACT_ELU, out_image=True,
                             k_block_ranges=None if kb_fwd is None else kb_fwd[l])
            imgs.append(img)                 # bf16 row image: next operand, ELU' operand and weight-gradient operand
        N, K = ws[-1].shape
        # (conditioning features are not mapped by any unit: they pass through)
        y = x.clone() if spec.get('passthrough') else torch.empty_like(x)
        logdet = torch.zeros(B, dtype=torch.float32, device=x.device)
        wimg, wt = _weight_images(ws[-1], keep)
        wts.append(wt)
        tc_gemm(img, wimg, B, N, K, bias=bs[-1], k_block_ranges=None if kb_fwd is None else kb_fwd[-1],
                tx=TcTx(spec['kind'], spec['cols'], x, y=y, logdet=logdet, max_radius=spec['max_radius'],
                        unit_sphere=spec['unit_sphere'], spline=spec.get('spline')))
        ctx.save_for_backward(x, *(imgs if keep else []), *ws, bs[-1], *([wimg] if keep else []))
        ctx.wts = wts if keep else None
        ctx.meta = (L, kb_fwd, kb_bwd, rr_w, spec, xc is not None)
        if spec['kind'] == 'sos':
            ctx.mark_non_differentiable(logdet)                  # the reference's SOS log-det carries no gradient (sos.py:233)
        return y, logdet

    @staticmethod
    def backward(ctx, grad_y, grad_ld):
        L, kb_fwd, kb_bwd, rr_w, spec, embedded = ctx.meta
        saved = list(ctx.saved_tensors)
        x, imgs, ws, b_last, w_last_img = saved[0], saved[1:1 + L], saved[1 + L:1 + 2 * L], saved[1 + 2 * L], saved[2 + 2 * L]
        B = x.shape[0]
        grad_y = torch.zeros_like(x) if grad_y is None else _rows(grad_y.contiguous())
        if grad_ld is not None:
            grad_ld = grad_ld.contiguous()
        need_w = [ctx.needs_input_grad[7 + l] or ctx.needs_input_grad[7 + L + l] for l in range(L)]
        gx = grad_y.clone() if spec.get('passthrough') else torch.empty_like(x)
        N, K = ws[-1].shape
        _, gimg, _, gb = tc_gemm(imgs[-1], w_last_img, B, N, K, bias=b_last, out_image=True, column_sums=True,
                                 k_block_ranges=None if kb_fwd is None else kb_fwd[-1],
                                 tx=TcTx(spec['kind'], spec['cols'], x, grad_y=grad_y, grad_logdet=grad_ld, grad_x=gx,
                                         max_radius=spec['max_radius'], unit_sphere=spec['unit_sphere'],
                                         spline=spec.get('spline')))
        if embedded:
            # the cotangent of the embedded conditioner input goes back through the embedding's own graph
            gxc, gws, gbs = _made_tc_backward_layers(B, imgs, ws, kb_bwd, rr_w, need_w, ctx.needs_input_grad[1], gimg, gb, wts=ctx.wts)
        else:
            gxc = None
            _, gws, gbs = _made_tc_backward_layers(B, imgs, ws, kb_bwd, rr_w, need_w, ctx.needs_input_grad[0], gimg, gb, gx_into=gx,
                                                   wts=ctx.wts)
        ctx.wts = None
        return (gx if ctx.needs_input_grad[0] else None, gxc, None, None, None, None, None, *gws, *gbs)


def made_tx_forward_tc(x, weights, biases, kb_fwd, kb_bwd, rr_w, spec, xc=None):
    """See :class:`MadeTxFunctionTC`; ``spec``: dict(kind, cols, max_radius, unit_sphere[, spline]); ``xc``: the conditioner
    input if an embedding layer sits in front of the MADE."""
    return MadeTxFunctionTC.apply(x, xc, len(weights), kb_fwd, kb_bwd, rr_w, spec, *weights, *biases)


def made_forward_tc(x, weights, biases, kb_fwd=None, kb_bwd=None, rr_w=None, n_split=1, weight_images=None):
    """All layers of a MADE on the tensor cores.  ``n_split`` = 2 / 3: split-precision operands (3 / 6 bf16 products per
    reduction step, fp32-class accuracy); an inference path -- under autograd use n_split = 1 or the exact kernels.
    ``weight_images``: optional cached ``tc_pack(w, 256, n_split=n_split)`` of every layer."""
    if x.shape[0] == 0:                       # nothing to launch; zero-sized result tied to the inputs for autograd
        return x.new_zeros((0, weights[-1].shape[0])) + 0.0 * sum(w.sum() + b.sum() for w, b in zip(weights, biases))
    if n_split == 1:
        return MadeFunctionTC.apply(x, len(weights), kb_fwd, kb_bwd, rr_w, *weights, *biases)
    if torch.is_grad_enabled() and (x.requires_grad or any(t.requires_grad for t in (*weights, *biases))):
        raise NotImplementedError("tfep_b200: the split-precision tensor-core conditioner (precision='bf16x3' / 'bf16x6') "
                                  "is an inference path; use torch.no_grad(), or precision='fp32' / 'bf16' for training")
    B = x.shape[0]
    img = tc_pack(x.contiguous(), 128, n_split=n_split)
    h = None
    for l, (w, b) in enumerate(zip(weights, biases)):
        last = l == len(weights) - 1
        N, K = w.shape
        wimg = weight_images[l] if weight_images is not None else tc_pack(w, 256, n_split=n_split)
        h, img = tc_gemm(img, wimg, B, N, K, c=True if last else None, bias=b, activation=ACT_NONE if last else ACT_ELU,
                         out_image=not last, k_block_ranges=None if kb_fwd is None else kb_fwd[l], n_split=n_split)
    return h
