"""General tensor-core GEMM (tfepb_tc_pack / tfepb_tc_gemm) against a double-precision product of the same
bf16-rounded operands: forward (bias, ELU, image of the result chained into a second product), backward input
(ELU' multiplier), weight gradient (transposed images, split-K atomics), staircase k-block ranges, ragged sizes."""

import pytest
import torch

from tfep_b200 import _ops

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _bf(t):
    return t.to(torch.bfloat16).double()


def _rand(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g).to(DEV)


@pytest.mark.parametrize('m,n,k', [(128, 256, 64), (1000, 670, 300), (257, 1500, 670), (5, 7, 9)])
def test_forward_bias_elu_and_chained_image(m, n, k):
    x, w, b = _rand((m, k), 1), _rand((n, k), 2) / k ** 0.5, _rand((n,), 3)
    ai, bi = _ops.tc_pack(x, 128), _ops.tc_pack(w, 256)
    c, img = _ops.tc_gemm(ai, bi, m, n, k, c=True, bias=b, activation=_ops.ACT_ELU, out_image=True)
    ref = torch.nn.functional.elu(_bf(x) @ _bf(w).T + b.double())
    assert c.shape == (m, n)
    assert float((c.double() - ref).abs().max()) < 2e-4 * (1 + float(ref.abs().max()))
    # the image of the result is the A operand of a second product
    n2 = 130
    w2 = _rand((n2, n), 4) / n ** 0.5
    c2, _ = _ops.tc_gemm(img, _ops.tc_pack(w2, 256), m, n2, n, c=True)
    ref2 = _bf(c) @ _bf(w2).T
    assert float((c2.double() - ref2).abs().max()) < 2e-4 * (1 + float(ref2.abs().max()))


def test_backward_products_and_k_ranges():
    m, n, k = 700, 328, 1650                          # dX = dY W: reduction over the 1650 outputs
    gy, w, h = _rand((m, k), 5), _rand((k, n), 6) / k ** 0.5, _rand((m, n), 7)
    gx, _ = _ops.tc_gemm(_ops.tc_pack(gy, 128), _ops.tc_pack(w, 256, transpose=True), m, n, k, c=True, aux=h)
    ref = (_bf(gy) @ _bf(w)) * torch.where(h > 0, torch.ones_like(h), h + 1).double()
    assert float((gx.double() - ref).abs().max()) < 3e-4 * (1 + float(ref.abs().max()))
    # weight gradient dW = dY^T X over a batch of 5000, split-K with atomics into a zero-filled C
    B, no, ni = 5000, 300, 200
    gy, x = _rand((B, no), 8), _rand((B, ni), 9)
    gw, _ = _ops.tc_gemm(_ops.tc_pack(gy, 128, transpose=True), _ops.tc_pack(x, 256, transpose=True), no, ni, B, c=True, split_k=16)
    ref = _bf(gy).T @ _bf(x)
    assert float((gw.double() - ref).abs().max()) < 3e-4 * (1 + float(ref.abs().max()))
    # staircase: only the k-blocks inside the given range are multiplied
    m, n, k = 300, 512, 640
    x, w = _rand((m, k), 10), _rand((n, k), 11)
    ranges = torch.tensor([[0, 3], [2, 10]], dtype=torch.int32, device=DEV)      # n-tile 0: k < 192; n-tile 1: k >= 128
    c, _ = _ops.tc_gemm(_ops.tc_pack(x, 128), _ops.tc_pack(w, 256), m, n, k, c=True, k_block_ranges=ranges)
    wm = w.clone()
    wm[:256, 192:] = 0
    wm[256:, :128] = 0
    ref = _bf(x) @ _bf(wm).T
    assert float((c.double() - ref).abs().max()) < 3e-4 * (1 + float(ref.abs().max()))


@pytest.mark.parametrize('m,n,k,rows', [(1000, 670, 300, 256), (257, 300, 670, 128), (5, 7, 9, 256), (4096, 1500, 128, 128)])
def test_transposed_image_and_column_sums_from_the_epilogue(m, n, k, rows):
    """The epilogue can also emit the image of the TRANSPOSED result (the operand of the weight-gradient product) and
    the column sums (the bias gradient): bit-identical to packing the fp32 result, sums equal to the row sum."""
    x, w, b, h = _rand((m, k), 21), _rand((n, k), 22) / k ** 0.5, _rand((n,), 23), _rand((m, n), 24)
    ai, bi = _ops.tc_pack(x, 128), _ops.tc_pack(w, 256)
    for kw in (dict(bias=b, activation=_ops.ACT_ELU), dict(aux=h)):
        c, img, img_t, sums = _ops.tc_gemm(ai, bi, m, n, k, c=True, out_image=True, out_image_t=rows, column_sums=True, **kw)
        assert torch.equal(img_t, _ops.tc_pack(c, rows, transpose=True))
        ref = c.double().sum(dim=0)
        assert float((sums.double() - ref).abs().max()) < 1e-5 * (1 + float(c.abs().sum(dim=0).max()))
        # without the fp32 result
        _, _, img_t2, sums2 = _ops.tc_gemm(ai, bi, m, n, k, out_image_t=rows, column_sums=True, **kw)
        assert torch.equal(img_t2, img_t)
        assert float((sums2 - sums).abs().max()) < 1e-5 * (1 + float(c.abs().sum(dim=0).max()))


@pytest.mark.parametrize('rows,cols,t_rows', [(1000, 670, 128), (257, 300, 256), (5, 7, 128), (4096, 1500, 256), (64, 64, 256)])
def test_dual_pack_equals_the_separate_packs(rows, cols, t_rows):
    src = _rand((rows, cols), 31)
    img, img_t, sums = _ops.tc_pack_dual(src, t_rows, column_sums=True)
    assert torch.equal(img, _ops.tc_pack(src, 128))
    assert torch.equal(img_t, _ops.tc_pack(src, t_rows, transpose=True))
    assert float((sums.double() - src.double().sum(dim=0)).abs().max()) < 1e-5 * (1 + float(src.abs().sum(dim=0).max()))
    img2, none_t, none_s = _ops.tc_pack_dual(src)
    assert torch.equal(img2, img) and none_t is None and none_s is None
    # padded leading dimension (a column slice of a wider tensor)
    wide = _rand((rows, cols + 5), 32)
    i3, t3, _ = _ops.tc_pack_dual(wide[:, :cols], t_rows)
    assert torch.equal(i3, _ops.tc_pack(wide[:, :cols].contiguous(), 128))
    assert torch.equal(t3, _ops.tc_pack(wide[:, :cols].contiguous(), t_rows, transpose=True))


@pytest.mark.parametrize('n_split,tol', [(2, 4e-5), (3, 1.5e-5)])
@pytest.mark.parametrize('m,n,k', [(128, 256, 64), (1000, 670, 300), (257, 1650, 330), (5, 7, 9), (300, 512, 96)])
def test_split_precision_products(m, n, k, n_split, tol):
    """x = x_0 + x_1 (+ x_2) in bf16 terms, products A_i B_j with i + j < n_split (3 or 6 MMAs per k-step): against the
    double-precision product of the UNROUNDED fp32 operands; the image of the result (split the same way by the
    epilogue) chains into a second product; staircase k-ranges apply per 64-wide block as in the plain mode."""
    x, w, b = _rand((m, k), 41), _rand((n, k), 42) / k ** 0.5, _rand((n,), 43)
    ai, bi = _ops.tc_pack(x, 128, n_split=n_split), _ops.tc_pack(w, 256, n_split=n_split)
    c, img = _ops.tc_gemm(ai, bi, m, n, k, c=True, bias=b, activation=_ops.ACT_ELU, out_image=True, n_split=n_split)
    ref = torch.nn.functional.elu(x.double() @ w.double().T + b.double())
    scale = 1 + float(ref.abs().max())
    assert float((c.double() - ref).abs().max()) < tol * scale
    n2 = 130
    w2 = _rand((n2, n), 44) / n ** 0.5
    c2, _ = _ops.tc_gemm(img, _ops.tc_pack(w2, 256, n_split=n_split), m, n2, n, c=True, n_split=n_split)
    ref2 = c.double() @ w2.double().T
    assert float((c2.double() - ref2).abs().max()) < tol * (1 + float(ref2.abs().max()))
    # the first term of a split image IS the plain bf16 image
    nbytes = _ops._lib.load().tfepb_tc_image_bytes(m, k, 128)
    assert torch.equal(ai[:nbytes], _ops.tc_pack(x, 128))
    if n >= 512 and k >= 96:
        tiles = (n + 255) // 256                        # tile 0 sees k-block 0 only, the others k-block 1 only
        ranges = torch.tensor([[0, 1]] + [[1, 2]] * (tiles - 1), dtype=torch.int32, device=DEV)
        cr, _ = _ops.tc_gemm(ai, bi, m, n, k, c=True, k_block_ranges=ranges, n_split=n_split)
        wm = w.clone()
        wm[:256, 64:] = 0
        wm[256:, :64] = 0
        wm[256:, 128:] = 0
        refr = x.double() @ wm.double().T
        assert float((cr.double() - refr).abs().max()) < tol * (1 + float(refr.abs().max()))


@pytest.mark.parametrize('batch,n_out,n_in,split', [(5000, 300, 200, 16), (130, 1600, 670, 3), (64, 7, 9, 1), (1000, 670, 300, 5),
                                                    (5, 33, 300, 4), (129, 128, 256, 2)])
def test_weight_gradient_from_row_images_mn_major(batch, n_out, n_in, split):
    """dW = dY^T X read MN-major out of the SAME row images that the forward / backward-input products use (no transposed
    images): reduction over the rows of both images, batch sizes that end inside a 64-row k-block and inside a 128-row
    image block, output sizes that end inside a 128-row / 256-column tile; also from an image written by a product's
    epilogue (its rows beyond the batch must be zero)."""
    gy, x = _rand((batch, n_out), 20), _rand((batch, n_in), 21)
    gw, _ = _ops.tc_gemm(_ops.tc_pack(gy, 128), _ops.tc_pack(x, 128), n_out, n_in, batch, c=True, split_k=split, mn_major=True)
    ref = _bf(gy).T @ _bf(x)
    assert gw.shape == (n_out, n_in)
    assert float((gw.double() - ref).abs().max()) < 3e-4 * (1 + float(ref.abs().max()))
    # the image of an activation as the forward epilogue writes it (bias makes the padding rows non-zero before masking)
    k0 = 100
    z, w, b = _rand((batch, k0), 22), _rand((n_in, k0), 23) / k0 ** 0.5, _rand((n_in,), 24) + 1.0
    _, himg = _ops.tc_gemm(_ops.tc_pack(z, 128), _ops.tc_pack(w, 256), batch, n_in, k0, bias=b, activation=_ops.ACT_ELU,
                           out_image=True)
    h = torch.nn.functional.elu(_bf(z) @ _bf(w).T + b.double())
    gw2, _ = _ops.tc_gemm(_ops.tc_pack(gy, 128), himg, n_out, n_in, batch, c=True, split_k=split, mn_major=True)
    ref2 = _bf(gy).T @ _bf(h.float())
    assert float((gw2.double() - ref2).abs().max()) < 3e-4 * (1 + float(ref2.abs().max()))


@pytest.mark.parametrize('m,n,k', [(1000, 670, 300), (257, 1500, 670), (129, 40, 70), (128 * 9, 256, 64)])
def test_cluster_mode_shares_the_b_operand(m, n, k):
    """Clusters of two CTAs, each fetching half of every B block and multicasting it to both: bit-identical results to the
    plain launch (same products in the same order), for even and odd numbers of m-tiles (an odd count leaves a ghost half
    that only keeps the hand-shakes going) and more tiles than clusters."""
    x, w, b = _rand((m, k), 31), _rand((n, k), 32) / k ** 0.5, _rand((n,), 33)
    ai, bi = _ops.tc_pack(x, 128), _ops.tc_pack(w, 256)
    c0, i0 = _ops.tc_gemm(ai, bi, m, n, k, c=True, bias=b, activation=_ops.ACT_ELU, out_image=True, cluster=False)
    c1, i1 = _ops.tc_gemm(ai, bi, m, n, k, c=True, bias=b, activation=_ops.ACT_ELU, out_image=True, cluster=True)
    assert torch.equal(c0, c1) and torch.equal(i0, i1)
    ref = torch.nn.functional.elu(_bf(x) @ _bf(w).T + b.double())
    assert float((c1.double() - ref).abs().max()) < 2e-4 * (1 + float(ref.abs().max()))
