"""tfepb_wn_pack: the packed effective weight of a weight-normalised masked layer (nn/masked.py:369-371, 433-439, 270 + the
degree-sorted packing) and its VJP in one launch each, against the tensor-algebra formulation that autograd differentiates
(``masked.effective_weight`` + ``MadePlan.pack``).  fp32 both ways: agreement to summation order (1e-6 relative)."""

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize('rows,cols,pad', [(37, 19, 0), (670, 300, 0), (1500, 670, 100), (5, 1, 3)])
def test_wn_pack_matches_effective_weight_and_autograd(rows, cols, pad):
    from tfep_b200 import _ops
    from tfep_b200.nn import masked
    g0 = torch.Generator().manual_seed(rows * 1000 + cols)
    v = torch.randn(rows, cols, generator=g0)
    mask = (torch.rand(rows, cols, generator=g0) > 0.4).float()
    v = v * mask
    v[rows // 2] = 0.0                                           # a row of zero norm: weight 0, no NaN, g gets no gradient
    g = torch.randn(rows, 1, generator=g0)
    b = torch.randn(rows, generator=g0)
    row_perm = torch.randperm(rows, generator=g0)
    if pad:
        row_perm = torch.cat([row_perm, torch.full((pad,), -1, dtype=torch.long)])[torch.randperm(rows + pad, generator=g0)]
    col_perm = torch.randperm(cols, generator=g0)
    cw = torch.randn(len(row_perm), cols, generator=g0).to(DEV)
    cb = torch.randn(len(row_perm), generator=g0).to(DEV)

    def leaves():
        return [t.clone().to(DEV).requires_grad_(True) for t in (v, g, b)]

    # reference formulation
    v1, g1, b1 = leaves()
    w = masked.effective_weight(v1, g1, mask.to(DEV))
    sel = row_perm.clamp(min=0).to(DEV)
    keep = (row_perm >= 0).to(DEV)
    w_ref = w.index_select(0, sel).index_select(1, col_perm.to(DEV)) * keep[:, None]
    b_ref = b1.index_select(0, sel) * keep
    ((w_ref * cw).sum() + (b_ref * cb).sum()).backward()
    # one launch
    v2, g2, b2 = leaves()
    w_out, b_out = _ops.wn_pack(v2, g2, b2, mask.to(DEV), row_perm.to(DEV, torch.int32), col_perm.to(DEV, torch.int32))
    assert w_out.shape == w_ref.shape and w_out.stride(0) % 4 == 0
    ((w_out * cw).sum() + (b_out * cb).sum()).backward()
    assert _rel(w_out.detach(), w_ref.detach()) < 1e-6 and torch.equal(b_out.detach(), b_ref.detach())
    assert torch.isfinite(v2.grad).all() and torch.isfinite(g2.grad).all()
    def close(a, b, tol):            # (a single-column layer cancels to rounding noise: absolute floor)
        return float((a - b).abs().max()) < tol * (1.0 + float(b.abs().max()))

    assert close(v2.grad, v1.grad, 5e-6) and close(g2.grad, g1.grad, 5e-6) and close(b2.grad, b1.grad, 1e-7)
    # masked entries of v (zero by construction) receive zero gradient, as with the reference's hooks (nn/masked.py:400-402)
    assert float((v2.grad * (1 - mask.to(DEV))).abs().max()) == 0.0


def test_training_step_uses_the_fused_weight_path():
    """MADE.packed_weights under autograd = wn_pack per layer: same conditioner output and parameter gradients as the
    tensor-algebra path (forced by a bias-free check bypass: float64 parameters take the generic route)."""
    from tfep_b200.nn.conditioners.made import MADE, generate_degrees
    torch.manual_seed(3)
    made = MADE(degrees_in=generate_degrees(12), degrees_out=generate_degrees(12).tile((2,)), hidden_layers=2).to(DEV)
    x = torch.randn(50, 12, device=DEV)
    c = torch.randn(50, 24, device=DEV)
    (made(x) * c).sum().backward()
    fused = {k: p.grad.clone() for k, p in made.named_parameters()}
    made64 = MADE(degrees_in=generate_degrees(12), degrees_out=generate_degrees(12).tile((2,)), hidden_layers=2).to(DEV).double()
    made64.load_state_dict({k: t.double() for k, t in made.state_dict().items()})
    (made64(x.double()) * c.double()).sum().backward()
    for k, p in made64.named_parameters():
        assert _rel(fused[k].double(), p.grad) < 1e-4, k
