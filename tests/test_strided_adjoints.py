"""The fused strided forms are each other's adjoints (what DESIGN.md 8 item 2 builds on: under autograd the input gradient of the up-2 composite is
the down-2 composite with channel-transposed weights, and vice versa, so both directions already exist as forward kernels).  Checked here in
float64 with the oracle's definitional pipeline against autograd, for the two call forms the networks use (SynthesisLayer up-2:
conv2d_resample(x, w, f, up=2, padding=1, flip_weight=False), training/networks.py:84-94 / :310; Conv2dLayer down-2: down=2, padding=1,
flip_weight=True, :197)."""
import pytest
import torch

from oracle import ops_oracle as O


@pytest.mark.parametrize('n,ci,co,h,w', [(2, 3, 4, 6, 6), (1, 5, 2, 7, 9), (3, 2, 2, 4, 12)])
def test_up2_and_down2_composites_are_adjoint(n, ci, co, h, w):
    torch.manual_seed(n * 100 + h)
    f = O.setup_filter([1, 3, 3, 1]).double()
    wt = torch.randn(co, ci, 3, 3, dtype=torch.double)
    # up-2: d/dx <y, dy> = 4 * down2(dy, w^T)
    x = torch.randn(n, ci, h, w, dtype=torch.double, requires_grad=True)
    y = O.conv2d_resample(x, wt, f, up=2, padding=1, flip_weight=False)
    assert tuple(y.shape) == (n, co, 2 * h, 2 * w)
    dy = torch.randn_like(y)
    dx, = torch.autograd.grad(y, x, dy)
    dx_adj = 4 * O.conv2d_resample(dy, wt.transpose(0, 1).contiguous(), f, down=2, padding=1, flip_weight=True)
    assert float((dx - dx_adj).abs().max()) < 1e-12 * float(dx.abs().max())
    # down-2: d/dx <y, dy> = up2(dy, w^T) / 4
    xd = torch.randn(n, ci, 2 * h, 2 * w, dtype=torch.double, requires_grad=True)
    yd = O.conv2d_resample(xd, wt, f, down=2, padding=1, flip_weight=True)
    assert tuple(yd.shape) == (n, co, h, w)
    dyd = torch.randn_like(yd)
    dxd, = torch.autograd.grad(yd, xd, dyd)
    dxd_adj = 0.25 * O.conv2d_resample(dyd, wt.transpose(0, 1).contiguous(), f, up=2, padding=1, flip_weight=False)
    assert float((dxd - dxd_adj).abs().max()) < 1e-12 * float(dxd.abs().max())
