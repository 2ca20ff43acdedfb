"""The unfused drop-in operators (csrc/geometry.cu) behind layers.BackprojectDepth / Project3D / Cat_xy /
SSIM / transformation_from_parameters and the layers.F proxy (interpolate, grid_sample) -- forward AND
gradients against the CPU oracle / torch's own CPU ops on identical inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF

from tests._util import rel_err
from oracle import step_oracle as SO

pytestmark = pytest.mark.gpu


def _K(B, H, W):
    K = torch.tensor([[0.58 * W, 0, 0.5 * W, 0], [0, 1.92 * H, 0.5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1.]])
    K = K.repeat(B, 1, 1)
    return K, torch.linalg.pinv(K)


@pytest.mark.parametrize("B,H,W", [(2, 16, 24), (3, 96, 160)])
def test_geometry_modules_forward_backward(cuda, B, H, W):
    from fusiondepth_b200 import layers as L
    g = torch.Generator().manual_seed(H)
    depth = (0.5 + torch.rand(B, 1, H, W, generator=g) * 10)
    K, invK = _K(B, H, W)
    aa = (0.02 * torch.randn(B, 1, 3, generator=g))
    tt = (0.1 * torch.randn(B, 1, 3, generator=g))
    img = torch.rand(B, 3, H, W, generator=g)
    tgt = torch.rand(B, 3, H, W, generator=g)
    wgt = torch.randn(B, 3, H, W, generator=g)

    def chain(depth, aa, tt, img, tgt, invert, lay, Fn, dev):
        T = lay["pose"](aa, tt, invert)
        pts = lay["bp"](depth, invK.to(dev))
        grid = lay["proj"](pts, K.to(dev), T)
        warped = Fn.grid_sample(img, grid, padding_mode="border", align_corners=False)
        s = lay["ssim"](warped, tgt)
        return T, pts, grid, warped, s

    for invert in (False, True):
        # oracle (CPU, torch autograd)
        od, oa, ot = (t.clone().requires_grad_(True) for t in (depth, aa, tt))
        olay = {"pose": SO.pose_matrix, "bp": SO.backproject,
                "proj": lambda p, K_, T_: SO.project(p, K_, T_, H, W), "ssim": SO.ssim}
        oT, opts, ogrid, owarp, oss = chain(od, oa, ot, img, tgt, invert, olay, TF, "cpu")
        ((oss * wgt).sum() / wgt.numel()).backward()
        # ours
        cd, ca, ct = (t.clone().cuda().requires_grad_(True) for t in (depth, aa, tt))
        clay = {"pose": L.transformation_from_parameters, "bp": L.BackprojectDepth(B, H, W).cuda(),
                "proj": L.Project3D(B, H, W).cuda(), "ssim": L.SSIM().cuda()}
        cT, cpts, cgrid, cwarp, css = chain(cd, ca, ct, img.cuda(), tgt.cuda(), invert, clay, L.F, "cuda")
        ((css * wgt.cuda()).sum() / wgt.numel()).backward()
        torch.cuda.synchronize()
        assert rel_err(cT.detach().cpu(), oT.detach()) < 1e-6
        assert rel_err(cpts.detach().cpu(), opts.detach()) < 1e-6
        assert rel_err(cgrid.detach().cpu(), ogrid.detach()) < 1e-5
        assert rel_err(cwarp.detach().cpu(), owarp.detach()) < 2e-4
        assert rel_err(css.detach().cpu(), oss.detach()) < 2e-4
        assert rel_err(cd.grad.cpu(), od.grad) < 2e-3
        assert rel_err(ca.grad.cpu(), oa.grad) < 2e-3
        assert rel_err(ct.grad.cpu(), ot.grad) < 2e-3
    # Cat_xy and a batch mismatch
    assert rel_err(L.Cat_xy(B, H, W).cuda()(depth.cuda(), invK.cuda()).cpu(), SO.cat_xy(depth, invK)) < 1e-6
    with pytest.raises(RuntimeError):
        L.BackprojectDepth(B + 1, H, W).cuda()(depth.cuda(), invK.cuda())


def test_ssim_gradient_both_arguments(cuda):
    from fusiondepth_b200 import layers as L
    g = torch.Generator().manual_seed(3)
    x, y = torch.rand(2, 3, 20, 28, generator=g), torch.rand(2, 3, 20, 28, generator=g)
    w = torch.randn(2, 3, 20, 28, generator=g)
    ox, oy = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    (SO.ssim(ox, oy) * w).sum().backward()
    cx, cy = x.cuda().requires_grad_(True), y.cuda().requires_grad_(True)
    out = L.SSIM()(cx, cy)
    (out * w.cuda()).sum().backward()
    assert rel_err(out.detach().cpu(), SO.ssim(x, y)) < 1e-5
    assert rel_err(cx.grad.cpu(), ox.grad) < 1e-3 and rel_err(cy.grad.cpu(), oy.grad) < 1e-3


@pytest.mark.parametrize("shape,size", [((2, 1, 24, 80), (192, 640)), ((2, 3, 12, 20), (48, 80)),
                                        ((1, 1, 192, 640), (24, 80)), ((2, 1, 17, 23), (40, 31))])
def test_interpolate_proxy_matches_torch(cuda, shape, size):
    from fusiondepth_b200 import layers as L
    g = torch.Generator().manual_seed(shape[2])
    x = torch.rand(*shape, generator=g)
    w = torch.randn(shape[0], shape[1], *size, generator=g)
    ox = x.clone().requires_grad_(True)
    oy = TF.interpolate(ox, list(size), mode="bilinear", align_corners=False)
    (oy * w).sum().backward()
    cx = x.cuda().requires_grad_(True)
    cy = L.F.interpolate(cx, list(size), mode="bilinear", align_corners=False)
    (cy * w.cuda()).sum().backward()
    assert rel_err(cy.detach().cpu(), oy.detach()) < 1e-6
    assert rel_err(cx.grad.cpu(), ox.grad) < 1e-5
    # anything else is torch's own functional
    assert torch.equal(L.F.interpolate(cx.detach(), scale_factor=2, mode="nearest"),
                       TF.interpolate(cx.detach(), scale_factor=2, mode="nearest"))


def test_grid_sample_proxy_input_gradient(cuda):
    from fusiondepth_b200 import layers as L
    g = torch.Generator().manual_seed(9)
    img = torch.rand(2, 3, 16, 24, generator=g)
    grid = torch.rand(2, 10, 14, 2, generator=g) * 2.4 - 1.2            # some samples beyond the border
    w = torch.randn(2, 3, 10, 14, generator=g)
    oi, og = img.clone().requires_grad_(True), grid.clone().requires_grad_(True)
    (TF.grid_sample(oi, og, padding_mode="border", align_corners=False) * w).sum().backward()
    ci, cg = img.cuda().requires_grad_(True), grid.cuda().requires_grad_(True)
    out = L.F.grid_sample(ci, cg, padding_mode="border")
    (out * w.cuda()).sum().backward()
    assert rel_err(out.detach().cpu(), TF.grid_sample(img, grid, padding_mode="border", align_corners=False)) < 1e-5
    assert rel_err(cg.grad.cpu(), og.grad) < 1e-4 and rel_err(ci.grad.cpu(), oi.grad) < 1e-4


def test_masked_median_matches_torch(cuda):
    from fusiondepth_b200 import ops
    g = torch.Generator().manual_seed(5)
    for n_on in (1, 2, 7, 4000):
        x = torch.randn(2, 1, 192, 640, generator=g)
        m = torch.zeros(2 * 192 * 640)
        m[torch.randperm(m.numel(), generator=g)[:n_on * 40]] = 1.0
        m = m.view(2, 1, 192, 640)
        win = (78, 190, 23, 617)
        crop = torch.zeros_like(m)
        crop[:, :, win[0]:win[1], win[2]:win[3]] = 1
        sel = (m * crop) > 0
        if int(sel.sum()) == 0:
            continue
        want = torch.median(x[sel] * 100.0)
        got = ops.masked_median(x.cuda(), m.cuda(), win, 100.0)
        assert float(got) == float(want), (n_on, float(got), float(want))
