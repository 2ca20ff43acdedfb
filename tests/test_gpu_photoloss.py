"""Fused photometric loss chain (CUDA) vs reference fixtures and the CPU oracle.
Tolerance: 1e-4 relative fp32 on loss / depth / gradient tensors (north_star)."""
import numpy as np
import pytest
import torch

from tests._util import GOLDEN, rel_err
from fusiondepth_b200 import synth
from oracle import step_oracle as SO

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _run_cuda(inputs, noise, disps, Ts, materialize=True):
    from fusiondepth_b200 import training
    dev = torch.device("cuda:0")
    inp = synth.to_device(inputs, dev)
    nz = {s: noise[s].to(dev) for s in noise}
    outputs = {}
    leaves = {}
    for s in range(4):
        leaves["disp%d" % s] = disps[("disp", s)].detach().to(dev).requires_grad_(True)
        outputs[("disp", s)] = leaves["disp%d" % s]
    for f in (-1, 1):
        leaves["T%d" % f] = Ts[f].detach().to(dev).requires_grad_(True)
        outputs[("cam_T_cam", 0, f)] = leaves["T%d" % f]
    losses = training.fused_losses(inp, outputs, nz, None, materialize)
    losses["loss"].backward()
    torch.cuda.synchronize()
    return losses, outputs, leaves


def test_loss_chain_vs_reference_fixture(cuda):
    from tests.test_oracle_golden import loss_chain_inputs
    g = np.load(GOLDEN + "/loss_chain.npz")
    inputs, noise = loss_chain_inputs(g)
    disps = {("disp", s): torch.from_numpy(g["in:disp%d" % s]) for s in range(4)}
    Ts = {f: torch.from_numpy(g["in:T%d" % f]) for f in (-1, 1)}
    losses, outputs, leaves = _run_cuda(inputs, noise, disps, Ts)
    for k in ("loss", "loss/0", "loss/1", "loss/2", "loss/3", "loss/si_loss0", "loss/si_loss1",
              "loss/si_loss2", "loss/si_loss3"):
        assert rel_err(losses[k].cpu(), g["loss:" + k]) < RTOL, (k, float(losses[k]), float(g["loss:" + k]))
    for s in range(4):
        assert rel_err(outputs[("depth", 0, s)].cpu(), g["depth%d" % s]) < RTOL
        mism = (outputs["identity_selection/%d" % s].cpu().numpy() != g["identity_selection%d" % s]).mean()
        assert mism < 1e-3, (s, mism)
    for s in (0, 3):
        for f in (-1, 1):
            assert rel_err(outputs[("color", f, s)].cpu(), g["color%d_%d" % (f, s)]) < RTOL
    for s in range(4):
        assert rel_err(leaves["disp%d" % s].grad.cpu(), g["grad:disp%d" % s]) < 2e-3, s
    for f in (-1, 1):
        assert rel_err(leaves["T%d" % f].grad.cpu(), g["grad:T%d" % f]) < 2e-3, f


@pytest.mark.parametrize("B,H,W,mode", [(1, 32, 64, "uniform"), (2, 64, 96, "coherent"),
                                        (3, 96, 128, "coherent")])
def test_loss_chain_vs_oracle(cuda, B, H, W, mode):
    inputs = synth.make_batch(B, H, W, seed=B * 7 + H, mode=mode)
    noise = inputs.pop("noise")
    g = torch.Generator().manual_seed(H)
    disps = {("disp", s): torch.sigmoid(torch.randn(B, 1, H >> s, W >> s, generator=g) * 0.5 - 1)
             for s in range(4)}
    Ts = {}
    for f in (-1, 1):
        Ts[f] = SO.pose_matrix(0.01 * torch.randn(B, 1, 3, generator=g), 0.05 * torch.randn(B, 1, 3, generator=g), f < 0)
    # a beam map that populates the si-loss mask
    with torch.no_grad():
        up = torch.nn.functional.interpolate(disps[("disp", 0)], [H, W], mode="bilinear", align_corners=False)
        dep = 26.0 / (0.01 + 9.99 * up)
        mask = (torch.rand(B, 1, H, W, generator=g) < 0.1).float()
        inputs["4beam"] = mask * (dep + (torch.rand(B, 1, H, W, generator=g) - 0.5) * 3.0) / 100.0
    od = {k: v.clone().requires_grad_(True) for k, v in disps.items()}
    oT = {f: Ts[f].clone().requires_grad_(True) for f in Ts}
    ol, oo = SO.photometric_chain(inputs, od, oT, noise)
    ol["loss"].backward()
    losses, outputs, leaves = _run_cuda(inputs, noise, disps, Ts)
    for k in ol:
        assert rel_err(losses[k].cpu(), ol[k].detach()) < RTOL, (k, float(losses[k]), float(ol[k]))
    for s in range(4):
        assert rel_err(outputs[("depth", 0, s)].cpu(), oo[("depth", 0, s)].detach()) < RTOL
        assert rel_err(outputs["to_optimise/%d" % s].cpu(), oo["to_optimise/%d" % s].detach()) < 2e-4
        for f in (-1, 1):
            assert rel_err(outputs[("color", f, s)].cpu(), oo[("color", f, s)].detach()) < 2e-4
        assert rel_err(leaves["disp%d" % s].grad.cpu(), od[("disp", s)].grad) < 2e-3, s
    for f in (-1, 1):
        assert rel_err(leaves["T%d" % f].grad.cpu(), oT[f].grad) < 2e-3, f


def test_full_size_properties(cuda):
    """640x192 batch 6 (BASELINE config 2's micro-batch): size-independent properties."""
    from fusiondepth_b200 import training
    dev = torch.device("cuda:0")
    B, H, W = 6, 192, 640
    inputs = synth.to_device(synth.make_batch(B, H, W, seed=2, mode="coherent"), dev)
    noise = inputs.pop("noise")
    g = torch.Generator().manual_seed(0)
    outputs = {("disp", s): torch.sigmoid(torch.randn(B, 1, H >> s, W >> s, generator=g) - 1).to(dev)
               for s in range(4)}
    eye = torch.eye(4, device=dev).repeat(B, 1, 1)
    outputs[("cam_T_cam", 0, -1)] = eye.clone()
    outputs[("cam_T_cam", 0, 1)] = eye.clone()
    # 1. sources identical to the target: identity terms are exactly noise * 1e-5
    same = dict(inputs)
    same[("color", -1, 0)] = inputs[("color", 0, 0)]
    same[("color", 1, 0)] = inputs[("color", 0, 0)]
    zero_noise = {s: torch.zeros_like(noise[s]) for s in noise}
    o1 = dict(outputs)
    l1 = training.fused_losses(same, o1, zero_noise, {"use_si": False}, materialize=True)
    for s in range(4):
        assert float(o1["to_optimise/%d" % s].abs().max()) == 0.0
        assert float(o1["identity_selection/%d" % s].sum()) == 0.0
    # 2. determinism: two launches give bit-identical losses
    a = training.fused_losses(inputs, dict(outputs), noise, None)
    b = training.fused_losses(inputs, dict(outputs), noise, None)
    for k in a:
        assert torch.equal(a[k], b[k])
    # 3. total = (sum loss/s + sum si)/4
    tot = sum(float(a["loss/%d" % s]) + float(a["loss/si_loss%d" % s]) for s in range(4)) / 4
    assert abs(tot - float(a["loss"])) < 1e-6 * max(1.0, abs(tot))
    # 4. depth range and identity warp sub-pixel shift (grid normalised by W-1, sampled with
    #    align_corners=False -- SURVEY.md Appendix E.3)
    o4 = dict(outputs)
    training.fused_losses(inputs, o4, noise, None, materialize=True)
    d = o4[("depth", 0, 0)]
    assert float(d.min()) >= 0.1 - 1e-6 and float(d.max()) <= 100.0 + 1e-4
    x = torch.arange(W, device=dev, dtype=torch.float32)
    sx = (x * W / (W - 1) - 0.5).clamp(0, W - 1)
    src = inputs[("color", 1, 0)]
    x0 = sx.floor().long()
    x1 = (x0 + 1).clamp(max=W - 1)
    fx = sx - x0
    y = torch.arange(H, device=dev, dtype=torch.float32)
    sy = (y * H / (H - 1) - 0.5).clamp(0, H - 1)
    y0 = sy.floor().long()
    y1 = (y0 + 1).clamp(max=H - 1)
    fy = (sy - y0).view(1, 1, H, 1)
    row = lambda yy: src[:, :, yy][:, :, :, x0] * (1 - fx) + src[:, :, yy][:, :, :, x1] * fx
    expect = row(y0) * (1 - fy) + row(y1) * fy
    assert float((o4[("color", 1, 0)] - expect).abs().max()) < 2e-4
