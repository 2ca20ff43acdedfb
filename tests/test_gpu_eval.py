"""Inference / evaluation path: BatchNorm-folded forward vs the reference fixture,
evaluate_depth.py's AbsRel harness and Trainer.compute_depth_losses on the device vs their CPU restatements."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests._util import GOLDEN, rel_err, synth_weights
from oracle import step_oracle as SO

pytestmark = pytest.mark.gpu


def _eval_models(num_layers=18):
    from fusiondepth_b200 import networks
    enc = networks.ResnetEncoder(num_layers, False)
    benc = networks.ResnetEncoder(num_layers, False, beam_encoder=True)
    dec = networks.DepthDecoder(enc.num_ch_enc, [0, 1, 2, 3])
    for i, m in enumerate((enc, benc, dec)):
        m.load_state_dict(synth_weights(m.state_dict(), 1000 + num_layers * 10 + i))
        m.cuda().eval()
    return {"encoder": enc, "beam_encoder": benc, "depth": dec}


@pytest.mark.parametrize("num_layers", [18, 50])
def test_folded_eval_forward_vs_reference_fixture(cuda, num_layers):
    """EvalRunner (BatchNorm folded into the convolutions, CUDA graph) against the reference's eval-mode
    disparities, and its AbsRel against the AbsRel of the reference's disparity on the same ground truth."""
    from fusiondepth_b200 import evaluation
    g = np.load(GOLDEN + "/forward_variants.npz")
    rgb, two = torch.from_numpy(g["rgb"]).cuda(), torch.from_numpy(g["two"]).cuda()
    H, W = rgb.shape[-2:]
    run = evaluation.EvalRunner(_eval_models(num_layers), 1, H, W)
    for _ in range(2):                                               # the graph replays
        disp = run({("color", 0, 0): rgb, "2channel": two})
    ref = SO.disp_to_depth(torch.from_numpy(g["r%d/disp0" % num_layers]), 0.1, 100.0)[0][:, 0]
    assert rel_err(disp.cpu(), ref) < 1e-4
    rng = np.random.RandomState(7)
    gt_h, gt_w = 2 * H - 9, 2 * W - 38
    gt = rng.uniform(2.0, 60.0, (gt_h, gt_w)).astype(np.float32)
    gt[rng.uniform(size=gt.shape) < 0.7] = 0.0
    ours = evaluation.evaluate_frame(disp[0], torch.from_numpy(gt).cuda()).cpu()
    want = SO.eval_abs_rel(ref[0].numpy(), gt)
    assert abs(float(ours[0]) - want) < 1e-4, (float(ours[0]), want)


def _np_errors(gt, pred):
    # evaluate_depth.py:42-60
    thresh = np.maximum(gt / pred, pred / gt)
    return np.array([np.mean(np.abs(gt - pred) / gt), np.mean((gt - pred) ** 2 / gt), np.sqrt(((gt - pred) ** 2).mean()),
                     np.sqrt(((np.log(gt) - np.log(pred)) ** 2).mean()), (thresh < 1.25).mean(),
                     (thresh < 1.25 ** 2).mean(), (thresh < 1.25 ** 3).mean()])


@pytest.mark.parametrize("n_even", [True, False])
def test_evaluate_frame_matches_evaluate_depth_recipe(cuda, n_even):
    """All seven metrics + the scaling ratio vs a numpy restatement of evaluate_depth.py:344-378, 470-478
    (numpy.median: mean of the middle two for even counts)."""
    from fusiondepth_b200 import evaluation
    rng = np.random.RandomState(3 + n_even)
    h, w, gh, gw = 192, 640, 375, 1242
    disp = (0.02 + rng.uniform(size=(h, w)) * 0.5).astype(np.float32)
    gt = rng.uniform(0.5, 95.0, (gh, gw)).astype(np.float32)             # some beyond MAX_DEPTH
    gt[rng.uniform(size=gt.shape) < 0.95] = 0.0
    c = evaluation.garg_crop(gh, gw)
    mask = np.logical_and(gt > 1e-3, gt < 80)
    cm = np.zeros(mask.shape, bool); cm[c[0]:c[1], c[2]:c[3]] = True
    mask &= cm
    if (mask.sum() % 2 == 0) != n_even:                                   # force the parity under test
        ys, xs = np.nonzero(mask)
        gt[ys[0], xs[0]] = 0.0
        mask[ys[0], xs[0]] = False
    up = F.interpolate(torch.from_numpy(disp)[None, None], (gh, gw), mode="bilinear", align_corners=False)[0, 0].numpy()
    pred = 1 / up
    ratio = np.median(gt[mask]) / np.median(pred[mask])
    p = np.clip(pred[mask] * ratio, 1e-3, 80)
    want = _np_errors(gt[mask].astype(np.float64), p.astype(np.float64))
    got = evaluation.evaluate_frame(torch.from_numpy(disp).cuda(), torch.from_numpy(gt).cuda()).cpu().numpy()
    assert int(got[7]) == int(mask.sum())
    assert abs(got[8] - ratio) < 1e-6 * ratio, (got[8], ratio)
    assert np.allclose(got[:7], want, rtol=2e-5, atol=1e-7), (got[:7], want)


def test_compute_depth_losses_matches_trainer(cuda):
    """Trainer.compute_depth_losses (trainer.py:598-630) restated with torch ops on the CPU vs the device path."""
    from fusiondepth_b200 import evaluation
    from fusiondepth_b200.layers import compute_depth_errors
    g = torch.Generator().manual_seed(5)
    depth = 0.5 + 30 * torch.rand(2, 1, 192, 640, generator=g)
    gt = 1.0 + 70 * torch.rand(2, 1, 375, 1242, generator=g)
    gt = gt * (torch.rand(2, 1, 375, 1242, generator=g) < 0.05).float()
    # reference recipe on the CPU
    dp = torch.clamp(F.interpolate(depth, [375, 1242], mode="bilinear", align_corners=False), 1e-3, 80)
    mask = gt > 0
    crop = torch.zeros_like(mask); crop[:, :, 153:371, 44:1197] = 1
    mask = mask * crop
    gm, pm = gt[mask], dp[mask]
    pm = torch.clamp(pm * (torch.median(gm) / torch.median(pm)), min=1e-3, max=80)
    want = [float(v) for v in compute_depth_errors(gm, pm)]
    losses = {}
    evaluation.compute_depth_losses({"depth_gt": gt.cuda()}, {("depth", 0, 0): depth.cuda()}, losses)
    got = [float(losses[k]) for k in evaluation.DEPTH_METRIC_NAMES]
    assert np.allclose(got, want, rtol=2e-5, atol=1e-7), (got, want)


def test_folded_forward_at_completion_resolution(cuda):
    """The completion driver's working size (completor.py:31-34 forces 352 x 1216): BN-folded encoder + beam encoder
    + decoder forward against the CPU oracle's eval-mode forward on the same input."""
    from fusiondepth_b200 import evaluation, synth
    H, W = 352, 1216
    models = _eval_models(18)
    g = torch.Generator().manual_seed(3)
    rgb = torch.rand(1, 3, H, W, generator=g)
    two = torch.zeros(1, 2, H, W)
    mask = torch.rand(1, 1, H, W, generator=g) < 0.03
    two[:, 0:1] = mask * torch.rand(1, 1, H, W, generator=g) * 0.6
    two[:, 1:2] = mask.float()
    sds = {k: {n: t.detach().cpu() for n, t in m.state_dict().items()} for k, m in models.items()}
    with torch.no_grad():
        feats = SO.resnet_encoder(sds["encoder"], rgb, 18, training=False)
        beam = SO.resnet_encoder(sds["beam_encoder"], two, 18, training=False)
        ref = SO.disp_to_depth(SO.depth_decoder(sds["depth"], feats, beam_feats=beam)[("disp", 0)], 0.1, 100.0)[0][:, 0]
    run = evaluation.EvalRunner(models, 1, H, W)
    disp = run({("color", 0, 0): rgb.cuda(), "2channel": two.cuda()})
    assert rel_err(disp.cpu(), ref) < 1e-4
