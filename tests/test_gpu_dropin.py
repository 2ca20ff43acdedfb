"""The drop-in boundary, executed: the reference's UNCHANGED trainer.py / refiner.py (staged byte-for-byte
under oracle/_ref by oracle/make_ref.py, or /root/reference in the build container) driven against
fusiondepth_b200/dropin, i.e. `import networks` / `from layers import *` resolve to this repo's CUDA modules.

Three levels, all against the reference-generated fixture tests/golden/step_r18.npz (same seeds):
  * unpatched: Trainer.process_batch as written -- every layers.* module and F.* call of its loss chain lands
    on this library's unfused kernels (csrc/geometry.cu);
  * patch_trainer: generate_images_pred + compute_losses swapped for the fused fd_photoloss pair;
  * Trainer.val-style process_batch(val=True) + compute_depth_losses through the patched trainer.
"""
import numpy as np
import pytest
import torch

from tests._util import GOLDEN, rel_err, synth_weights
from fusiondepth_b200 import synth
from oracle import ref_harness as RH
from tests.test_gpu_refiner import _gtol

pytestmark = pytest.mark.gpu


def _trainer(B, H, W, seed=0):
    if not RH.available():
        pytest.skip("reference sources not staged (oracle/_ref)")
    ns = RH.load(dropin=True)
    assert ns.networks.ResnetEncoder.__module__.startswith("fusiondepth_b200"), "drop-in did not resolve"
    assert ns.trainer.__file__.startswith(ns.root), "trainer.py must be the reference's own file"
    models = RH.make_models(ns, 18)
    for i, (name, m) in enumerate(sorted(models.items())):
        m.load_state_dict(synth_weights(m.state_dict(), seed * 100 + i))
        m.cuda().train()
    return ns, models, RH.make_trainer(ns, models, B, H, W, device="cuda")


def _check_against_fixture(outputs, losses, models, g, grad_tol=5e-3):
    for k in losses:
        assert rel_err(losses[k].detach().cpu(), g["loss:" + k]) < 1e-4, (k, float(losses[k]), float(g["loss:" + k]))
    for s in range(4):
        assert rel_err(outputs[("disp", s)].detach().cpu(), g["disp%d" % s]) < 1e-4, s
        assert rel_err(outputs[("depth", 0, s)].detach().cpu(), g["depth%d" % s]) < 1e-4, s
        mism = (outputs["identity_selection/%d" % s].cpu().numpy() != g["identity_selection%d" % s]).mean()
        assert mism < 1e-3, (s, mism)
    for f in (-1, 1):
        assert rel_err(outputs[("cam_T_cam", 0, f)].detach().cpu(), g["cam_T_cam%d" % f]) < 1e-5
        assert rel_err(outputs[("color", f, 0)].detach().cpu(), g["color%d_0" % f]) < 2e-4
    bad, n = [], 0
    for key in g.files:
        if key.startswith("gnorm:"):
            name, pk = key[6:].split("/", 1)
            p = dict(models[name].named_parameters())[pk]
            got, want = float(p.grad.double().norm()), float(g[key])
            n += 1
            if abs(got - want) > grad_tol * want + 1e-9:
                bad.append((key, got, want))
    assert n > 200 and not bad, bad[:8]


@pytest.mark.parametrize("patched", [False, True])
def test_unchanged_trainer_dropin(cuda, patched):
    from fusiondepth_b200 import training
    g = np.load(GOLDEN + "/step_r18.npz")
    ns, models, tr = _trainer(3, 96, 160)
    if patched:
        training.patch_trainer(tr)
    inputs = synth.make_batch(3, 96, 160, seed=1, mode="coherent", lidar_density=0.25)
    noise = inputs.pop("noise")
    with RH.FixedNoise([noise[s] for s in range(4)]):
        outputs, losses = tr.process_batch(dict(inputs))
    losses["loss"].backward()
    torch.cuda.synchronize()
    _check_against_fixture(outputs, losses, models, g)


def test_unchanged_trainer_validation_path(cuda):
    """Trainer.val()'s body for one batch through the patched trainer: process_batch(val=True) must leave
    ("depth", 0, 0) for compute_depth_losses (trainer.py:402-405, 598-630)."""
    from fusiondepth_b200 import training
    ns, models, tr = _trainer(1, 96, 160)
    training.patch_trainer(tr)
    tr.depth_metric_names = ["de/abs_rel", "de/sq_rel", "de/rms", "de/log_rms", "da/a1", "da/a2", "da/a3"]
    for m in models.values():
        m.eval()
    inputs = synth.make_batch(1, 96, 160, seed=3, mode="coherent", lidar_density=0.25)
    inputs.pop("noise")
    with torch.no_grad():
        outputs, _ = tr.process_batch(dict(inputs), val=True)
        depth = outputs[("depth", 0, 0)]
        assert depth.shape == (1, 1, 96, 160)
        # reference semantics of the validation depth: disp_to_depth(bilinear(disp_0 -> HxW))
        want = 1.0 / (0.01 + 9.99 * outputs[("disp", 0)])
        assert rel_err(depth.cpu(), want.cpu()) < 1e-6
        rng = np.random.RandomState(0)
        gt = rng.uniform(2.0, 60.0, (1, 1, 375, 1242)).astype(np.float32)
        gt[rng.uniform(size=gt.shape) < 0.9] = 0
        batch = {"depth_gt": torch.from_numpy(gt).cuda()}
        losses = {}
        tr.compute_depth_losses(batch, outputs, losses)
    assert set(losses) == set(tr.depth_metric_names) and all(np.isfinite(v) for v in losses.values())


def test_unchanged_refiner_dropin(cuda):
    """refiner.py's Refiner.process_batch, unchanged, on this repo's modules (unfused kernels), and with
    refine.patch_refiner (fused pack + loss): both against the reference-generated fixture."""
    from fusiondepth_b200 import refine
    if not RH.available():
        pytest.skip("reference sources not staged (oracle/_ref)")
    g = np.load(GOLDEN + "/refiner.npz")
    ns = RH.load(dropin=True, with_refiner=True)
    for patched in (False, True):
        models = RH.make_models(ns, 18)
        models["refine2d_decoder"] = RH.make_refine_decoder(ns, models["encoder"].num_ch_enc)
        for i, (name, m) in enumerate(sorted(models.items())):
            m.load_state_dict(synth_weights(m.state_dict(), 4 * 100 + i))
            m.cuda().train()
        rf = RH.make_refiner(ns, models, 2, 192, 640, device="cuda")
        if patched:
            refine.patch_refiner(rf)
        inputs = synth.make_refiner_batch(2, 192, 640, seed=6)
        noise = inputs.pop("noise")
        with RH.FixedNoise([noise[s] for s in range(4)]):
            outputs, losses = rf.process_batch(dict(inputs))
        losses["loss"].backward()
        torch.cuda.synchronize()
        for k in losses:
            assert rel_err(losses[k].detach().cpu(), g["loss:" + k]) < 1e-4, (patched, k, float(losses[k]))
        for s in range(4):
            d = outputs[("disp", s)].detach().cpu()
            assert rel_err(d if s else d[:, :, ::4, ::4], g["disp%d" % s]) < 1e-4, (patched, s)
        dec = dict(models["refine2d_decoder"].named_parameters())
        bad = []
        for key in g.files:
            if key.startswith("gnorm:"):
                got, want = float(dec[key[6:]].grad.double().norm()), float(g[key])
                if abs(got - want) > _gtol(key) * want + 1e-9:
                    bad.append((key, got, want))
        assert not bad, (patched, bad[:8])


@pytest.mark.parametrize("patched", [False, True])
def test_unchanged_completor_dropin(cuda, patched):
    """The completion driver (driver half): the reference's UNCHANGED completor.py --
    Completor.process_batch + backward -- against fusiondepth_b200/dropin, unpatched (every layers.* / F.* call of
    its loss code on this library's kernels) and with patch_completor (fused loss), vs the fixture the reference
    itself produced (tests/make_golden.py gen_completor)."""
    from fusiondepth_b200 import training
    if not RH.available():
        pytest.skip("reference sources not staged (oracle/_ref)")
    g = np.load(GOLDEN + "/step_completor.npz")
    ns = RH.load(dropin=True, with_completor=True)
    assert ns.networks.ResnetEncoder.__module__.startswith("fusiondepth_b200"), "drop-in did not resolve"
    assert ns.completor.__file__.startswith(ns.root), "completor.py must be the reference's own file"
    models = RH.make_models(ns, 18)
    for i, (name, m) in enumerate(sorted(models.items())):
        m.load_state_dict(synth_weights(m.state_dict(), 300 + i))
        m.cuda().train()
    cp = RH.make_completor(ns, models, 2, 96, 160, device="cuda")
    if patched:
        training.patch_completor(cp)
    inputs = synth.make_batch(2, 96, 160, seed=6, mode="coherent", lidar_density=0.25)
    noise = inputs.pop("noise")
    with RH.FixedNoise([noise[s] for s in range(4)]):
        outputs, losses = cp.process_batch(dict(inputs))
    losses["loss"].backward()
    torch.cuda.synchronize()
    for key in g.files:
        if key.startswith("loss:"):
            k = key[5:]
            assert rel_err(losses[k].detach().cpu(), g[key]) < 1e-4, (k, float(losses[k]), float(g[key]))
    for s in range(4):
        assert rel_err(outputs[("disp", s)].detach().cpu(), g["disp%d" % s]) < 1e-4, s
        assert rel_err(outputs[("depth", 0, s)].detach().cpu(), g["depth%d" % s]) < 1e-4, s
        mism = (outputs["identity_selection/%d" % s].cpu().numpy() != g["identity_selection%d" % s]).mean()
        assert mism < 1e-3, (s, mism)
    for f in (-1, 1):
        assert rel_err(outputs[("cam_T_cam", 0, f)].detach().cpu(), g["cam_T_cam%d" % f]) < 1e-5
        assert rel_err(outputs[("color", f, 0)].detach().cpu(), g["color%d_0" % f]) < 2e-4
    bad, n = [], 0
    for key in g.files:
        if key.startswith("gnorm:"):
            name, pk = key[6:].split("/", 1)
            p = dict(models[name].named_parameters())[pk]
            got, want = float(p.grad.double().norm()), float(g[key])
            n += 1
            if abs(got - want) > _gtol(key) * want + 1e-9:
                bad.append((key, got, want))
    assert n > 200 and not bad, bad[:8]
