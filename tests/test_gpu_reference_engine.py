"""SURVEY.md 8(a20) executed by the consumer itself: the reference's OWN training engines run on the drop-in loaders.

``oracle/_ref/*.bin`` is the reference's ``bestengine.py`` / ``ViT_engine.py`` / ``ViT_model.py`` byte-compiled where they lie
(oracle/build_ref.py, build container only; git-ignored binaries that travel with the snapshot).  Their ``train_model`` /
``validate_model`` / ``test_model`` / ``visualize_sample_images`` are imported unmodified and driven for one epoch on
``my_dataloader.create_dataloaders`` / ``ViT_dataloader.create_dataloaders`` of THIS repo -- ``tqdm(loader)``, ``len()``,
``next(iter())``, ``inputs.to(device)``, ``labels[:, i]``, ``images[indices]``, the six-head label list: whatever the engines
do with a batch, they do here.  Stand-ins only for what the image lacks: matplotlib / seaborn (plot sinks) and the two
network downloads (ImageNet ResNet18 weights, the facebook/dino-vits8 checkpoint -> same architecture, random weights).
Skipped where oracle/_ref was not built (no /root/reference at build time).
"""
import os
import sys
import types
from unittest import mock

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build_ref  # noqa: E402

pytestmark = pytest.mark.gpu


def _plot_sinks():
    """matplotlib / seaborn are not installed: every plotting call lands in a MagicMock."""
    mods = {}
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        m = mock.MagicMock(name=name)
        m.__spec__ = None
        mods[name] = m
    mods["matplotlib"].pyplot = mods["matplotlib.pyplot"]
    # plt.subplots(r, c) must unpack
    mods["matplotlib.pyplot"].subplots.side_effect = lambda *a, **k: (mock.MagicMock(), mock.MagicMock())
    return mods


@pytest.fixture(scope="module")
def feature_dirs(tmp_path_factory, lib):
    """A small dataset written by this repo's own front end: dB feature .npy files + (6,19) label .npy files."""
    from gtc_b200 import audio_io, ops, synth, CqtRecipe
    root = tmp_path_factory.mktemp("engine_ds")
    (root / "cqt").mkdir(); (root / "tab").mkdir()
    dev = torch.device("cuda", torch.cuda.current_device())
    recipe = CqtRecipe()
    sr = int(recipe.sr)
    audio = synth.pluck_clips(3, sr * 8, sr=sr, seed=5)
    plan = ops.CqtPlan(recipe)
    lens = [audio.shape[1]] * 3
    clip_off, seg_off = plan.offsets(lens)
    n_seg = int(seg_off[-1])
    db = plan.segments_db(audio.reshape(-1).to(dev), torch.from_numpy(clip_off).to(dev), torch.from_numpy(seg_off).to(dev), n_seg)
    on, du, pi, eoff = synth.note_events([8.0] * 3, seed=6)
    per = n_seg // 3
    times = np.concatenate([(np.arange(per) + 0.5) * (8.0 / per)] * 3)
    t_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tabs, _ = ops.rasterize_tabs(t_(on), t_(du), t_(pi), t_(eoff), t_(times), t_(seg_off))
    db, tabs = db.cpu().numpy(), tabs.cpu().numpy()
    for i in range(n_seg):
        audio_io.save_feature(root / "cqt" / f"clip_segment_{i:05d}.npy", db[i])
        audio_io.save_label(root / "tab" / f"clip_segment_{i:05d}.npy", tabs[i])
    return root, n_seg


@pytest.mark.skipif(not build_ref.available("bestengine"), reason="oracle/_ref/bestengine.bin not built (python oracle/build_ref.py needs /root/reference)")
def test_reference_cnn_engine_trains_on_the_drop_in_loader(feature_dirs, tmp_path, monkeypatch):
    import torchvision
    root, n_seg = feature_dirs
    monkeypatch.chdir(tmp_path)                                  # the engine writes best_guitar_tab_model.pt / *.png to the cwd
    real_resnet18 = torchvision.models.resnet18
    monkeypatch.setattr(torchvision.models, "resnet18", lambda pretrained=False, **kw: real_resnet18(weights=None))   # no network
    with mock.patch.dict(sys.modules, _plot_sinks()):
        be = build_ref.load("bestengine", "reference_bestengine")
        # torch 2.11 dropped ReduceLROnPlateau(verbose=...), which bestengine.py:875 still passes: drop that keyword only
        real_plateau = be.ReduceLROnPlateau
        be.ReduceLROnPlateau = lambda opt, **kw: real_plateau(opt, **{k: v for k, v in kw.items() if k != "verbose"})
        import my_dataloader                                     # this repo's drop-in (the engine imports it by this name, bestengine.py:1043)
        train_loader, val_loader, test_loader = my_dataloader.create_dataloaders(str(root / "cqt"), str(root / "tab"), batch_size=32)
        assert len(train_loader.dataset) + len(val_loader.dataset) + len(test_loader.dataset) == n_seg
        be.set_seed(0)
        model = be.GuitarTabNet().to("cuda")
        be.visualize_sample_images(train_loader)                 # next(iter(loader)), images[indices], labels[idx] (bestengine.py:440-455)
        model, best_epoch, accuracies, (train_losses, val_losses, string_acc) = be.train_model(
            model=model, train_loader=train_loader, val_loader=val_loader, epochs=2, device="cuda", lr=0.0005)
        assert len(train_losses) == 2 and np.isfinite(train_losses).all() and np.isfinite(val_losses).all()
        assert len(accuracies) == 6 and all(0.0 <= a <= 100.0 for a in accuracies)
        assert os.path.exists("best_guitar_tab_model.pt")
        crit = be.LabelSmoothingLoss(classes=19, smoothing=0.05)
        loss, acc = be.validate_model(model, test_loader, crit, "cuda")
        assert np.isfinite(loss) and len(acc) == 6
        be.test_model(model, test_loader, "cuda")


@pytest.mark.skipif(not (build_ref.available("ViT_engine") and build_ref.available("ViT_model")),
                    reason="oracle/_ref/ViT_engine.bin not built (python oracle/build_ref.py needs /root/reference)")
def test_reference_vit_engine_trains_on_the_drop_in_loader(feature_dirs, tmp_path, monkeypatch):
    import transformers
    root, n_seg = feature_dirs
    monkeypatch.chdir(tmp_path)
    cfg = transformers.ViTConfig(hidden_size=384, num_hidden_layers=12, num_attention_heads=6, intermediate_size=1536,
                                 patch_size=8, image_size=224)   # the facebook/dino-vits8 architecture, random weights
    monkeypatch.setattr(transformers.ViTModel, "from_pretrained", classmethod(lambda cls, *a, **k: transformers.ViTModel(cfg)))
    monkeypatch.setattr(transformers.ViTImageProcessor, "from_pretrained", classmethod(lambda cls, *a, **k: mock.MagicMock()))
    mods = _plot_sinks()
    with mock.patch.dict(sys.modules, mods):
        sys.modules["ViT_model"] = build_ref.load("ViT_model", "ViT_model")
        try:
            # transformers 5 removed the deprecated ViTFeatureExtractor that ViT_engine.py:14 imports and never uses
            # (the lazy module re-registers itself in sys.modules on first use, so patch the registered object)
            tr = sys.modules["transformers"]
            if not hasattr(tr, "ViTFeatureExtractor"):
                monkeypatch.setattr(tr, "ViTFeatureExtractor", mock.MagicMock(), raising=False)
            ve = build_ref.load("ViT_engine", "reference_ViT_engine")     # imports THIS repo's ViT_dataloader (ViT_engine.py:13)
            import ViT_dataloader
            assert ve.create_dataloaders is ViT_dataloader.create_dataloaders
            train_loader, val_loader, test_loader = ve.create_dataloaders(str(root / "cqt"), str(root / "tab"), batch_size=25)
            model = ve.ViTGuitarTabModel().to("cuda")
            model, best_epoch, accuracies = ve.train_model(model, train_loader, val_loader, epochs=1, device="cuda", lr=0.0005)
            assert len(accuracies) == 6 and os.path.exists("best_vit_guitar_tab_model.pt")
            crit = ve.LabelSmoothingLoss(classes=19, smoothing=0.1)
            loss, acc = ve.validate_model(model, test_loader, crit, "cuda")[:2]
            assert np.isfinite(loss)
        finally:
            sys.modules.pop("ViT_model", None)
