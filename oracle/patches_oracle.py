"""CPU oracle: dB feature (n_bins, T) -> (3, H, W) float32 patch tensors.  TEST INFRASTRUCTURE ONLY.

* ``vit_patch``      restates /root/reference/ViT_dataloader.py:27-51 with NumPy: ``(x+120)/120`` -> clip[0,1] ->
  bicubic ``F.interpolate(size, align_corners=False)`` -> ``repeat(3,1,1)``.  The bicubic restatement follows ATen's
  ``upsample_bicubic2d`` (A = -0.75, src = scale*(dst+0.5)-0.5 in float32, taps floor-1..floor+2 clamped) and is
  PINNED against the real ``torch.nn.functional.interpolate`` (installed here) by tests/test_oracle_patches.py.
* ``vit_patch_torch`` is the reference line itself executed with CPU torch.
* ``cnn_patch``      the tensor contract bestengine.py consumes from my_dataloader.py:17-21,29-30:
  (3,224,224) float32, ImageNet-normalised.  PARITY UNPINNED: the reference feeds matplotlib PNG pictures
  (new_cqt.py:33-42) through PIL; matplotlib's rasteriser/colormap is not a numeric contract.
  The contract used here: grey image = clip((dB+120)/120, 0, 1) with the highest bin on the top row
  (specshow's origin), bilinear half-pixel resize to 224x224, 3 identical channels, (x-mean)/std.
"""
from __future__ import annotations

import numpy as np

IMAGENET_MEAN = (0.485, 0.456, 0.406)     # my_dataloader.py:20
IMAGENET_STD = (0.229, 0.224, 0.225)


def _cubic_coeffs(t):
    """ATen get_cubic_upsample_coefficients, float32."""
    A = np.float32(-0.75)
    one = np.float32(1.0)

    def conv1(x):   # |x| <= 1
        return ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one

    def conv2(x):   # 1 < |x| < 2
        return ((A * x - np.float32(5) * A) * x + np.float32(8) * A) * x - np.float32(4) * A

    t = np.asarray(t, dtype=np.float32)
    return np.stack([conv2(t + one), conv1(t), conv1(one - t), conv2((one - t) + one)], axis=-1).astype(np.float32)


def _axis_taps(n_in, n_out):
    """Per output index: 4 clamped source indices and 4 float32 weights (align_corners=False, cubic)."""
    scale = np.float32(n_in) / np.float32(n_out)
    dst = np.arange(n_out, dtype=np.float32)
    src = scale * (dst + np.float32(0.5)) - np.float32(0.5)
    fl = np.floor(src)
    t = (src - fl).astype(np.float32)
    base = fl.astype(np.int64)
    idx = np.clip(base[:, None] + np.arange(-1, 3)[None, :], 0, n_in - 1)
    return idx, _cubic_coeffs(t)


def bicubic_resize(img, out_h, out_w):
    """img (H, W) float32 -> (out_h, out_w) float32; x-interpolation first, then y (ATen order)."""
    img = np.asarray(img, dtype=np.float32)
    iy, wy = _axis_taps(img.shape[0], out_h)
    ix, wx = _axis_taps(img.shape[1], out_w)
    rows = np.zeros((img.shape[0], out_w), dtype=np.float32)
    for j in range(4):
        rows += img[:, ix[:, j]] * wx[None, :, j]
    out = np.zeros((out_h, out_w), dtype=np.float32)
    for i in range(4):
        out += rows[iy[:, i], :] * wy[:, i, None]
    return out


def vit_normalize(db):
    """ViT_dataloader.py:27-32."""
    audio = np.asarray(db).astype(np.float32)
    return np.clip((audio + 120) / 120, 0, 1)


def vit_patch(db, img_size=(224, 224)):
    x = bicubic_resize(vit_normalize(db), img_size[0], img_size[1])
    return np.repeat(x[None], 3, axis=0)


def vit_patch_torch(db, img_size=(224, 224)):
    """The reference's own lines (ViT_dataloader.py:35-51) on CPU torch."""
    import torch
    t = torch.tensor(vit_normalize(db)).unsqueeze(0)
    t = torch.nn.functional.interpolate(t.unsqueeze(0), size=img_size, mode='bicubic', align_corners=False).squeeze(0)
    return t.repeat(3, 1, 1).numpy()


def bilinear_resize(img, out_h, out_w):
    """Half-pixel-centre bilinear (what PIL/torchvision Resize does when up-sampling), float32."""
    img = np.asarray(img, dtype=np.float32)

    def taps(n_in, n_out):
        scale = np.float32(n_in) / np.float32(n_out)
        src = scale * (np.arange(n_out, dtype=np.float32) + np.float32(0.5)) - np.float32(0.5)
        src = np.maximum(src, np.float32(0))
        i0 = np.minimum(np.floor(src).astype(np.int64), n_in - 1)
        i1 = np.minimum(i0 + 1, n_in - 1)
        l1 = (src - i0.astype(np.float32)).astype(np.float32)
        return i0, i1, (np.float32(1) - l1), l1

    y0, y1, wy0, wy1 = taps(img.shape[0], out_h)
    x0, x1, wx0, wx1 = taps(img.shape[1], out_w)
    rows = img[:, x0] * wx0[None, :] + img[:, x1] * wx1[None, :]
    return (rows[y0, :] * wy0[:, None] + rows[y1, :] * wy1[:, None]).astype(np.float32)


def cnn_patch(db, img_size=(224, 224), flip=True):
    g = vit_normalize(db)
    if flip:
        g = g[::-1, :]
    x = bilinear_resize(g, img_size[0], img_size[1])
    mean = np.asarray(IMAGENET_MEAN, dtype=np.float32)[:, None, None]
    std = np.asarray(IMAGENET_STD, dtype=np.float32)[:, None, None]
    return ((x[None] - mean) / std).astype(np.float32)
