"""Device-side datasets and loaders behind the drop-in ``my_dataloader`` / ``ViT_dataloader`` modules.

The reference forks DataLoader workers that np.load one tiny file per item and resize it on the CPU
(ViT_dataloader.py:22-56, :74-86).  Here all dB features [N, n_bins, T] and labels [N, 6, 19] of a dataset live in HBM
(2 KB + 114 B per item) and a batch is assembled by two kernel launches (gtc_patches + a label view kernel) from an
index tensor -- the sampler's permutation -- so the 313x data expansion to (3, 224, 224) fp32 happens on the GPU and
nothing crosses PCIe per step.  The loader objects support what the engines use: ``len()``, ``iter()``,
``next(iter())``, ``.dataset``, ``.batch_size`` (bestengine.py:440-445, :899-901; ViT_engine.py:277-296).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.GtcError("the dataloaders assemble batches on the GPU; no CUDA device is available (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _load_dir(directory: str, suffix: str):
    """(names, arrays) in the order sorted(os.listdir) gives the reference (my_dataloader.py:10-11 / ViT_dataloader.py:10-11).
    Packed per-clip files (audio_io.save_*_packed) are expanded into the items -- and the names -- the reference's
    per-segment files would have had, so the pairing by sorted position is unchanged."""
    from . import audio_io
    items = []
    for f in os.listdir(directory):
        if not f.endswith(suffix) or f.endswith(".idx.npy"):
            continue
        a = np.load(os.path.join(directory, f))
        if a.ndim == 3 and (f.endswith(audio_io.FEATURE_PACK_SUFFIX) or f.endswith(audio_io.LABEL_PACK_SUFFIX)):
            idx_path = os.path.join(directory, f[: -len(".npy")] + ".idx.npy")
            idx = np.load(idx_path) if os.path.exists(idx_path) else None
            items.extend(zip(audio_io.packed_item_names(f, len(a), idx), a))
        else:
            items.append((f, a))
    items.sort(key=lambda it: it[0])
    return [n for n, _ in items], [a for _, a in items]


def load_feature_dir(audio_dir: str, suffix: str = ".npy"):
    """sorted(listdir) order, as my_dataloader.py:10 / ViT_dataloader.py:10.  Returns (names, [N, n_bins, T] fp32)."""
    names, arrs = _load_dir(audio_dir, suffix)
    arrs = [a.astype(np.float32) for a in arrs]
    if arrs and any(a.shape != arrs[0].shape for a in arrs):
        raise ValueError("feature files of one dataset must share a shape (n_bins, T)")
    return names, (np.stack(arrs) if arrs else np.zeros((0, 96, 5), np.float32))


def load_png_dir(audio_dir: str, size=(224, 224), workers: int = 8):
    """The picture route of my_dataloader.py: every ``.png`` of ``audio_dir`` in sorted order (:10), decoded to RGB (:29) and
    resized by PIL exactly as ``transforms.Resize((224, 224))`` does for a PIL image (bilinear, PIL's own antialiasing;
    :18).  Decode and resize stay on the host -- they are the reference's own library calls -- and run ONCE per dataset
    (the reference repeats them every epoch); the result is [N, H, W, 3] uint8, 150 KB per item."""
    from concurrent.futures import ThreadPoolExecutor
    from PIL import Image
    names = sorted(f for f in os.listdir(audio_dir) if f.endswith('.png'))
    out = np.empty((len(names), size[0], size[1], 3), dtype=np.uint8)

    def one(i):
        with Image.open(os.path.join(audio_dir, names[i])) as im:
            out[i] = np.asarray(im.convert("RGB").resize((size[1], size[0]), Image.BILINEAR))

    with ThreadPoolExecutor(max_workers=max(1, workers)) as ex:
        list(ex.map(one, range(len(names))))
    return names, out


def load_label_dir(annotation_dir: str):
    names, arrs = _load_dir(annotation_dir, ".npy")
    for f, a in zip(names, arrs):
        if a.shape != (6, 19):
            print(f"Warning: Annotation has unexpected shape: {a.shape} ({f})")
    return names, (np.stack(arrs).astype(np.int8) if arrs else np.zeros((0, 6, 19), np.int8))


class DeviceTabDataset:
    """Features + labels resident on the GPU.  ``mode`` selects the tensor contract of __getitem__/batches."""

    def __init__(self, db: torch.Tensor, tabs: torch.Tensor, mode: int, img_size=(224, 224), label_kind: str = "argmax",
                 audio_files: Optional[List[str]] = None, annotation_files: Optional[List[str]] = None):
        assert db.shape[0] == tabs.shape[0], "Mismatch in audio and annotation file counts."
        self.db, self.tabs = db.contiguous(), tabs.contiguous()
        self.rgb = None                 # [N, H, W, 3] uint8 when the dataset was built from pictures (my_dataloader)
        self.mode, self.img_size, self.label_kind = mode, tuple(img_size), label_kind
        self.audio_files, self.annotation_files = audio_files or [], annotation_files or []

    def __len__(self):
        return self.db.shape[0]

    def batch(self, index: torch.Tensor):
        """index int64 [B] (device) -> (inputs [B,3,H,W] fp32, labels) assembled on the current stream."""
        if self.rgb is not None:                                   # picture route: ToTensor + Normalize of decoded PNGs
            x = ops.patches_rgb8(self.rgb, index=index)
        else:
            x = ops.patches(self.db, index=index, img_size=self.img_size, mode=self.mode)
        if self.label_kind == "argmax":
            y = ops.labels_argmax(self.tabs, index)                # (B, 6) int64       my_dataloader.py:40-44
        else:
            y = ops.labels_vit_heads(self.tabs, index)             # 6 x (B, 19) int64  ViT_dataloader.py:54
        return x, y

    def __getitem__(self, idx):
        idx = int(idx)
        if idx < 0:
            idx += len(self)
        if not 0 <= idx < len(self):
            raise IndexError(idx)
        x, y = self.batch(torch.tensor([idx], dtype=torch.int64, device=self.db.device))
        return (x[0], y[0]) if self.label_kind == "argmax" else (x[0], [h[0] for h in y])


class Subset:
    """What torch.utils.data.random_split returns, minus the Dataset base class."""

    def __init__(self, dataset: DeviceTabDataset, indices: Sequence[int]):
        self.dataset = dataset
        self.indices = list(int(i) for i in indices)

    def __len__(self):
        return len(self.indices)

    def __getitem__(self, i):
        return self.dataset[self.indices[i]]


def random_split(dataset, lengths, generator=None) -> List[Subset]:
    """torch.utils.data.random_split for integer lengths: one randperm, consecutive slices (same RNG consumption)."""
    assert sum(lengths) == len(dataset), "Sum of input lengths does not equal the length of the input dataset!"
    perm = torch.randperm(sum(lengths), generator=generator).tolist() if generator is not None else torch.randperm(sum(lengths)).tolist()
    out, at = [], 0
    for n in lengths:
        out.append(Subset(dataset, perm[at:at + n]))
        at += n
    return out


def sampler_order(n: int, shuffle: bool) -> torch.Tensor:
    """Item order of one epoch exactly as torch's DataLoader produces it, consuming the global CPU generator the same way
    (torch/utils/data/dataloader.py): creating the iterator draws the workers' base seed, then RandomSampler.__iter__ draws
    its own seed and permutes with a private generator.  Under ``torch.manual_seed(s)`` the batches therefore come in the
    order the reference's ``DataLoader(..., shuffle=True)`` (my_dataloader.py:62, ViT_dataloader.py:74) yields them."""
    torch.empty((), dtype=torch.int64).random_()                       # _BaseDataLoaderIter.__init__: base seed
    if not shuffle:
        return torch.arange(n, dtype=torch.int64)
    seed = int(torch.empty((), dtype=torch.int64).random_().item())    # RandomSampler.__iter__
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g)


class DeviceLoader:
    """Iterable of device batches over a Subset.  Replaces DataLoader(batch_size, shuffle, num_workers, pin_memory)."""

    def __init__(self, subset, batch_size: int, shuffle: bool, drop_last: bool = False):
        self.dataset = subset
        self.batch_size = int(batch_size)
        self.shuffle = bool(shuffle)
        self.drop_last = bool(drop_last)
        base = subset.dataset if isinstance(subset, Subset) else subset
        idx = subset.indices if isinstance(subset, Subset) else list(range(len(subset)))
        self._base = base
        self._index = torch.tensor(idx, dtype=torch.int64, device=base.db.device)

    def __len__(self):
        n = self._index.numel()
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = self._index.numel()
        order = self._index[sampler_order(n, self.shuffle).to(self._index.device)]
        for b in range(len(self)):
            idx = order[b * self.batch_size: (b + 1) * self.batch_size].contiguous()
            yield self._base.batch(idx)


def split_sizes(n: int, train_ratio: float, val_ratio: float):
    """my_dataloader.py:56-58 / ViT_dataloader.py:63-65."""
    train_size = int(train_ratio * n)
    val_size = int(val_ratio * n)
    return train_size, val_size, n - train_size - val_size
