"""Drop-in for the reference's ``ViT_dataloader.py``: same class / function names and signatures, batches assembled on
the GPU by libgtc (reference /root/reference/ViT_dataloader.py:8-88).

    GuitarTabDataset(audio_dir, annotation_dir, img_size=(224, 224))
    create_dataloaders(audio_dir, annotation_dir, batch_size=50, train_ratio=0.8, val_ratio=0.1, img_size=(224, 224))

Per item (reference :27-54): dB features -> (x+120)/120 -> clip [0,1] -> bicubic resize (align_corners=False) ->
3 identical channels; labels -> six int64 (19,) heads.  After collation a batch is ``(inputs (B,3,H,W) fp32, [6 x (B,19)
int64])`` -- exactly what ViT_engine.py:277-296 iterates over; here both already live on the GPU, so the engine's
``.to(device)`` is a no-op.  The split is the reference's: int(0.8 n) / int(0.1 n) / rest with
``torch.Generator().manual_seed(42)`` (:63-71), train shuffled, val/test not (:84-86).  ``num_workers``, ``pin_memory``
and ``prefetch_factor`` have no meaning here (no worker processes, no per-step H2D copy).
"""
import torch

from gtc_b200 import _lib, loaders


class GuitarTabDataset(loaders.DeviceTabDataset):
    def __init__(self, audio_dir, annotation_dir, img_size=(224, 224)):
        dev = loaders._device()
        audio_files, db = loaders.load_feature_dir(audio_dir, ".npy")
        annotation_files, tabs = loaders.load_label_dir(annotation_dir)
        assert len(audio_files) == len(annotation_files), "Mismatch in audio and annotation file counts."
        super().__init__(torch.from_numpy(db).to(dev), torch.from_numpy(tabs).to(dev), _lib.GTC_PATCH_VIT, img_size,
                         label_kind="heads", audio_files=audio_files, annotation_files=annotation_files)
        self.audio_dir, self.annotation_dir = audio_dir, annotation_dir

    @classmethod
    def from_tensors(cls, db, tabs, img_size=(224, 224)):
        """Index-aligned in-memory path: device features [N, n_bins, T] fp32 + labels [N, 6, 19] int8 (no files)."""
        self = cls.__new__(cls)
        loaders.DeviceTabDataset.__init__(self, db, tabs, _lib.GTC_PATCH_VIT, img_size, label_kind="heads")
        return self


def _make_loaders(dataset, batch_size, train_ratio, val_ratio, generator):
    train_size, val_size, test_size = loaders.split_sizes(len(dataset), train_ratio, val_ratio)
    train_dataset, val_dataset, test_dataset = loaders.random_split(dataset, [train_size, val_size, test_size], generator=generator)
    return (loaders.DeviceLoader(train_dataset, batch_size, shuffle=True),
            loaders.DeviceLoader(val_dataset, batch_size, shuffle=False),
            loaders.DeviceLoader(test_dataset, batch_size, shuffle=False))


def create_dataloaders(audio_dir, annotation_dir, batch_size=50, train_ratio=0.8, val_ratio=0.1, img_size=(224, 224)):
    dataset = GuitarTabDataset(audio_dir, annotation_dir, img_size)
    return _make_loaders(dataset, batch_size, train_ratio, val_ratio, torch.Generator().manual_seed(42))
