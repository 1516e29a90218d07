"""Drop-in for the reference's ``my_dataloader.py`` (the CNN / bestengine.py loader): same names and signatures,
batches assembled on the GPU by libgtc (reference /root/reference/my_dataloader.py:8-72).

    GuitarTabDataset(audio_dir, annotation_dir)
    create_dataloaders(audio_dir, annotation_dir, batch_size=64, train_ratio=0.8, val_ratio=0.1)

Tensor contract handed to bestengine.py:899-920: ``inputs (B,3,224,224) fp32`` ImageNet-normalised with the reference's
mean/std (:20) and ``labels (B,6) int64`` = argmax over the 19 frets of each string (first 1 wins, all-zero row -> 0,
:40-44).  Two kinds of ``audio_dir``:
* ``.png`` pictures, the reference's own input (:10,:29): decoded and resized by PIL on the host exactly as the reference
  does (once per dataset instead of once per item and epoch), ToTensor + Normalize per batch on the GPU -- the batches
  are bit-identical to the reference's (tests/test_gpu_dropins.py compares with torchvision's own transform).
* (96, T) dB feature ``.npy`` files (what this repo's cqt.py writes): rendering the pictures is matplotlib's business and
  not a numeric contract, so here the "picture" is the grey image clip((dB+120)/120, 0, 1) with the
  highest bin on the top row, resized bilinearly to 224x224 and replicated to 3 channels.
The split is unseeded like the reference's (:60).
"""
import os

import torch

from gtc_b200 import _lib, loaders


class GuitarTabDataset(loaders.DeviceTabDataset):
    def __init__(self, audio_dir, annotation_dir):
        dev = loaders._device()
        annotation_files, tabs = loaders.load_label_dir(annotation_dir)
        if any(f.endswith('.png') for f in os.listdir(audio_dir)):
            # the reference's own input (my_dataloader.py:10): pictures rendered by new_cqt.py.  PIL decode + resize on
            # the host once, then ToTensor + Normalize per batch on the GPU -- bit-identical to the reference's transform
            audio_files, rgb = loaders.load_png_dir(audio_dir)
            assert len(audio_files) == len(annotation_files), "Mismatch in audio and annotation file counts."
            super().__init__(torch.zeros((len(audio_files), 1, 1), device=dev), torch.from_numpy(tabs).to(dev), _lib.GTC_PATCH_CNN,
                             (224, 224), label_kind="argmax", audio_files=audio_files, annotation_files=annotation_files)
            self.rgb = torch.from_numpy(rgb).to(dev)
        else:
            audio_files, db = loaders.load_feature_dir(audio_dir, ".npy")
            assert len(audio_files) == len(annotation_files), "Mismatch in audio and annotation file counts."
            super().__init__(torch.from_numpy(db).to(dev), torch.from_numpy(tabs).to(dev), _lib.GTC_PATCH_CNN, (224, 224),
                             label_kind="argmax", audio_files=audio_files, annotation_files=annotation_files)
        self.audio_dir, self.annotation_dir = audio_dir, annotation_dir

    @classmethod
    def from_tensors(cls, db, tabs):
        """Index-aligned in-memory path: device features [N, n_bins, T] fp32 + labels [N, 6, 19] int8 (no files)."""
        self = cls.__new__(cls)
        loaders.DeviceTabDataset.__init__(self, db, tabs, _lib.GTC_PATCH_CNN, (224, 224), label_kind="argmax")
        return self


def create_dataloaders(audio_dir, annotation_dir, batch_size=64, train_ratio=0.8, val_ratio=0.1):
    dataset = GuitarTabDataset(audio_dir, annotation_dir)
    train_size, val_size, test_size = loaders.split_sizes(len(dataset), train_ratio, val_ratio)
    train_dataset, val_dataset, test_dataset = loaders.random_split(dataset, [train_size, val_size, test_size])
    return (loaders.DeviceLoader(train_dataset, batch_size, shuffle=True),
            loaders.DeviceLoader(val_dataset, batch_size, shuffle=False),
            loaders.DeviceLoader(test_dataset, batch_size, shuffle=False))
