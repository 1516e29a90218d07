"""CPU oracle for the hot path (TEST INFRASTRUCTURE ONLY -- never shipped, never measured as the product).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this package.  The product path (``guitar-tablature-classification_b200/``) must never import it.

PARITY STATUS: **parity unpinned**.  The arithmetic of the reference's hot path lives in third-party
libraries that are neither vendored under ``/root/reference`` nor installable here
(librosa 0.10.2/0.11.0 un-pinned, soxr/libsoxr 0.1.3, jams); the reference ships no tests or golden
vectors for this path.  The oracle therefore *restates* the published algorithms:

* ``cqt_oracle``     -- librosa.cqt/vqt + amplitude_to_db + the reference's ``cqt_lim``   (cqt.py:5-67, new_cqt.py:8-30)
* ``labels_oracle``  -- GuitarTablatureExtractor's label arithmetic                         (jam_to_tablature.py:55-178, 245-333)
* ``patches_oracle`` -- the two dataloaders' tensor contracts                              (ViT_dataloader.py:22-56, my_dataloader.py:26-50)

The one third-party function that *is* installed here, ``torch.nn.functional.interpolate`` (CPU), is used
directly to pin the bicubic patch restatement (tests/test_oracle_patches.py).
"""
