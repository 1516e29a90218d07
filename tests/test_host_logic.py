"""CPU tests of the host-side logic: WAV/.npy I/O, JAMS marshalling, splits, sharding (no CUDA calls)."""
import json
import os

import numpy as np
import pytest
import torch

from gtc_b200 import audio_io, events, loaders, ops, shard
from oracle import labels_oracle as lo


def test_wav_roundtrip_matches_librosa_semantics(tmp_path):
    rng = np.random.default_rng(0)
    y = rng.uniform(-0.9, 0.9, (1000, 2))
    q = np.clip(np.round(y * 32768), -32768, 32767).astype(np.int16)
    import scipy.io.wavfile
    scipy.io.wavfile.write(tmp_path / "a.wav", 22050, q)
    got, sr = audio_io.load_wav(tmp_path / "a.wav")
    assert sr == 22050 and got.dtype == np.float32 and got.shape == (1000,)
    want = (q.astype(np.float32) / 32768).mean(axis=1)               # soundfile scaling + librosa.to_mono
    assert np.abs(got - want).max() < 1e-7
    seg, _ = audio_io.load_wav(tmp_path / "a.wav", offset=0.01, duration=0.02)
    assert np.array_equal(seg, got[int(0.01 * 22050): int(0.01 * 22050) + int(0.02 * 22050)])
    assert audio_io.wav_duration(tmp_path / "a.wav") == 1000 / 22050


def test_npy_layouts(tmp_path):
    f = np.arange(96 * 5, dtype=np.float32).reshape(96, 5)
    audio_io.save_feature(tmp_path / "f.npy", f)
    raw = open(tmp_path / "f.npy", "rb").read()
    assert b"'fortran_order': True" in raw[:128] and b"'<f4'" in raw[:128] and len(raw) == 128 + 1920
    assert np.array_equal(np.load(tmp_path / "f.npy"), f)
    tab = np.zeros((6, 19), np.int8)
    tab[2, 3] = 1
    audio_io.save_label(tmp_path / "t.npy", tab)
    raw = open(tmp_path / "t.npy", "rb").read()
    assert len(raw) == 242 and b"'|i1'" in raw[:128] and b"'fortran_order': False" in raw[:128] and b"(6, 19)" in raw[:128]


def test_label_fixture_format_matches_reference_files():
    """tests/golden/labels_*.npy are verbatim copies of three reference label files (format pin only)."""
    gold = os.path.join(os.path.dirname(__file__), "golden")
    names = sorted(f for f in os.listdir(gold) if f.startswith("ref_label_"))
    assert len(names) >= 3
    for n in names:
        raw = open(os.path.join(gold, n), "rb").read()
        a = np.load(os.path.join(gold, n))
        assert a.shape == (6, 19) and a.dtype == np.int8 and set(np.unique(a)) <= {0, 1} and len(raw) == 242


def test_jams_json_reader_and_marshalling(tmp_path):
    doc = {"file_metadata": {"duration": 3.0},
           "annotations": [{"namespace": "note_midi", "data": [{"time": 0.5, "duration": 1.0, "value": 45.2, "confidence": None},
                                                                {"time": 1.0, "duration": 0.5, "value": {"pitch": 50}, "confidence": 1},
                                                                {"time": 1.0, "duration": 0.5, "value": {"other": 1}, "confidence": 1},
                                                                {"time": 1.0, "duration": 0.5, "value": "abc", "confidence": 1}]},
                           {"namespace": "pitch_contour", "data": {"time": [0.0, 0.01, 0.02], "duration": [0, 0, 0],
                                                                   "value": [{"frequency": 110.0, "voiced": True}, {"frequency": 0.0}, 220.0],
                                                                   "confidence": [0.9, 0.9, None]}}]}
    p = tmp_path / "x.jams"
    p.write_text(json.dumps(doc))
    jam = events.load_jams(p)
    on, du, pi = events.marshal_notes(jam)
    assert list(on) == [0.5, 1.0] and list(pi) == [45.2, 50.0]
    ct, cm, cc, ck = events.marshal_contours(jam)
    assert list(ct) == [0.0, 0.02] and list(ck) == [0, 1] and abs(cm[0] - 45.0) < 1e-12
    # the same objects drive the oracle (attribute-compatible with jams.JAMS)
    assert lo.extract_tablature_from_jams(jam, 0.6).sum() == 1


def test_random_split_reproduces_torch(tmp_path):
    class Fake:
        def __len__(self):
            return 103
    sizes = loaders.split_sizes(103, 0.8, 0.1)
    assert sizes == (82, 10, 11)
    mine = loaders.random_split(Fake(), list(sizes), generator=torch.Generator().manual_seed(42))
    ref = torch.utils.data.random_split(range(103), list(sizes), generator=torch.Generator().manual_seed(42))
    for a, b in zip(mine, ref):
        assert a.indices == list(b.indices)


def test_segment_counts():
    assert list(ops.segment_counts([661500, 4410, 4409, 0, 6615], 4410, 2205)) == [299, 1, 0, 0, 2]


def test_partitions_cover_every_clip_once():
    for world in (1, 2, 3, 8):
        parts = [shard.partition_round_robin(10, r, world) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(10))
        durs = [30.0, 14.6, 22.0, 29.5, 15.0, 16.0, 30.0, 21.0, 18.5, 25.0]
        parts = [shard.partition_balanced(durs, r, world) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(10))
        loads = [sum(durs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= 30.0


def test_merge_sharded_restores_clip_order():
    counts = np.array([3, 0, 2, 4, 1])
    full = np.arange(10 * 2).reshape(10, 2)
    off = np.concatenate([[0], np.cumsum(counts)])
    owners = [shard.partition_round_robin(5, r, 2) for r in range(2)]
    outs = [np.concatenate([full[off[c]:off[c + 1]] for c in own]) for own in owners]
    assert np.array_equal(shard.merge_sharded(outs, owners, counts), full)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    own = shard.partition_round_robin(7, rank, world)
    st = shard.ShardStats(n_clips=len(own), n_segments=int(own.sum()) + 1, n_samples=100 * (rank + 1), total=10 * (rank + 1),
                          with_notes=rank + 2, with_first_string=rank, n_skipped=0, elapsed_ns=1000 * (rank + 1))
    g = shard.gather_stats(st.as_tensor())
    q.put((rank, g.tolist(), shard.reduce_stats(g)))
    dist.destroy_process_group()


def test_stats_gather_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert res[0][1] == res[1][1]                                  # every rank sees the same [2, 8] table
    tot = res[0][2]
    assert tot["n_clips"] == 7 and tot["total"] == 30 and tot["with_notes"] == 5 and tot["elapsed_ns"] == 2000


# ---------------------------------------------------------------------------------------------------- packed files
def test_packed_files_explode_to_the_reference_layout(tmp_path):
    from gtc_b200 import audio_io
    rng = np.random.default_rng(0)
    feats = rng.uniform(-120, 0, (13, 96, 5)).astype(np.float32)
    tabs = (rng.random((13, 6, 19)) < 0.05).astype(np.int8)
    keep = np.array([0, 1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13])            # segment 4 had no picture (:307-309)
    fp = tmp_path / ("clipA" + audio_io.FEATURE_PACK_SUFFIX)
    lp = tmp_path / ("clipA" + audio_io.LABEL_PACK_SUFFIX)
    audio_io.save_features_packed(fp, feats)
    audio_io.save_labels_packed(lp, tabs, keep)
    assert audio_io.explode_features(fp, tmp_path / "f") == 13
    assert audio_io.explode_labels(lp, tmp_path / "l") == 13
    for k in range(13):
        a = np.load(tmp_path / "f" / f"clipA_segment_{k}.npy")
        assert a.dtype == np.float32 and a.shape == (96, 5) and a.flags.f_contiguous and np.array_equal(a, feats[k])
        b = np.load(tmp_path / "l" / "clipA" / f"clipA_{keep[k]:04d}.npy")
        assert b.dtype == np.int8 and b.shape == (6, 19) and np.array_equal(b, tabs[k])
    assert (tmp_path / "l" / "clipA" / "clipA_0000.npy").stat().st_size == 242          # the reference's 242-byte file


def test_loader_reads_packed_and_exploded_dirs_identically(tmp_path):
    """sorted(listdir) pairing order (un-padded counters sort _10 before _2) is kept for packed files."""
    from gtc_b200 import audio_io, loaders
    rng = np.random.default_rng(1)
    packed, flat = tmp_path / "packed", tmp_path / "flat"
    packed.mkdir(); flat.mkdir()
    for base, n in (("b_clip", 12), ("a_clip", 3)):
        feats = rng.uniform(-120, 0, (n, 96, 5)).astype(np.float32)
        p = packed / (base + audio_io.FEATURE_PACK_SUFFIX)
        audio_io.save_features_packed(p, feats)
        audio_io.explode_features(p, flat)
    n1, a1 = loaders.load_feature_dir(str(packed))
    n2, a2 = loaders.load_feature_dir(str(flat))
    assert n1 == n2 == sorted(n2) and np.array_equal(a1, a2)
    assert n1.index("b_clip_segment_10.npy") < n1.index("b_clip_segment_2.npy")


def test_inference_segment_tables():
    """tablature_generator.py:655-664 and "tablature-generator (1).py":299-323 window arithmetic."""
    from gtc_b200 import inference
    starts, valid = inference.segment_table(100000, 66150, 33075)
    assert starts.tolist() == [0, 33075, 66150, 99225] and valid.tolist() == [66150, 66150, 33850, 775]
    s, v, L = inference.vit_window_table(44100, 44100)
    assert L == 8820 and len(s) == 9 and s[-1] == 8 * 4410 and set(v.tolist()) == {8820}
    s, v, L = inference.vit_window_table(5000, 44100)                  # shorter than a window but >= half: one padded window
    assert s.tolist() == [0] and v.tolist() == [5000]
    s, v, L = inference.vit_window_table(4000, 44100)                  # shorter than half a window: skipped (:317-318)
    assert len(s) == 0


def test_chunk_planner_properties():
    """gtc_b200/chunks.py: chunks tile the clip list exactly, respect the segment limit, end on full GEMM waves where a
    choice exists, and ramp up from small chunks for the host-input path."""
    from gtc_b200 import chunks as ck
    sm, parts, n_out = 148, 2, 960
    eff = lambda n, c: ck.wave_efficiency(n, c, parts, n_out, sm)
    # BASELINE.json configs[1]: 360 clips of 299 segments
    nseg = np.full(360, 299)
    plain = ck.plan_bounds(nseg, 19200)
    assert plain[0] == (0, 64) and plain[-1][1] == 360 and all(a1 == b0 for (_, a1), (b0, _) in zip(plain, plain[1:]))
    wave = ck.plan_bounds(nseg, 19200, efficiency=eff)
    assert [c1 - c0 for c0, c1 in wave] == [63, 63, 63, 63, 63, 45]                       # 63 clips = 18 900 rows = 592 tiles
    assert ck.gemm_tiles(63 * 299, 63, parts, n_out) == 4 * sm and eff(63 * 299, 63) == 1.0
    assert abs(eff(54 * 299, 54) - 508 / 592) < 1e-12                                     # the old 54-clip chunks: 3.43 waves
    wave_rows = ck.rows_per_wave(n_out, sm)
    assert wave_rows == 148 * 128 // 4
    ramp = ck.plan_bounds(nseg, 19200, ramp=True, efficiency=eff, wave_rows=wave_rows)
    sizes = [c1 - c0 for c0, c1 in ramp]
    assert sizes[:7] == [15, 15, 31, 31, 47, 47, 63] and sum(sizes) == 360 and max(sizes) <= 64   # 1,1,2,2,3,3,4 full waves
    assert all(eff(299 * n, n) > 0.95 for n in sizes[:7])
    # the default limit is six waves: 94 clips per chunk, the same ramp in front
    six = ck.plan_bounds(nseg, 28400, efficiency=eff)
    assert [c1 - c0 for c0, c1 in six] == [94, 94, 94, 78] and eff(94 * 299, 94) > 0.99
    six_ramp = [c1 - c0 for c0, c1 in ck.plan_bounds(nseg, 28400, ramp=True, efficiency=eff, wave_rows=wave_rows)]
    assert six_ramp[:6] == [15, 15, 31, 31, 47, 47] and sum(six_ramp) == 360
    # ragged shard: zero-segment clips, one clip longer than the limit, random lengths
    rng = np.random.default_rng(5)
    nseg = rng.integers(0, 400, size=97)
    nseg[10] = 0
    nseg[40] = 5000
    for ramp_on in (False, True):
        for e in (None, eff):
            b = ck.plan_bounds(nseg, 2000, ramp=ramp_on, efficiency=e)
            assert b[0][0] == 0 and b[-1][1] == 97 and all(a1 == b0 for (_, a1), (b0, _) in zip(b, b[1:]))
            for c0, c1 in b:
                assert c1 > c0 and (nseg[c0:c1].sum() <= 2000 or c1 - c0 == 1)
    assert ck.plan_bounds([], 100) == [] and ck.plan_bounds([7], 100) == [(0, 1)]
    # tile widths follow cqt_gemm_tc.cu: 960 -> 240, 840*... ragged -> 256
    assert ck.gemm_tiles(128, 0, 1, 960) == 4 and ck.gemm_tiles(129, 0, 1, 960) == 8 and ck.gemm_tiles(128, 0, 1, 2 * 84 * 130) == -(-21840 // 240)


def test_loader_epoch_order_equals_torch_dataloader():
    """gtc_b200.loaders.sampler_order reproduces the item order (and the global-RNG consumption) of the reference's
    DataLoader(shuffle=True / False) under torch.manual_seed, over several epochs
    (my_dataloader.py:62-70, ViT_dataloader.py:74-86)."""
    import torch
    from torch.utils.data import DataLoader, Dataset
    from gtc_b200.loaders import sampler_order

    class Items(Dataset):
        def __len__(self):
            return 53

        def __getitem__(self, i):
            return i

    for workers in (0,):      # worker processes draw nothing more from the parent's generator (checked once with 2 workers;
        for shuffle in (True, False):   # not repeated here: forking a multi-threaded pytest process is a deadlock risk)
            torch.manual_seed(1234)
            dl = DataLoader(Items(), batch_size=8, shuffle=shuffle, num_workers=workers)
            ref = [torch.cat([b for b in dl]).tolist() for _ in range(3)]
            ref_next = torch.rand(1).item()                         # the global generator is left in the same state
            torch.manual_seed(1234)
            got = [sampler_order(53, shuffle).tolist() for _ in range(3)]
            assert got == ref and torch.rand(1).item() == ref_next
