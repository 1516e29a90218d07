"""GPU parity of the inference-time front-ends against the oracle:
tablature_generator.py:599-666 (3 s segments, C2/84 bins/hop 512, |C| -> dB ref=max) and
"tablature-generator (1).py":282-372 (44.1 kHz 0.2 s windows, cqt.py recipe, (x+120)/120, bicubic 224)."""
import os

import numpy as np
import pytest
import torch

from conftest import make_test_audio
from oracle import cqt_oracle as co

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fe(lib):
    from gtc_b200.inference import TabCnnFrontEnd
    return TabCnnFrontEnd()


def test_halve_rate_is_the_soxr_hq_stage(fe):
    y = make_test_audio(44100 + 123, seed=61, sr=44100.0)
    t = torch.from_numpy(y).cuda()
    ref = co.resample_2to1(y)                                       # scale=True: x sqrt(2)
    got = fe.plan.halve_rate(t, scale=True).cpu().numpy()
    assert got.shape == ref.shape and np.abs(got - ref).max() < 2e-6 * np.abs(ref).max()
    got = fe.plan.halve_rate(t, scale=False).cpu().numpy()          # librosa.load(sr=22050) of a 44.1 kHz file
    assert np.abs(got - ref / np.sqrt(2.0)).max() < 2e-6 * np.abs(ref).max()


def test_load_resamples_44k_files(fe, tmp_path):
    from gtc_b200 import audio_io
    y = make_test_audio(44100, seed=62, sr=44100.0)
    audio_io.write_wav_pcm16(tmp_path / "a.wav", y, 44100)
    y44, _ = audio_io.load_wav(tmp_path / "a.wav")
    got = fe.load(str(tmp_path / "a.wav"))
    ref = co.resample_2to1(y44) / np.sqrt(2.0)
    assert got.shape == (22050,) and np.abs(got - ref).max() < 2e-6
    audio_io.write_wav_pcm16(tmp_path / "b.wav", y[:8000], 16000)
    with pytest.raises(ValueError):
        fe.load(str(tmp_path / "b.wav"))


def test_three_second_segments_match_oracle(fe):
    y = make_test_audio(int(22050 * 5.2), seed=63)
    db, times = fe.cqt_db_segments(y)                               # PCM_16 temp-file hop included (:878-882)
    db = db.cpu().numpy()
    assert db.shape == (4, 84, 130) and np.allclose(times, [0.0, 1.5, 3.0, 4.5])
    cache = {}
    yq = (np.clip(np.rint(y.astype(np.float64) * 32767.0), -32768, 32767).astype(np.int16).astype(np.float32) / np.float32(32768.0))
    for i, t0 in enumerate(times):
        s = int(round(t0 * 22050))
        seg = yq[s: s + 66150]
        seg = np.pad(seg, (0, 66150 - len(seg)))
        C = co.cqt(seg, sr=22050, hop_length=512, fmin=co.note_to_hz_C(2), n_bins=84, _basis_cache=cache)
        ref = co.amplitude_to_db_amax(np.abs(C))
        big = np.abs(C) > 3e-3 * np.abs(C).max()
        assert np.abs(db[i] - ref)[big].max() <= 0.01
        assert db[i].max() == 0.0 and db[i].min() >= -80.0


def test_drop_in_segment_audio_and_whole_file_features(lib, tmp_path):
    import tablature_generator as tg
    from gtc_b200 import audio_io
    y = make_test_audio(22050 * 4, seed=64)
    audio_io.write_wav_pcm16(tmp_path / "song.wav", y, 22050)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        gen = tg.TablatureImageGenerator(model_path=None)
        segments, sr = gen.segment_audio(str(tmp_path / "song.wav"))
        assert sr == 22050 and [round(t, 3) for _, t in segments] == [0.0, 1.5, 3.0]
        assert all(len(s) == 66150 and s.dtype == np.float32 for s, _ in segments)
        yl, _ = audio_io.load_wav(tmp_path / "song.wav")
        assert np.array_equal(segments[1][0][: len(yl) - 33075], yl[33075:]) and np.all(segments[2][0][len(yl) - 66150:] == 0)
        path = gen.audio_to_cqt_image(str(tmp_path / "song.wav"))
        got = np.load(path)
        C = co.cqt(yl, sr=22050, hop_length=512, fmin=co.note_to_hz_C(2), n_bins=84)
        ref = co.amplitude_to_db_amax(np.abs(C))
        big = np.abs(C) > 3e-3 * np.abs(C).max()
        assert got.shape == ref.shape == (84, 1 + len(yl) // 512) and np.abs(got - ref)[big].max() <= 0.01
    finally:
        os.chdir(cwd)


def test_vit_preprocess_and_prepare(lib):
    from gtc_b200 import inference
    y = make_test_audio(44100, seed=65, sr=44100.0)
    norm, stamps = inference.vit_preprocess(y, sr=44100)
    assert norm.shape == (9, 96, 9) and np.allclose(stamps, np.arange(9) * 0.1)
    got = norm.cpu().numpy()
    for i in range(9):
        seg = y[i * 4410: i * 4410 + 8820]
        cut, pre, _ = co.segment_features(seg, 44100, fmin=co.note_to_hz_C(1), return_pre_cut=True)
        ref = np.clip((cut + 120.0) / 120.0, 0, 1)
        away = np.abs(pre + 60.0) > 0.02
        assert np.abs(got[i] - ref)[away].max() <= 0.01 / 120.0 + 1e-6
    img = inference.prepare_for_vit(norm)
    ref_img = torch.nn.functional.interpolate(norm.cpu().unsqueeze(1), size=(224, 224), mode="bicubic", align_corners=False).repeat(1, 3, 1, 1)
    assert img.shape == (9, 3, 224, 224) and (img.cpu() - ref_img).abs().max().item() < 3e-5
    short, stamps = inference.vit_preprocess(y[:5000], sr=44100)      # one zero-padded window (:321-323)
    assert short.shape == (1, 96, 9) and stamps == [0.0]
    cut = co.segment_features(np.pad(y[:5000], (0, 3820)), 44100, fmin=co.note_to_hz_C(1))
    assert np.abs(short[0].cpu().numpy() - np.clip((cut + 120.0) / 120.0, 0, 1)).mean() < 1e-3
