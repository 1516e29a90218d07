"""BASELINE.json configs[3] and [4] as streams: sharded corpus -> CNN batches (128) / ViT batches (50), consumed like the
engines do (bestengine.py:899-920 indexes labels[:, i]; ViT_engine.py:277-296 iterates six heads and argmaxes)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_vit_stream_of_a_sharded_corpus(lib):
    from gtc_b200 import streams, ops
    seen = {"n": 0, "items": 0}

    def consumer(x, labels):
        assert x.shape[1:] == (3, 224, 224) and x.dtype == torch.float32 and x.is_cuda
        assert len(labels) == 6 and all(h.shape == (x.shape[0], 19) and h.dtype == torch.int64 for h in labels)
        idx = [torch.argmax(h, dim=1) for h in labels]                      # ViT_engine.py:291-293
        assert all(i.shape == (x.shape[0],) for i in idx)
        seen["n"] += 1
        seen["items"] += x.shape[0]

    # rank 3 of 8 of an 800-clip corpus (100 clips x 10 s): clip ids 3, 11, ...
    rep = streams.stream_corpus(800, 10.0, rank=3, world_size=8, batch_size=50, mode="vit", consumer=consumer, clips_per_block=64)
    per_clip = (int(22050 * 10.0) - 4410) // 2205 + 1
    assert rep.n_clips == 100 and rep.n_segments == 100 * per_clip == seen["items"]
    assert rep.n_batches == seen["n"] and rep.n_full_batches >= rep.n_batches - 4
    assert rep.label_stats[0] == rep.n_segments and 0 < rep.label_stats[1] <= rep.n_segments
    assert rep.seconds_of_audio == pytest.approx(1000.0)


def test_cnn_stream_feeds_a_six_head_step(lib):
    """One optimisation step of a small six-head CNN on every batch: the tensor contract of bestengine.py:899-920."""
    from gtc_b200 import streams
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 7, stride=4), torch.nn.ReLU(), torch.nn.AdaptiveAvgPool2d(4),
                              torch.nn.Flatten(), torch.nn.Linear(128, 6 * 19)).cuda()
    opt = torch.optim.SGD(net.parameters(), lr=0.01)
    losses = []

    def consumer(inputs, labels):
        assert inputs.shape[1:] == (3, 224, 224) and labels.shape == (inputs.shape[0], 6) and labels.dtype == torch.int64
        assert int(labels.min()) >= 0 and int(labels.max()) < 19
        out = net(inputs).view(-1, 6, 19)
        loss = sum(torch.nn.functional.cross_entropy(out[:, i], labels[:, i]) for i in range(6))      # labels[:, i] (:905-911)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))

    rep = streams.stream_corpus(16, 10.0, batch_size=128, mode="cnn", consumer=consumer, clips_per_block=16)
    assert rep.n_batches == len(losses) == -(-rep.n_segments // 128) and np.isfinite(losses).all()
    assert losses[-1] < losses[0]
