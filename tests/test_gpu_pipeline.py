"""GPU: the chunked three-stream FrontEnd gives the same results as single calls, for host and device inputs."""
import numpy as np
import pytest
import torch

from conftest import make_test_audio
from oracle import labels_oracle as lo, patches_oracle as po

pytestmark = pytest.mark.gpu
SR = 22050


@pytest.fixture(scope="module")
def shard_inputs():
    from gtc_b200 import synth
    lens = np.array([SR * 3, 4000, SR * 2 + 123, SR * 5, 4410, SR * 4], dtype=np.int64)
    audio = np.concatenate([make_test_audio(int(n), 50 + i) for i, n in enumerate(lens)])
    on, du, pi, eoff = synth.note_events([n / SR for n in lens], seed=7)
    return audio, lens, np.stack([on, du, pi]), eoff


def test_frontend_matches_direct_ops(lib, shard_inputs, recipe):
    from gtc_b200 import ops
    from gtc_b200.pipeline import FrontEnd, ShardInputs
    audio, lens, ev, eoff = shard_inputs
    dev = torch.device("cuda")
    fe = FrontEnd(recipe, chunk_segments=40, patch_batch=16)           # forces several chunks and patch batches
    host = ShardInputs(torch.from_numpy(audio).pin_memory(), lens, torch.from_numpy(ev).pin_memory(), eoff, sr=SR)
    seen = []

    def consumer(patches, tabs, g0):
        seen.append((g0, patches[:1].clone(), tabs[:1].clone()))

    out = fe.run(host, consumer=consumer)
    torch.cuda.synchronize()
    stats = fe.stats()
    db_e2e, tabs_e2e = out.db.numpy().copy(), out.tabs.numpy().copy()
    assert out.h2d_bytes >= audio.nbytes and out.d2h_bytes >= db_e2e.nbytes

    # direct single-call path
    plan = ops.CqtPlan(recipe)
    clip_off, seg_off = plan.offsets(lens)
    n_seg = int(seg_off[-1])
    assert out.n_seg == n_seg and len(fe.plan_chunks(host)) > 2
    db = plan.segments_db(torch.from_numpy(audio).to(dev), torch.from_numpy(clip_off).to(dev), torch.from_numpy(seg_off).to(dev), n_seg)
    assert np.array_equal(db.cpu().numpy(), db_e2e)                    # chunking must not change a single bit

    # labels against the oracle (times = (i + 0.5) * duration / n_seg per clip)
    want = []
    for c, n in enumerate(lens):
        k = int(seg_off[c + 1] - seg_off[c])
        if k:
            t = lo.segment_times(float(n) / SR, k)
            want.append(lo.rasterize_events_numpy(ev[0, eoff[c]:eoff[c + 1]], ev[1, eoff[c]:eoff[c + 1]], ev[2, eoff[c]:eoff[c + 1]], t))
    want = np.concatenate(want)
    assert np.array_equal(tabs_e2e, want)
    assert list(stats) == [n_seg, int((want.sum(axis=(1, 2)) > 0).sum()), int((want[:, 0].sum(axis=1) > 0).sum())]

    # device-resident inputs give identical outputs
    devin = ShardInputs(torch.from_numpy(audio).to(dev), lens, torch.from_numpy(ev).to(dev), eoff, sr=SR)
    out2 = fe.run(devin, device_inputs=True)
    torch.cuda.synchronize()
    assert np.array_equal(out2.db.cpu().numpy(), db_e2e) and np.array_equal(out2.tabs.cpu().numpy(), tabs_e2e)

    # patch batches handed to the consumer are the ViT patches of the right segments
    assert len(seen) >= 3
    for g0, p, t in seen[:4]:
        assert np.abs(p[0].cpu().numpy() - po.vit_patch(db_e2e[g0])).max() < 3e-5
        assert np.array_equal(t[0].cpu().numpy(), want[g0])


def test_sharded_run_equals_single_run(lib, shard_inputs, recipe):
    """shard the clip list N ways, concatenate, compare bit-for-bit with the 1-way result;
    the gathered stats equal the serial sums (ranks emulated one after the other on the one GPU of the test box)."""
    from gtc_b200 import ops, shard
    from gtc_b200.pipeline import FrontEnd, ShardInputs
    audio, lens, ev, eoff = shard_inputs
    dev = torch.device("cuda")
    fe = FrontEnd(recipe, chunk_segments=64, patch_batch=32)
    clip_off = np.concatenate([[0], np.cumsum(lens)])

    def run(clip_ids):
        a = np.concatenate([audio[clip_off[c]:clip_off[c + 1]] for c in clip_ids])
        e = np.concatenate([ev[:, eoff[c]:eoff[c + 1]] for c in clip_ids], axis=1)
        eo = np.concatenate([[0], np.cumsum([eoff[c + 1] - eoff[c] for c in clip_ids])]).astype(np.int64)
        inp = ShardInputs(torch.from_numpy(a).to(dev), lens[clip_ids], torch.from_numpy(np.ascontiguousarray(e)).to(dev), eo, sr=SR)
        out = fe.run(inp, device_inputs=True, emit_patches=False)
        torch.cuda.synchronize()
        return out.db.cpu().numpy().copy(), out.tabs.cpu().numpy().copy(), fe.stats().copy()

    full_db, full_tabs, full_stats = run(np.arange(len(lens)))
    counts = ops.segment_counts(lens, 4410, 2205)
    for world in (2, 3):
        owners = [shard.partition_round_robin(len(lens), r, world) for r in range(world)]
        res = [run(o) for o in owners]
        assert np.array_equal(shard.merge_sharded([r[0] for r in res], owners, counts), full_db)
        assert np.array_equal(shard.merge_sharded([r[1] for r in res], owners, counts), full_tabs)
        gathered = torch.stack([shard.ShardStats(n_clips=len(o), n_segments=int(counts[o].sum()), total=int(r[2][0]),
                                                 with_notes=int(r[2][1]), with_first_string=int(r[2][2])).as_tensor()
                                for o, r in zip(owners, res)])
        tot = shard.reduce_stats(gathered)
        assert [tot["total"], tot["with_notes"], tot["with_first_string"]] == list(full_stats)
        assert tot["n_clips"] == len(lens) and tot["n_segments"] == int(counts.sum())


def test_host_staging_modes_agree(lib, shard_inputs, recipe):
    """Host inputs: whole-shard staging (one train of piece copies) and the per-chunk double-buffer fallback used for
    shards larger than ``stage_bytes_limit`` give the same bits; int16 PCM input equals its fp32 conversion; a second
    run over new inputs through the same FrontEnd (pinned staging reuse) is not disturbed by the first."""
    from gtc_b200.pipeline import FrontEnd, ShardInputs
    audio, lens, ev, eoff = shard_inputs
    pcm = np.clip(np.round(audio * 32768.0), -32768, 32767).astype(np.int16)
    f32 = pcm.astype(np.float32) / 32768.0
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    res = {}
    for name, limit, piece, a in (("staged", 4 << 30, 2, pcm), ("fallback", 0, 8, pcm), ("staged_f32", 4 << 30, 8, f32)):
        fe = FrontEnd(recipe, chunk_segments=40, patch_batch=16)
        fe.stage_bytes_limit, fe.stage_piece_clips = limit, piece
        out = fe.run(ShardInputs(pin(a), lens, pin(ev), eoff, sr=SR))
        torch.cuda.synchronize()
        res[name] = (out.db.numpy().copy(), out.tabs.numpy().copy(), fe.stats().copy(), out.h2d_bytes)
        # same FrontEnd, different shard (clips reversed): results must be those of a fresh FrontEnd
        order = np.arange(len(lens))[::-1]
        off = np.concatenate([[0], np.cumsum(lens)])
        a2 = np.concatenate([a[off[c]:off[c + 1]] for c in order])
        e2 = np.concatenate([ev[:, eoff[c]:eoff[c + 1]] for c in order], axis=1)
        eo2 = np.concatenate([[0], np.cumsum([eoff[c + 1] - eoff[c] for c in order])]).astype(np.int64)
        out2 = fe.run(ShardInputs(pin(a2), lens[order], pin(e2), eo2, sr=SR))
        torch.cuda.synchronize()
        res[name + "_rev"] = (out2.db.numpy().copy(), out2.tabs.numpy().copy())
    for k in ("fallback", "staged_f32"):
        assert np.array_equal(res[k][0], res["staged"][0]) and np.array_equal(res[k][1], res["staged"][1])
        assert list(res[k][2]) == list(res["staged"][2])
        assert np.array_equal(res[k + "_rev"][0], res["staged_rev"][0]) and np.array_equal(res[k + "_rev"][1], res["staged_rev"][1])
    assert res["staged"][3] == res["fallback"][3] and res["staged_f32"][3] > res["staged"][3]
    # reversed-clip run == per-clip blocks of the forward run, in reverse order
    from gtc_b200 import ops
    counts = ops.segment_counts(lens, 4410, 2205)
    so = np.concatenate([[0], np.cumsum(counts)])
    want = np.concatenate([res["staged"][0][so[c]:so[c + 1]] for c in np.arange(len(lens))[::-1]])
    assert np.array_equal(res["staged_rev"][0], want)


def test_full_size_shard_spot_checks(lib, recipe, basis_cache):
    """BASELINE.json configs[1] at full size (360 clips x 30 s, 107 640 segments, the chunks bench.py times): the
    pipeline's dB features, labels and patches of 48 randomly chosen segments against the CPU oracle, plus whole-run
    invariants (value set, 0 dB peak per segment, identical channels, label stats)."""
    from gtc_b200 import synth
    from gtc_b200.pipeline import FrontEnd, ShardInputs
    from oracle import cqt_oracle as co
    dev = torch.device("cuda")
    n_clips, n = 360, SR * 30
    audio = synth.pluck_clips(n_clips, n, sr=SR, seed=1, device=dev, block=24).reshape(-1)
    on, du, pi, eoff = synth.note_events([30.0] * n_clips, seed=2)
    ev = np.stack([on, du, pi])
    lens = np.full(n_clips, n, dtype=np.int64)
    fe = FrontEnd(recipe)
    n_seg = n_clips * 299
    rng = np.random.default_rng(11)
    picks = np.sort(rng.choice(n_seg, size=48, replace=False))
    grabbed = {}
    chan_diff = []

    def consumer(patches, tabs, g0):
        sel = picks[(picks >= g0) & (picks < g0 + patches.shape[0])]
        for g in sel:
            grabbed[int(g)] = patches[int(g) - g0].clone()
        chan_diff.append(((patches[:, 0] != patches[:, 1]) | (patches[:, 0] != patches[:, 2])).any())

    inp = ShardInputs(audio, lens, torch.from_numpy(ev).to(dev), eoff, sr=SR)
    out = fe.run(inp, device_inputs=True, consumer=consumer)
    torch.cuda.synchronize()
    assert out.n_seg == n_seg and len(fe.plan_chunks(inp)) == 4 and len(grabbed) == 48
    assert not any(bool(c) for c in chan_diff)                      # ViT_dataloader.py:50 repeat(3, 1, 1)
    db = out.db
    assert bool(((db == -120) | ((db >= -60) & (db <= 0))).all())  # cqt_lim value set
    assert bool((db.amax(dim=(1, 2)) == 0).all())                   # ref=np.amax: every segment peaks at exactly 0 dB
    tabs = out.tabs.cpu().numpy()
    stats = fe.stats()
    assert list(stats) == [n_seg, int((tabs.sum(axis=(1, 2)) > 0).sum()), int((tabs[:, 0].sum(axis=1) > 0).sum())]
    db_h = db.cpu().numpy()
    audio_h = audio.cpu().numpy()
    tol = 0.01
    for g in picks:
        c, i = divmod(int(g), 299)
        seg = audio_h[c * n + i * 2205: c * n + i * 2205 + 4410]
        _, pre, _ = co.segment_features(seg, SR, fmin=co.note_to_hz_C(1), _basis_cache=basis_cache, return_pre_cut=True)
        above, below = pre > -60 + 2 * tol, pre < -60 - 2 * tol
        assert np.abs(db_h[g][above] - pre[above]).max() < tol and (db_h[g][below] == -120).all(), f"segment {g}"
        t = lo.segment_times(30.0, 299)[i:i + 1]
        want = lo.rasterize_events_numpy(ev[0, eoff[c]:eoff[c + 1]], ev[1, eoff[c]:eoff[c + 1]], ev[2, eoff[c]:eoff[c + 1]], t)[0]
        assert np.array_equal(tabs[g], want), f"labels of segment {g}"
        assert np.abs(grabbed[int(g)][0].cpu().numpy() - po.vit_patch(db_h[g])).max() < 3e-5, f"patch of segment {g}"


def test_prefetch_of_the_next_shard(lib, shard_inputs, recipe):
    """run(A, next_inp=B) then run(B): B's host->device copies are queued behind A's on the un-joined staging stream;
    results equal those of independent runs, also when the prefetched shard is not the one that is run next and when
    the same pinned buffers are prefetched again and again (what bench.py's e2e arm does)."""
    from gtc_b200.pipeline import FrontEnd, ShardInputs
    audio, lens, ev, eoff = shard_inputs
    pcm = np.clip(np.round(audio * 32768.0), -32768, 32767).astype(np.int16)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    off = np.concatenate([[0], np.cumsum(lens)])

    def shard(order):
        a = np.concatenate([pcm[off[c]:off[c + 1]] for c in order])
        e = np.concatenate([ev[:, eoff[c]:eoff[c + 1]] for c in order], axis=1)
        eo = np.concatenate([[0], np.cumsum([eoff[c + 1] - eoff[c] for c in order])]).astype(np.int64)
        return ShardInputs(pin(a), lens[order], pin(e), eo, sr=SR)

    orders = [np.arange(len(lens)), np.arange(len(lens))[::-1], np.array([2, 0, 5, 3, 1, 4])]
    shards = [shard(o) for o in orders]

    def result(fe, inp, **kw):
        out = fe.run(inp, **kw)
        torch.cuda.synchronize()
        return out.db.numpy().copy(), out.tabs.numpy().copy(), fe.stats().copy()

    want = [result(FrontEnd(recipe, chunk_segments=40, patch_batch=16), s) for s in shards]
    fe = FrontEnd(recipe, chunk_segments=40, patch_batch=16)
    fe.stage_piece_clips = 2
    got = [result(fe, shards[0], next_inp=shards[1]), result(fe, shards[1], next_inp=shards[2]),
           result(fe, shards[0], next_inp=shards[0]),          # shard 2 was prefetched but is skipped
           result(fe, shards[0], next_inp=shards[0]), result(fe, shards[0]), result(fe, shards[2])]
    for g, w in zip(got, [want[0], want[1], want[0], want[0], want[0], want[2]]):
        assert np.array_equal(g[0], w[0]) and np.array_equal(g[1], w[1]) and list(g[2]) == list(w[2])
