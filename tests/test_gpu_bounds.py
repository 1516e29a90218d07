"""GPU: no kernel writes outside the buffers it is given.  compute-sanitizer is closed on the B200 pool, so every output
(and the caller-owned workspace, sized exactly as the library asks) is placed between two guard bands filled with a
sentinel and the bands are checked after the call -- ragged inputs, index gathers, tail tiles and padded rows included."""
import numpy as np
import pytest
import torch

from conftest import make_test_audio

pytestmark = pytest.mark.gpu
SR = 22050
GUARD = 4096          # bytes on each side (keeps 256-byte alignment of the payload)


class Guarded:
    def __init__(self, shape, dtype, dev):
        self.nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        pad = (-self.nbytes) % 256
        self.raw = torch.full((GUARD + self.nbytes + pad + GUARD,), 0xA5, dtype=torch.uint8, device=dev)
        self.t = self.raw[GUARD:GUARD + self.nbytes].view(dtype).view(shape)

    def intact(self):
        return bool((self.raw[:GUARD] == 0xA5).all()) and bool((self.raw[GUARD + self.nbytes:] == 0xA5).all())


def test_outputs_and_workspace_stay_inside_their_buffers(lib, recipe):
    from gtc_b200 import ops, synth, augment, _lib
    dev = torch.device("cuda")
    lens = np.array([SR * 2 + 37, 4000, SR + 777, 4410, SR * 3], dtype=np.int64)     # ragged, one clip without a window
    audio = torch.from_numpy(np.concatenate([make_test_audio(int(n), 70 + i) for i, n in enumerate(lens)])).to(dev)
    guards = []

    def g(shape, dtype):
        guards.append(Guarded(shape, dtype, dev))
        return guards[-1].t

    for engine in (_lib.GTC_GEMM_TCGEN05_FP16X2, _lib.GTC_GEMM_TCGEN05_3XTF32, _lib.GTC_GEMM_SIMT_FP32):
        plan = ops.CqtPlan(recipe, engine=engine)
        co, so = plan.offsets(lens)
        n_seg = int(so[-1])
        co_t, so_t = torch.from_numpy(co).to(dev), torch.from_numpy(so).to(dev)
        ws = g((plan.workspace_bytes(n_seg, len(lens)),), torch.uint8)                # exactly what the library asks for
        db = plan.segments_db(audio, co_t, so_t, n_seg, out=g((n_seg, 96, 5), torch.float32), workspace=ws)   # fused dB finish (default)
        if engine != _lib.GTC_GEMM_SIMT_FP32:
            plan.configure(_lib.GTC_OPT_FUSE_FINISH, 0)                                # separate finish_db_kernel
            plan.segments_db(audio, co_t, so_t, n_seg, out=g((n_seg, 96, 5), torch.float32), workspace=ws)
        torch.cuda.synchronize()
        plan.close()

    sp = ops.StructuredCqtPlan(recipe)
    starts = torch.tensor([0, 2205, int(lens[0]) + 5, int(lens[0]) + 100], dtype=torch.int64, device=dev)
    valid = torch.tensor([4410, 4410, 3000, 1234], dtype=torch.int32, device=dev)
    seglen = torch.tensor([4410, 4410, 4410, 3000], dtype=torch.int32, device=dev)
    sp.segments_db(audio, starts, valid, seglen, 4410, out=g((4, 96, sp.frames(4410)), torch.float32))
    sp.close()

    on, du, pi, eoff = synth.note_events([n / SR for n in lens], seed=3)
    t_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    times = np.concatenate([(np.arange(int(so[c + 1] - so[c])) + 0.5) * 0.1 for c in range(len(lens))])
    tabs, _ = ops.rasterize_tabs(t_(on), t_(du), t_(pi), t_(eoff), t_(times), t_(so), out=g((n_seg, 6, 19), torch.int8),
                                 stats=g((3,), torch.int64).zero_())
    idx = torch.tensor([5, 0, 3, 3, n_seg - 1], dtype=torch.int64, device=dev)

    for mode in (_lib.GTC_PATCH_VIT, _lib.GTC_PATCH_CNN):
        ops.patches(db, mode=mode, out=g((n_seg, 3, 224, 224), torch.float32))
        ops.patches(db, index=idx, mode=mode, out=g((5, 3, 224, 224), torch.float32))
    ops.patches(db, img_size=(37, 50), out=g((n_seg, 3, 37, 50), torch.float32))      # generic path, odd sizes
    rgb = torch.randint(0, 256, (7, 224, 224, 3), dtype=torch.uint8, device=dev)
    ops.patches_rgb8(rgb, index=torch.tensor([6, 1, 1], dtype=torch.int64, device=dev), out=g((3, 3, 224, 224), torch.float32))

    x = torch.rand((5, 3, 224, 224), device=dev) * 120 - 120
    augment.apply_ops(x, [1, 3, 4], shift=11, freq=(50, 5), time=(30, 10), out=g(tuple(x.shape), torch.float32))
    augment.apply_ops(x, [2, 4, 3], freq=(50, 5), time=(30, 10), noise_level=0.005, noise_seed=1, normalize_ref_db=-120.0,
                      out=g(tuple(x.shape), torch.float32))
    torch.cuda.synchronize()
    bad = [i for i, gd in enumerate(guards) if not gd.intact()]
    assert not bad, f"guard bands overwritten for buffers {bad} of {len(guards)}"


def test_inputs_are_not_read_beyond_their_ends(lib, recipe):
    """Read side: the inputs sit between guard bands of NaN (fp32 / fp64) or of an impossible value (int8 labels = 77).
    A read past either end that reaches a result shows up as a NaN or as a different result -- the operator GEMM would
    spread a single NaN sample over its whole segment."""
    from gtc_b200 import ops, synth, _lib
    dev = torch.device("cuda")
    lens = np.array([SR * 2 + 37, 4000, SR + 777, 4410], dtype=np.int64)
    a_np = np.concatenate([make_test_audio(int(n), 80 + i) for i, n in enumerate(lens)])

    def nan_guarded(arr, fill=float("nan")):
        arr = np.ascontiguousarray(arr)
        pad = 1024
        raw = torch.full((pad + arr.size + pad,), fill, dtype=torch.from_numpy(arr).dtype, device=dev)
        raw[pad:pad + arr.size] = torch.from_numpy(arr.reshape(-1)).to(dev)
        return raw[pad:pad + arr.size].view(arr.shape)

    for engine in (_lib.GTC_GEMM_TCGEN05_FP16X2, _lib.GTC_GEMM_SIMT_FP32):
        plan = ops.CqtPlan(recipe, engine=engine)
        co, so = plan.offsets(lens)
        n_seg = int(so[-1])
        co_t, so_t = torch.from_numpy(co).to(dev), torch.from_numpy(so).to(dev)
        want = plan.segments_db(torch.from_numpy(a_np).to(dev), co_t, so_t, n_seg).clone()
        got = plan.segments_db(nan_guarded(a_np), co_t, so_t, n_seg)
        assert bool(torch.isfinite(got).all()) and torch.equal(got, want)
        plan.close()
    sp = ops.StructuredCqtPlan(recipe)
    starts = torch.tensor([0, len(a_np) - 3000], dtype=torch.int64, device=dev)          # the last segment ends at the buffer end
    valid = torch.tensor([4410, 3000], dtype=torch.int32, device=dev)
    seglen = torch.tensor([4410, 4410], dtype=torch.int32, device=dev)
    s_want = sp.segments_db(torch.from_numpy(a_np).to(dev), starts, valid, seglen, 4410).clone()
    s_got = sp.segments_db(nan_guarded(a_np), starts, valid, seglen, 4410)
    assert bool(torch.isfinite(s_got).all()) and torch.equal(s_got, s_want)
    sp.close()

    db = want
    idx = torch.tensor([n_seg - 1, 0, n_seg - 1], dtype=torch.int64, device=dev)
    for mode in (_lib.GTC_PATCH_VIT, _lib.GTC_PATCH_CNN):
        p_want = ops.patches(db, index=idx, mode=mode)
        p_got = ops.patches(nan_guarded(db.cpu().numpy()), index=idx, mode=mode)
        assert bool(torch.isfinite(p_got).all()) and torch.equal(p_got, p_want)

    on, du, pi, eoff = synth.note_events([n / SR for n in lens], seed=3)
    t_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    times = np.concatenate([(np.arange(int(so[c + 1] - so[c])) + 0.5) * 0.1 for c in range(len(lens))])
    l_want, st_want = ops.rasterize_tabs(t_(on), t_(du), t_(pi), t_(eoff), t_(times), t_(so))
    l_got, st_got = ops.rasterize_tabs(nan_guarded(on), nan_guarded(du), nan_guarded(pi), t_(eoff), nan_guarded(times), t_(so))
    assert torch.equal(l_got, l_want) and torch.equal(st_got, st_want)
    tabs_g = nan_guarded(l_want.cpu().numpy(), fill=77)
    assert torch.equal(ops.labels_argmax(tabs_g, idx), ops.labels_argmax(l_want, idx))
    assert all(torch.equal(a, b) for a, b in zip(ops.labels_vit_heads(tabs_g, idx), ops.labels_vit_heads(l_want, idx)))
