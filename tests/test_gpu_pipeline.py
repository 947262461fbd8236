"""GPU tests of the device-side input pipeline (tacorl_b200/csrc/data_pipeline.cu) against the
CPU restatement of the reference's transforms / window collation (oracle/transforms_oracle.py)."""
import pytest
import torch

from oracle import transforms_oracle as TO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_window_gather_pad_and_random_shift_match_reference_semantics():
    from tacorl_b200.utils import transforms as T
    g = torch.Generator().manual_seed(0)
    F_, H, pad, Tn = 40, 32, 3, 8
    store = torch.randint(0, 256, (F_, 3, H, H), generator=g, dtype=torch.uint8)
    start = torch.tensor([0, 5, 30, 17])
    window = torch.tensor([8, 5, 8, 4])
    shift = torch.randint(0, 2 * pad + 1, (4, Tn, 2), generator=g)
    out = T.window_gather(store.to(DEV), start, window, Tn, shift, pad).cpu()
    assert out.shape == (4, Tn, 3, H, H) and out.dtype == torch.uint8
    for b in range(4):
        frames = TO.pad_frames(store[int(start[b]):int(start[b]) + int(window[b])], int(window[b]), Tn)
        want = TO.random_shifts(frames.float(), pad, shift[b].view(Tn, 1, 1, 2))
        # grid_sample lands on exact pixel centres up to float rounding of the grid: <= 1e-3 on the 0..255 scale
        assert (out[b].float() - want).abs().max() < 5e-3, (b, float((out[b].float() - want).abs().max()))
    plain = T.window_gather(store.to(DEV), start, window, Tn).cpu()
    for b in range(4):
        assert torch.equal(plain[b], TO.pad_frames(store[int(start[b]):int(start[b]) + int(window[b])], int(window[b]), Tn))


def test_actions_gather_pad_matches_pad_sequence():
    from tacorl_b200.utils import transforms as T
    g = torch.Generator().manual_seed(1)
    store = torch.rand(50, 7, generator=g) * 2 - 1
    start, window = torch.tensor([3, 20, 41]), torch.tensor([16, 9, 1])
    out = T.actions_gather_pad(store.to(DEV), start, window, 16).cpu()
    for b in range(3):
        s, w = int(start[b]), int(window[b])
        assert torch.equal(out[b], TO.pad_rel_actions(store[s:s + w], w, 16))
    rep = T.actions_gather_pad(store.to(DEV), start, window, 16, zero_pad=False).cpu()
    assert torch.equal(rep[1, 12], store[28])


@pytest.mark.parametrize("ops", [[0, 1, 3], [3, 0, 1], [1, 3, 0], [1], [3], []])
def test_color_jitter_matches_torchvision_ops(ops):
    from tacorl_b200.utils import transforms as T
    g = torch.Generator().manual_seed(2 + len(ops))
    N, H, W = 5, 24, 40
    x = torch.randint(0, 256, (N, 3, H, W), generator=g, dtype=torch.uint8)
    x[0, :, :4] = 128                       # grey patch: the max == min branch of rgb -> hsv
    fac = torch.stack([torch.rand(N, generator=g) * 0.2 + 0.9, torch.rand(N, generator=g) * 0.2 + 0.9,
                       torch.rand(N, generator=g) * 0.04 - 0.02], dim=1)
    order = torch.full((N,), T.pack_order(ops), dtype=torch.int32)
    out = T.color_jitter(x.to(DEV), order, fac).cpu()
    for n in range(N):
        want = TO.normalize(TO.color_jitter(TO.scale_image(x[n]), ops, float(fac[n, 0]), float(fac[n, 1]), float(fac[n, 2])))
        err = (out[n] - want).abs()
        # hue: a float rounding of h*6 at a sector boundary moves a channel by O(1e-6); 1e-4 on the [-1, 1] scale
        assert float(err.max()) < 1e-4, (ops, n, float(err.max()))
    plain = T.color_jitter(x.to(DEV)).cpu()
    assert torch.allclose(plain, TO.normalize(TO.scale_image(x)), atol=1e-6)


def test_frame_pipeline_feeds_the_encoder_like_host_side_transforms():
    """The fused pipeline (gather + pad + shift, uint8 out) followed by the encoder's on-load normalisation equals the
    host-side path: transforms on the CPU, float32 frames to the encoder."""
    from tacorl_b200 import ops
    from tacorl_b200.networks.visual_encoders.encoder import LMPVisionEncoder
    from tacorl_b200.utils import transforms as T
    from tacorl_b200.utils.rng import noise_tape
    ops.set_precision("fp32")
    g = torch.Generator().manual_seed(3)
    store = torch.randint(0, 256, (30, 3, 84, 84), generator=g, dtype=torch.uint8)
    start, window, Tn, pad = torch.tensor([2, 11]), torch.tensor([8, 6]), 8, 4
    shift = torch.randint(0, 2 * pad + 1, (2 * Tn, 1, 1, 2), generator=g)
    pipe = T.FramePipeline(pad=pad, contrast=0.0, brightness=0.0, hue=0.0)
    with noise_tape([shift]) as tape:
        frames = pipe(store.to(DEV), start, window, Tn)
        assert len(tape) == 0
    assert frames.dtype == torch.uint8
    torch.manual_seed(0)
    enc = LMPVisionEncoder().to(DEV)
    with torch.no_grad():
        got = enc(frames.view(2 * Tn, 3, 84, 84))
        host = []
        for b in range(2):
            f = TO.pad_frames(store[int(start[b]):int(start[b]) + int(window[b])], int(window[b]), Tn).float()
            f = TO.random_shifts(f, pad, shift.view(2, Tn, 1, 1, 2)[b])
            host.append(TO.normalize(TO.scale_image(f.round().to(torch.uint8))))
        want = enc(torch.cat(host).to(DEV))
    assert float((got - want).abs().max()) < 1e-5 * max(1.0, float(want.abs().max()))
