"""GPU clip loader: host RNG mirror of the reference's frame sampler on CPU; on the GPU a batch built
from JPEG bytes equals the reference's PIL chain (Image.open -> Scale -> RandomCrop -> ToTensor, third view = first
window again), and feeds a pretraining step."""
import io
import random
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from dualvar_b200 import loader as L

PIL = pytest.importorskip("PIL.Image")


def _reference_frame_sampler(total, num_frames, ds, repeat_prob=0.25):
    """dataset/local_dataset.py:245-258 literally (center_lower = 0, center_upper = total as sample_prototype passes)."""
    center_lower, center_upper = 0, total
    length = num_frames
    center_ind = np.random.randint(center_lower, center_upper)
    diff_seq = (np.arange(length) - length // 2) * ds
    if random.uniform(0., 1.) >= repeat_prob:
        center_lower = 0
    if random.uniform(0., 1.) >= repeat_prob:
        center_upper = total
    return np.clip(diff_seq + center_ind, center_lower, center_upper - 1).astype(np.int32)


def test_frame_sampler_mirrors_reference_rng_order():
    for seed, total, nf, ds in [(0, 40, 16, 1), (1, 10, 16, 2), (2, 300, 8, 4)]:
        np.random.seed(seed); random.seed(seed)
        want = [_reference_frame_sampler(total, nf, ds) for _ in range(5)]
        tail = (np.random.randint(0, 1000), random.random())
        np.random.seed(seed); random.seed(seed)
        got = [L.draw_frame_indices(total, nf, ds) for _ in range(5)]
        assert all(np.array_equal(a, b) for a, b in zip(want, got))
        assert tail == (np.random.randint(0, 1000), random.random())          # same number of draws consumed


def _dataset(rng, n_videos=3, n_frames=20, h=120, w=160):
    store = {}
    videos = []
    for v in range(n_videos):
        name = f"v{v}"
        base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3)).astype(np.uint8)
        img = np.asarray(PIL.fromarray(base).resize((w, h), PIL.BICUBIC)).astype(np.float32)
        for i in range(n_frames):
            fr = np.clip(img + 3.0 * i + rng.normal(0, 6, img.shape), 0, 255).astype(np.uint8)
            b = io.BytesIO()
            PIL.fromarray(fr).save(b, "JPEG", quality=85, subsampling=2)
            store[(name, i)] = b.getvalue()
        videos.append((name, n_frames, v))
    return videos, (lambda name, i: store[(name, i)]), store


@pytest.mark.gpu
def test_loader_batch_equals_the_pil_chain_and_feeds_a_step():
    from dualvar_b200 import models as PM
    from dualvar_b200.engine import RawClips
    rng = np.random.default_rng(0)
    videos, read, store = _dataset(rng)
    null_only = dict(weights=((1.0, 0.0, 0.0),) * 3)          # every view takes the null transform: Scale + RandomCrop + ToTensor
    ld = L.ClipLoader(videos, read, batch_size=2, device="cuda:0", indices=[2, 0], plan_kwargs=null_only, decode_threads=2)
    np.random.seed(3); random.seed(3); torch.manual_seed(3)
    batch = next(iter(ld))
    assert isinstance(batch["seq"], RawClips) and batch["vid"].tolist() == [2, 0]
    got = batch["seq"].frames.cpu().numpy()                   # uint8 (B, 3, 48, 112, 112)
    assert got.shape == (2, 3, 48, 112, 112) and got.dtype == np.uint8
    # the same draws again, then the reference chain with Pillow
    np.random.seed(3); random.seed(3); torch.manual_seed(3)
    for b, index in enumerate([2, 0]):
        name, label, (w1, w2), plan = ld.sample(index)
        frames = [PIL.open(io.BytesIO(store[(name, int(i))])) for i in np.concatenate((w1, w2))]
        frames = frames[:32] + frames[:16]                    # aug_series (dataset/local_dataset.py:288-289)
        for t, im in enumerate(frames):
            left, top = (int(v) for v in plan["crops"][0, t // 16])
            want = np.asarray(im.resize((128, 171), PIL.BICUBIC).crop((left, top, left + 112, top + 112))).transpose(2, 0, 1)
            assert np.array_equal(got[b, :, t], want), (b, t)
    # the default transform mix (jitter / blur branches) runs and a pretraining step consumes the batch
    ld2 = L.ClipLoader(videos, read, batch_size=2, device="cuda:0", indices=[0, 1, 2], drop_last=False)
    assert len(ld2) == 2
    model = PM.SimCLR_TimeSeriesV4("r3d", 128, 0.07, False, True, 2, 64, 0.07, 0.07, "clip-sr-tc",
                                   SimpleNamespace(shufflerank_theta=0.05)).to("cuda:0").train()
    sizes = []
    for bt in ld2:
        ret = model(bt["seq"])
        loss = sum(v for k, v in ret.items() if "loss" in k)
        assert torch.isfinite(loss)
        sizes.append(bt["seq"].block_shape[0])
    assert sizes == [2, 1]
    loss.backward()
