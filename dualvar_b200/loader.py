"""Pretraining clip loader with decode and augmentation on the GPU.

What the reference does per sample on DataLoader worker CPUs (dataset/local_dataset.py:268-308 + the transform of
pretrain.py:491-532): draw two 16-frame index windows, ``Image.open`` 32 JPEG files (libjpeg), re-use the first window
as the third view (``aug_series``), then per view Scale((128,171)) -> RandomCrop(112) -> ToTensor [-> ColorJitter ->
GaussianBlur] and stack to (3, 48, 112, 112) float32; the training loop adds Normalize + view/transpose on the GPU.

Here the host only draws random numbers - in the reference's order, from the generators it uses (``np.random`` /
``random`` / ``torch.rand``) - reads the compressed bytes and Huffman-decodes them on threads; everything that touches
pixels runs on the GPU for the whole batch at once and is bit-exact with the reference's PIL chain stage by stage:
    jpeg.decode_batch (csrc/jpeg.cu) -> frames.scale_crop / color_jitter / gaussian_blur (csrc/frames.cu, ...) ->
    engine.RawClips -> the ingest kernel (ToTensor's /255, Normalize, NDHWC / space-to-depth layout).
A batch is ``{'seq': RawClips, 'vid': labels}`` - what pretrain_loop.train_one_epoch and the models take.

Random-number order (single-process order, i.e. the reference with ``num_workers=0``): sample after sample, first the
dataset's draws (``draw_frame_indices`` x 2, preceded by the flip draw when ``rand_flip``), then the transform's
(``frames.draw_plan`` for that sample).
"""
import random

import numpy as np
import torch

from . import frames as FR
from . import jpeg as JP


def draw_frame_indices(total, num_frames=16, ds=1, repeat_prob=0.25, np_random=None, py_random=random):
    """``UCF101LMDB_2CLIP_Stage_Prototype.frame_sampler(total, 0, total)`` (dataset/local_dataset.py:245-258): a window of
    ``num_frames`` indices with stride ``ds`` around a random centre, clipped to the video; with probability
    ``repeat_prob`` per side the clip keeps that boundary at the centre-draw limits (here 0 / total, so both branches
    coincide - the draws are still consumed, in this order: np.random.randint, random.uniform, random.uniform)."""
    np_random = np.random if np_random is None else np_random
    center_lower, center_upper = 0, total
    center = np_random.randint(center_lower, center_upper)
    diff = (np.arange(num_frames) - num_frames // 2) * ds
    if py_random.uniform(0., 1.) >= repeat_prob:
        center_lower = 0
    if py_random.uniform(0., 1.) >= repeat_prob:
        center_upper = total
    return np.clip(diff + center, center_lower, center_upper - 1).astype(np.int32)


class ClipLoader:
    """Iterable of pretraining batches.

    videos      sequence of (name, n_frames, label) - the rows of the reference's ``video_subset``
    read_frame  callable (name, frame_index) -> bytes of ``{name}/image_{frame_index + 1:05d}.jpg`` (file, LMDB, ...)
    indices     iterable of dataset indices for one epoch (a DistributedSampler works); default: a fresh
                ``torch.randperm`` per epoch like RandomSampler
    """

    def __init__(self, videos, read_frame, batch_size, device, num_frames=16, ds=1, aug_series=True, rand_flip=False,
                 n_views=3, scaled=(128, 171), crop=(112, 112), indices=None, drop_last=True, decode_threads=None,
                 plan_kwargs=None):
        assert n_views == 3 and aug_series, "the DualVar loader yields 3 views: two windows + the first one again"
        self.videos, self.read_frame = list(videos), read_frame
        self.batch_size, self.device = batch_size, torch.device(device)
        self.num_frames, self.ds, self.rand_flip = num_frames, ds, rand_flip
        self.n_views, self.scaled, self.crop = n_views, tuple(scaled), tuple(crop)
        self.indices, self.drop_last, self.decode_threads = indices, drop_last, decode_threads
        self.plan_kwargs = dict(plan_kwargs or {})

    def __len__(self):
        n = len(self.indices) if self.indices is not None and hasattr(self.indices, "__len__") else len(self.videos)
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def sample(self, index):
        """The host half of ``__getitem__`` + transform draws for one dataset item: (name, label, two index windows, plan)."""
        name, vlen, label = self.videos[index]
        flip = random.randint(0, 1) if self.rand_flip else 0
        w1 = draw_frame_indices(vlen, self.num_frames, self.ds)
        if flip:
            w1 = w1[::-1]
        w2 = draw_frame_indices(vlen, self.num_frames, self.ds)
        if flip:
            w2 = w2[::-1]
        plan = FR.draw_plan(1, n_views=self.n_views, seq_len=self.num_frames, scaled=self.scaled, crop=self.crop,
                            **self.plan_kwargs)
        return name, label, (w1, w2), plan

    def collate(self, items):
        """Decode + augment a list of ``sample()`` results on the GPU -> {'seq': RawClips, 'vid': int64 labels}."""
        T = self.num_frames
        files = [self.read_frame(name, int(i)) for name, _, (w1, w2), _ in items for i in np.concatenate((w1, w2))]
        decoded = JP.decode_batch(files, self.device, threads=self.decode_threads)        # (B * 2T, H, W, 3) uint8
        B = len(items)
        H, W = decoded.shape[1:3]
        decoded = decoded.view(B, 2 * T, H, W, 3)
        # aug_series: seq[:2T] + seq[:T] - the third view re-uses the first window's frames (local_dataset.py:288-289)
        order = torch.cat([torch.arange(2 * T), torch.arange(T)]).to(self.device)
        fr = decoded.index_select(1, order)                                                 # (B, 3T, H, W, 3)
        crops = torch.cat([p["crops"] for *_, p in items])
        jitter = torch.cat([p["jitter"] for *_, p in items])
        blur = [s for *_, p in items for s in p["blur"]]
        any_aug = bool((jitter[:, 0] != 0).any()) or any(s > 0 for s in blur)
        clips = FR.stage_clips(fr, crops, self.n_views, jitter=jitter if any_aug else None, blur=blur if any_aug else None)
        return {"seq": clips, "vid": torch.tensor([label for _, label, _, _ in items], dtype=torch.int64)}

    def __iter__(self):
        idx = self.indices if self.indices is not None else torch.randperm(len(self.videos)).tolist()
        batch = []
        for i in idx:
            batch.append(self.sample(int(i)))
            if len(batch) == self.batch_size:
                yield self.collate(batch)
                batch = []
        if batch and not self.drop_last:
            yield self.collate(batch)
