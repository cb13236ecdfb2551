"""Tube masking for DAPT / MAE pre-training — drop-in for the reference's masking_generator.py.

`TubeMaskingGenerator` keeps the reference's interface and, for the same `np.random` state, its exact draws
(masking_generator.py:3-23): a flat float array [frames * H * W] of {0, 1}, 1 = masked, the same spatial mask in
every temporal slot.  `batch_masks` stacks draws into the bool [B, N] tensor `PretrainVisionTransformer.forward`
takes (what the DataLoader's collate does with the per-sample masks, run_mae_pretraining.py / engine_for_pretraining).
"""
import numpy as np
import torch

__all__ = ["TubeMaskingGenerator", "batch_masks"]


class TubeMaskingGenerator:
    def __init__(self, input_size, mask_ratio):
        self.frames, self.height, self.width = input_size
        self.num_patches_per_frame = self.height * self.width
        self.total_patches = self.frames * self.num_patches_per_frame
        self.num_masks_per_frame = int(mask_ratio * self.num_patches_per_frame)
        self.total_masks = self.frames * self.num_masks_per_frame

    def __repr__(self):
        return "Maks: total patches {}, mask patches {}".format(self.total_patches, self.total_masks)  # sic (mg:12-15)

    def __call__(self):
        keep = self.num_patches_per_frame - self.num_masks_per_frame
        mask_per_frame = np.concatenate([np.zeros(keep), np.ones(self.num_masks_per_frame)])
        np.random.shuffle(mask_per_frame)  # the one random draw of the reference (mg:21)
        return np.tile(mask_per_frame, (self.frames, 1)).flatten()

    @property
    def num_visible(self):
        return self.total_patches - self.total_masks


def batch_masks(generator, batch_size, device=None):
    """`batch_size` consecutive draws -> bool [B, N] (True = masked), optionally moved to `device`."""
    m = torch.from_numpy(np.stack([generator() for _ in range(batch_size)])).to(torch.bool)
    return m if device is None else m.to(device, non_blocking=True)
