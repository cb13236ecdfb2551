"""B200-native counterpart of the reference's run_inference_simple.py (cited as ris:line) — the self-contained
"load a checkpoint, score a folder of frames" path that BASELINE config 1 names.

    VisionTransformerInfer        ris:279-382   the classifier whose forward returns PROBABILITIES (softmax inside, ris:381)
    get_video_vit_small / _base   ris:385-407   its two factories (same arguments; `with_flash` is accepted and ignored:
                                                there is one attention kernel)
    prepare_image                 ris:18-37     host restatement (BGR uint8 HWC -> normalised RGB CHW fp32) for callers
                                                that still build fp32 windows themselves; the scoring loop below does
                                                the same arithmetic on the device (stad_normalize_frames_u8)
    score_frame_folder / main     ris:410-465   the frame-by-frame loop over a folder of images

State-dict keys are those of ris:279-356 (identical to modeling_finetune.VisionTransformer), so the released
checkpoints load with `model.load_state_dict(torch.load(ckpt))` as at ris:424-425.
"""
import os
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from .modeling_finetune import VisionTransformer
from .runner import IMAGENET_MEAN, IMAGENET_STD, StreamingScorer

IMG_EXT = (".png", ".jpg", ".jpeg", ".JPG", ".JPEG")   # ris:15

__all__ = ["IMG_EXT", "prepare_image", "VisionTransformerInfer", "get_video_vit_small", "get_video_vit_base",
           "iter_frame_folder", "score_frame_folder", "main"]


def prepare_image(img, mean, std, inplace=True):
    """ris:18-37 on the host: BGR uint8 [H, W, 3] -> RGB fp32 [3, H, W], / 255, (x - mean) / std."""
    if not (len(img.shape) == 3):
        raise TypeError(f'Input must be a 3D image tensor (C, H, W), but got shape: {img.shape}')
    rgb = np.ascontiguousarray(np.asarray(img)[:, :, ::-1].transpose(2, 0, 1))      # cvtColor(BGR2RGB) + HWC -> CHW
    out = torch.from_numpy(rgb).float().div_(255.0)
    mean = torch.as_tensor(mean, dtype=out.dtype).view(-1, 1, 1)
    std = torch.as_tensor(std, dtype=out.dtype).view(-1, 1, 1)
    return out.sub_(mean).div_(std)


class VisionTransformerInfer(VisionTransformer):
    """VisionTransformerInfer (ris:279-382): forward(x) = softmax(head(forward_features(x))), shape [B, num_classes]."""

    @torch.no_grad()
    def forward(self, x):
        return self.forward_probs(x)[1]


def _infer_model(embed_dim, depth, num_heads, with_flash):
    return VisionTransformerInfer(patch_size=16, embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=4,
                                  qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), num_classes=2, all_frames=16,
                                  tubelet_size=2, final_reduction="fc_norm", init_scale=0.001, use_flash_attn=with_flash)


def get_video_vit_small(with_flash=False):
    """ris:385-395."""
    return _infer_model(384, 12, 6, with_flash)


def get_video_vit_base(with_flash=False):
    """ris:397-407."""
    return _infer_model(768, 12, 12, with_flash)


def _natural_key(name):
    """natsorted (ris:433) for frame file names: digit runs compare as numbers."""
    import re
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", name)]


def score_frame_folder(model, frames_folder, size=(224, 224)):
    """The loop of ris:428-463 as a generator of (frame index, risk probability)."""
    for i, _, probs in iter_frame_folder(model, frames_folder, size):
        yield i, float(probs[1])


def iter_frame_folder(model, frames_folder, size=(224, 224)):
    """The loop of ris:428-463 / run_inference.py:69-109 as a generator of (frame index, logits [2], probs [2]).

    The reference fills the window with the first 16 images, predicts, and then — `if i < 16: continue`, ris:447-448 —
    SKIPS the next 16 images entirely (they never enter the window) before it appends one image per prediction; the
    same images enter the window here, so the probabilities line up with the reference's print-out.  Frames are read
    with cv2 (BGR) and handed over as uint8; resize (INTER_CUBIC, ris:438/451) and prepare_image run on the device."""
    import cv2
    names = sorted([f for f in os.listdir(frames_folder) if os.path.splitext(f)[1] in IMG_EXT], key=_natural_key)
    assert len(names) > 15, "We need at least 16 frames!"
    scorer = StreamingScorer(model, bgr=True, mean=IMAGENET_MEAN, std=IMAGENET_STD)
    assert (scorer.H, scorer.W) == tuple(size)

    def read(name):
        return torch.from_numpy(cv2.imread(os.path.join(frames_folder, name)))

    out = None
    for name in names[:16]:
        out = scorer.push(read(name))
    yield 15, out[0], out[1]                                     # "First prediction", ris:443-444
    for i, name in enumerate(names[16:]):
        if i < 16:                                               # ris:447-448
            continue
        out = scorer.push(read(name))
        yield i, out[0], out[1]                                  # ris:461-462 prints the loop index


def main(ckpt_file, frames_folder):
    """ris:410-465 with the ViT-S classifier."""
    model = get_video_vit_small(with_flash=False)
    model.default_cfg = {'url': "", 'num_classes': 400, 'input_size': (3, 224, 224), 'pool_size': None, 'crop_pct': .9,
                         'interpolation': 'bicubic', 'mean': IMAGENET_MEAN, 'std': IMAGENET_STD}
    model.load_state_dict(torch.load(ckpt_file, map_location='cpu'))
    model.to(torch.device("cuda")).eval()
    first = True
    for i, risk in score_frame_folder(model, frames_folder):
        print(f"First prediction: risk probability {risk:.4f}" if first else f"Frame {i}, risk probability: {risk:.2f}")
        first = False
    print("Done!")
