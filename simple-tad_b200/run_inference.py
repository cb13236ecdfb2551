"""B200-native counterpart of the reference's run_inference.py (cited as ri:line): score a folder of frames with a
registered classifier (`create_model`, ri:40-55), one prediction per new frame (ri:69-109).  The loop itself lives in
run_inference_simple.iter_frame_folder (the two reference scripts share it); this model returns logits, the softmax is
applied outside (ri:104-107)."""
import torch

from . import modeling_finetune  # noqa: F401  (registers the vit_* factories, ri:7)
from .registry import create_model
from .run_inference_simple import IMG_EXT, iter_frame_folder, prepare_image  # noqa: F401
from .runner import IMAGENET_MEAN, IMAGENET_STD


def build_model(model_name='vit_small_patch16_224', with_flash_attn=False):
    """The create_model call of ri:40-55, kwarg for kwarg, and the default_cfg of ri:56-62."""
    model = create_model(model_name=model_name, pretrained=False, num_classes=2, all_frames=16, tubelet_size=2,
                         fc_drop_rate=0.0, drop_rate=0.0, drop_path_rate=0.1, attn_drop_rate=0.0, drop_block_rate=None,
                         use_checkpoint=False, final_reduction="fc_norm", init_scale=0.001, use_flash_attn=with_flash_attn)
    model.default_cfg = {'url': "", 'num_classes': 400, 'input_size': (3, 224, 224), 'pool_size': None, 'crop_pct': .9,
                         'interpolation': 'bicubic', 'mean': IMAGENET_MEAN, 'std': IMAGENET_STD}
    return model


def main(ckpt_file, frames_folder, model_name='vit_small_patch16_224', with_flash_attn=False):
    """ri:37-109."""
    model = build_model(model_name, with_flash_attn)
    model.load_state_dict(torch.load(ckpt_file, map_location='cpu'))   # ri:64-65
    model.to(torch.device("cuda")).eval()
    first = True
    for i, logits, probs in iter_frame_folder(model, frames_folder):
        if first:
            print(f"First prediction: {logits.unsqueeze(0)}")
            first = False
        else:
            print(f"Frame {i}, risk logit: {float(logits[1]):.2f}, risk prob: {float(probs[1]):.2f}")
    print("Done!")
