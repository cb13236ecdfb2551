from .modeling_finetune import *  # noqa: F401,F403
