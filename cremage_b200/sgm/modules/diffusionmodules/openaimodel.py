"""sgm (SDXL) mirror of `sgm.modules.diffusionmodules.openaimodel.UNetModel`
(modules/sdxl/sgm/modules/diffusionmodules/openaimodel.py:476-874): same constructor and
`forward(x, timesteps, context, y)`; per-level transformer depth, linear proj_in/out, `label_emb` for the pooled
text / size vector `y` (:617-625,851-859).  Implementation shared with the ldm mirror."""
from typing import List, Optional, Tuple, Union

from ....ldm.modules.diffusionmodules.openaimodel import (Downsample, ResBlock, TimestepBlock,  # noqa: F401
                                                          TimestepEmbedSequential, Upsample)
from ....ldm.modules.diffusionmodules.openaimodel import UNetModel as _UNetCore
from ..attention import SpatialTransformer


class UNetModel(_UNetCore):
    _ST_CLS = SpatialTransformer

    def __init__(self, in_channels: int, model_channels: int, out_channels: int, num_res_blocks: int,
                 attention_resolutions: int, dropout: float = 0.0, channel_mult: Union[List, Tuple] = (1, 2, 4, 8),
                 conv_resample: bool = True, dims: int = 2, num_classes: Optional[Union[int, str]] = None,
                 use_checkpoint: bool = False, num_heads: int = -1, num_head_channels: int = -1,
                 num_heads_upsample: int = -1, use_scale_shift_norm: bool = False, resblock_updown: bool = False,
                 transformer_depth: int = 1, context_dim: Optional[int] = None,
                 disable_self_attentions: Optional[List[bool]] = None, num_attention_blocks: Optional[List[int]] = None,
                 disable_middle_self_attn: bool = False, disable_middle_transformer: bool = False,
                 use_linear_in_transformer: bool = False, spatial_transformer_attn_type: str = "softmax",
                 adm_in_channels: Optional[int] = None, lora_ranks: List[int] = None, lora_weights: List[float] = None):
        if disable_middle_transformer:
            raise NotImplementedError("cremage_b200: disable_middle_transformer is not implemented")
        super().__init__(image_size=None, in_channels=in_channels, model_channels=model_channels,
                         out_channels=out_channels, num_res_blocks=num_res_blocks,
                         attention_resolutions=attention_resolutions, dropout=dropout, channel_mult=channel_mult,
                         conv_resample=conv_resample, dims=dims, num_classes=num_classes,
                         use_checkpoint=use_checkpoint, num_heads=num_heads, num_head_channels=num_head_channels,
                         num_heads_upsample=num_heads_upsample, use_scale_shift_norm=use_scale_shift_norm,
                         resblock_updown=resblock_updown, use_spatial_transformer=True,
                         transformer_depth=transformer_depth, context_dim=context_dim, legacy=False,
                         disable_self_attentions=disable_self_attentions, num_attention_blocks=num_attention_blocks,
                         disable_middle_self_attn=disable_middle_self_attn,
                         use_linear_in_transformer=use_linear_in_transformer, lora_ranks=lora_ranks,
                         lora_weights=lora_weights, adm_in_channels=adm_in_channels)
