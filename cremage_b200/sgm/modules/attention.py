"""sgm (SDXL) mirror of `sgm.modules.attention` (modules/sdxl/sgm/modules/attention.py): SpatialTransformer (:902-1133)
honours `use_linear` (nn.Linear proj_in / proj_out on the [b, hw, c] view, :999,:1056); BasicTransformerBlock (:724-847)
and CrossAttention (:358-534, SDPA) compute the same function as their ldm twins and share the implementation."""
from ...ldm.modules.attention import (BasicTransformerBlock, CrossAttention, FeedForward, GEGLU,  # noqa: F401
                                      MemoryEfficientCrossAttention)
from ...ldm.modules.attention import SpatialTransformer as _LdmSpatialTransformer


class SpatialTransformer(_LdmSpatialTransformer):
    _HONOR_USE_LINEAR = True

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0.0, context_dim=None, disable_self_attn=False,
                 use_linear=False, attn_type="softmax", use_checkpoint=True, sdp_backend=None, lora_ranks=None,
                 lora_weights=None, ipa_scale=1.0, ipa_num_tokens=0):
        if isinstance(context_dim, (list, tuple)) and len(context_dim) != depth:
            context_dim = depth * [context_dim[0]]          # sgm attention.py:951-960
        elif context_dim is not None and not isinstance(context_dim, (list, tuple)):
            context_dim = depth * [context_dim]
        super().__init__(in_channels, n_heads, d_head, depth=depth, dropout=dropout, context_dim=context_dim,
                         disable_self_attn=disable_self_attn, use_linear=use_linear, use_checkpoint=use_checkpoint,
                         lora_ranks=lora_ranks, lora_weights=lora_weights, ipa_scale=ipa_scale,
                         ipa_num_tokens=ipa_num_tokens)
