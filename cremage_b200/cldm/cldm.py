"""ControlNet residual injection into the B200 UNet -- mirror of `ControlledUnetModel` (reference
modules/cldm/cldm.py:28-70).

The reference subclass runs the frozen UNet and adds the ControlNet's residuals in two places: `h += control.pop()`
after the middle block and `torch.cat([h, hs.pop() + control.pop()], 1)` before every output block (skipped when
`only_mid_control`).  Here the additions are one `cb_add_nchw_to_nhwc` launch each (the residuals arrive NCHW from the
ControlNet, which is NOT part of this repository: it produces an input of the hot path, like CLIP produces `context`),
and the concatenation stays virtual as in the plain UNet.  `control` is consumed exactly like the reference consumes it:
popped from the END of the caller's list.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from ..engine import GraphedCall, require_cuda
from ..ldm.modules.diffusionmodules.openaimodel import UNetModel


class ControlledUnetModel(UNetModel):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._graphed_ctrl: Optional[GraphedCall] = None

    def _forward_ctrl(self, x, t, ctx, *control):
        return self._forward_impl(x, t, ctx, None, control)

    def forward(self, x, timesteps=None, context=None, control: Optional[List[torch.Tensor]] = None,
                only_mid_control: bool = False, **kwargs):
        if control is None:   # "If control is None, it should act the same way as UNetModel" (cldm.py:42)
            return super().forward(x, timesteps=timesteps, context=context, **kwargs)
        if self.num_classes is not None:
            raise NotImplementedError("cremage_b200: ControlledUnetModel mirrors the SD1.5 (ldm) class: no vector conditioning")
        require_cuda(x, "ControlledUnetModel.forward")
        if timesteps is None or context is None:
            raise ValueError("ControlledUnetModel.forward needs timesteps and context")
        n_need = 1 if only_mid_control else 1 + len(self.output_blocks)
        if len(control) < n_need:
            raise IndexError("pop from empty list")   # what the reference's control.pop() raises
        used = [control.pop() for _ in range(n_need)]           # middle residual first, then one per output block
        t = timesteps.to(device=x.device, dtype=torch.float32).contiguous()
        ctx = context.to(device=x.device)
        xin = x.contiguous()
        used = [c.to(device=x.device).contiguous() for c in used]
        if self.use_cuda_graph and not torch.cuda.is_current_stream_capturing():
            if self._graphed_ctrl is None:
                self._graphed_ctrl = GraphedCall(self._forward_ctrl)
            self.packed(x.device)   # refresh packs (and drop stale graphs) if parameters changed
            out = self._graphed_ctrl(xin, t, ctx, *used)
        else:
            out = self._forward_ctrl(xin, t, ctx, *used)
        return out.to(x.dtype)

    def _reset_graphs(self):
        super()._reset_graphs()
        if self.__dict__.get("_graphed_ctrl") is not None:
            self._graphed_ctrl.reset()
