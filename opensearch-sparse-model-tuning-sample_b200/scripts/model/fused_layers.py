"""Fused replacements for hot elementwise/normalisation ops of the third-party encoder body.

``fuse_backbone(backbone)`` swaps every ``torch.nn.LayerNorm`` whose width the kernel supports for ``FusedLayerNorm``
(same parameters, same state_dict keys, same numerics up to fp32 rounding). Under autocast PyTorch runs layer_norm in
fp32 (upcasting bf16 activations and writing fp32 outputs); the fused module keeps bf16 in / bf16 out with fp32
statistics, which also removes the surrounding cast kernels. Unsupported widths keep the PyTorch module.
"""
import torch

from ... import ops


class FusedLayerNorm(torch.nn.LayerNorm):
    """Drop-in torch.nn.LayerNorm over the last dimension, sm_100a kernels for CUDA tensors."""

    def forward(self, x):
        if (x.is_cuda and self.elementwise_affine and self.bias is not None and len(self.normalized_shape) == 1
                and x.dtype in (torch.bfloat16, torch.float32)):
            return ops.layer_norm(x, self.weight, self.bias, self.eps)
        return super().forward(x)


def fuse_backbone(backbone):
    """In place; returns the number of modules replaced."""
    swapped = 0
    for parent in backbone.modules():
        for name, child in list(parent.named_children()):
            if type(child) is torch.nn.LayerNorm and len(child.normalized_shape) == 1 and child.elementwise_affine \
                    and child.bias is not None and ops.layer_norm_supported(child.normalized_shape[0]):
                fused = FusedLayerNorm(child.normalized_shape, eps=child.eps)
                fused.weight, fused.bias = child.weight, child.bias  # share the parameters (tied state_dict keys)
                fused.train(child.training)
                setattr(parent, name, fused)
                swapped += 1
    return swapped
