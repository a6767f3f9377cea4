"""Fused replacements for hot elementwise/normalisation ops of the third-party encoder body.

``fuse_backbone(backbone)`` swaps every ``torch.nn.LayerNorm`` whose width the kernel supports for ``FusedLayerNorm``
(same parameters, same state_dict keys, same numerics up to fp32 rounding). Under autocast PyTorch runs layer_norm in
fp32 (upcasting bf16 activations and writing fp32 outputs); the fused module keeps bf16 in / bf16 out with fp32
statistics, which also removes the surrounding cast kernels. Unsupported widths keep the PyTorch module.
"""
import torch

from ... import ops


class FusedLayerNorm(torch.nn.LayerNorm):
    """Drop-in torch.nn.LayerNorm over the last dimension on the sm_100a kernels. CUDA tensors only: like the rest of
    the package there is no eager / CPU path (fuse_backbone only swaps modules the kernel supports; fp16 activations
    are normalised in fp32, which is also what autocast does for torch's LayerNorm)."""

    def forward(self, x):
        if x.dtype not in (torch.bfloat16, torch.float32):
            return ops.layer_norm(x.float(), self.weight, self.bias, self.eps).to(x.dtype)
        return ops.layer_norm(x, self.weight, self.bias, self.eps)


class FusedLinear(torch.nn.Linear):
    """torch.nn.Linear whose bias gradient comes from the fused column-sum kernel. The GEMMs themselves are cuBLAS
    (library) in both branches: with gradients off there is no bias gradient to fuse and the stock op is called."""

    def forward(self, x):
        if not x.is_cuda:
            raise ops._lib.SparseB200Error("FusedLinear runs on CUDA tensors only (no CPU path)")
        if torch.is_grad_enabled() and self.bias.requires_grad:
            return ops.linear(x, self.weight, self.bias)
        return torch.nn.functional.linear(x, self.weight, self.bias)


def _skip_linear(module, backbone):
    """The vocabulary decoder is never run as a Linear (the fused head consumes its parameters directly)."""
    out = backbone.get_output_embeddings() if hasattr(backbone, "get_output_embeddings") else None
    return out is not None and module is out


def fuse_backbone(backbone):
    """In place; returns the number of modules replaced."""
    swapped = 0
    for parent in backbone.modules():
        for name, child in list(parent.named_children()):
            if type(child) is torch.nn.Linear and child.bias is not None and not _skip_linear(child, backbone) \
                    and ops.colsum_supported(child.out_features):
                fused = FusedLinear(child.in_features, child.out_features, bias=True, device="meta")
                fused.weight, fused.bias = child.weight, child.bias
                fused.train(child.training)
                setattr(parent, name, fused)
                swapped += 1
                continue
            if type(child) is torch.nn.LayerNorm and len(child.normalized_shape) == 1 and child.elementwise_affine \
                    and child.bias is not None and ops.layer_norm_supported(child.normalized_shape[0]):
                fused = FusedLayerNorm(child.normalized_shape, eps=child.eps)
                fused.weight, fused.bias = child.weight, child.bias  # share the parameters (tied state_dict keys)
                fused.train(child.training)
                setattr(parent, name, fused)
                swapped += 1
    return swapped
