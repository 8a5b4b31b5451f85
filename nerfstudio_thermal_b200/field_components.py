"""Field components with the reference's constructor signatures, attributes and state_dict keys.

Mirrors `nerfstudio/field_components/{encodings,mlp,activations,spatial_distortions,embedding}.py` for the
classes thermal-nerfacto instantiates.  `implementation` is accepted for signature compatibility; this
package has exactly one backend (libtn_b200, sm_100a) and the value is recorded but not dispatched on.
"""
from typing import Literal, Optional, Set, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from . import ops


class FieldComponent(nn.Module):
    def __init__(self, in_dim: Optional[int] = None, out_dim: Optional[int] = None) -> None:
        super().__init__()
        self.in_dim = in_dim
        self.out_dim = out_dim

    def get_out_dim(self) -> int:
        if self.out_dim is None:
            raise ValueError("Output dimension has not been set")
        return self.out_dim


class Encoding(FieldComponent):
    def __init__(self, in_dim: int) -> None:
        if in_dim <= 0:
            raise ValueError("Input dimension should be greater than zero")  # encodings.py:43-44
        super().__init__(in_dim=in_dim)


class HashEncoding(Encoding):
    """field_components/encodings.py:310-466.  Parameter `hash_table` is [L*T, F] float32 (source of truth)."""

    def __init__(self, num_levels: int = 16, min_res: int = 16, max_res: int = 1024, log2_hashmap_size: int = 19,
                 features_per_level: int = 2, hash_init_scale: float = 0.001,
                 implementation: Literal["tcnn", "torch", "b200"] = "b200",
                 interpolation: Optional[Literal["Nearest", "Linear", "Smoothstep"]] = None) -> None:
        super().__init__(in_dim=3)
        assert interpolation is None or interpolation == "Linear", \
            f"interpolation '{interpolation}' is not supported for torch encoding backend"  # encodings.py:370-373
        self.num_levels = num_levels
        self.min_res = min_res
        self.features_per_level = features_per_level
        self.hash_init_scale = hash_init_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.hash_table_size = 2**log2_hashmap_size
        self.implementation = implementation
        levels = torch.arange(num_levels)
        self.growth_factor = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1
        self.scalings = torch.floor(min_res * self.growth_factor**levels)  # encodings.py:343-345, verbatim
        self.hash_offset = levels * self.hash_table_size
        self.tcnn_encoding = None
        table = torch.rand(size=(self.hash_table_size * num_levels, features_per_level)) * 2 - 1
        self.hash_table = nn.Parameter(table * hash_init_scale)  # encodings.py:377-379
        self.spec = ops.HashGridSpec(self.scalings.tolist(), features_per_level, log2_hashmap_size)
        # optional fp16 gather cache of the table ("fp16 features" mode); refreshed by refresh_half_cache()
        self.use_half_table = False
        self._half_cache: Optional[Tensor] = None
        self._half_version = -1
        # optional accumulation target for the table gradient (set by parallel.FlatGradBuffer.attach_sinks)
        self.grad_sink: Optional[Tensor] = None

    def get_out_dim(self) -> int:
        return self.num_levels * self.features_per_level

    def hash_fn(self, in_tensor: Tensor) -> Tensor:
        """Rows of the eight corners are produced inside the kernel; this keeps the reference's helper
        (encodings.py:401-418) for callers that hash integer coordinates themselves (int64 torch math)."""
        t = in_tensor * torch.tensor([1, 2654435761, 805459861]).to(in_tensor.device)
        x = torch.bitwise_xor(torch.bitwise_xor(t[..., 0], t[..., 1]), t[..., 2])
        x %= self.hash_table_size
        x += self.hash_offset.to(x.device)
        return x

    def refresh_half_cache(self) -> None:
        self._half_cache = self.hash_table.detach().to(torch.float16)
        self._half_version = self.hash_table._version

    def _half(self) -> Optional[Tensor]:
        if not self.use_half_table:
            return None
        if self._half_cache is None or self._half_version != self.hash_table._version \
                or self._half_cache.device != self.hash_table.device:
            self.refresh_half_cache()
        return self._half_cache

    def corner_indices(self, in_tensor: Tensor) -> Tensor:
        """int32 [N, L, 8] table rows (reference corner order) -- parity-test hook."""
        return ops.hash_encode_indices(in_tensor.reshape(-1, 3), self.spec)

    def forward(self, in_tensor: Tensor) -> Tensor:
        assert in_tensor.shape[-1] == 3
        flat = in_tensor.reshape(-1, 3)
        # [R, S, 3] inputs are ray samples in ray-major order: tell the kernels (tiling hint, see tn_b200.h)
        spr = in_tensor.shape[-2] if in_tensor.dim() == 3 else 0
        out = ops.hash_encode(flat, self.hash_table, self.spec, self._half(), self.grad_sink, samples_per_ray=spr)
        return out.view(*in_tensor.shape[:-1], self.get_out_dim())


class SHEncoding(Encoding):
    """field_components/encodings.py:755-800 (levels=4 is what the fields use)."""

    def __init__(self, levels: int = 4, implementation: Literal["tcnn", "torch", "b200"] = "b200") -> None:
        super().__init__(in_dim=3)
        if levels <= 0 or levels > 4:
            raise ValueError(f"Spherical harmonic encoding only supports 1 to 4 levels, requested {levels}")
        self.levels = levels

    def get_out_dim(self) -> int:
        return self.levels**2

    @torch.no_grad()
    def forward(self, in_tensor: Tensor) -> Tensor:
        out = ops.sh4(in_tensor.reshape(-1, 3)).view(*in_tensor.shape[:-1], 16)
        return out[..., : self.levels**2]


class MLP(FieldComponent):
    """field_components/mlp.py:60-183.  Parameters live in `layers` (nn.Linear, same keys/layout as the
    reference); the forward runs all layers in one fused kernel."""

    def __init__(self, in_dim: int, num_layers: int, layer_width: int, out_dim: Optional[int] = None,
                 skip_connections: Optional[Tuple[int]] = None, activation: Optional[nn.Module] = nn.ReLU(),
                 out_activation: Optional[nn.Module] = None,
                 implementation: Literal["tcnn", "torch", "b200"] = "b200") -> None:
        super().__init__()
        self.in_dim = in_dim
        assert self.in_dim > 0
        self.out_dim = out_dim if out_dim is not None else layer_width
        self.num_layers = num_layers
        self.layer_width = layer_width
        self.skip_connections = skip_connections
        self._skip_connections: Set[int] = set(skip_connections) if skip_connections else set()
        self.activation = activation
        self.out_activation = out_activation
        self.tcnn_encoding = None
        layers = []
        if num_layers == 1:
            layers.append(nn.Linear(in_dim, self.out_dim))
        else:
            for i in range(num_layers - 1):
                if i == 0:
                    assert i not in self._skip_connections, "Skip connection at layer 0 doesn't make sense."
                    layers.append(nn.Linear(in_dim, layer_width))
                elif i in self._skip_connections:
                    layers.append(nn.Linear(layer_width + in_dim, layer_width))
                else:
                    layers.append(nn.Linear(layer_width, layer_width))
            layers.append(nn.Linear(layer_width, self.out_dim))
        self.layers = nn.ModuleList(layers)
        # optional accumulation targets [(dW_i, db_i)] for the backward kernels (parallel.FlatGradBuffer.attach_sinks)
        self.grad_sinks = None
        if isinstance(out_activation, nn.Sigmoid):
            self._out_act = ops.ACT_SIGMOID
        elif out_activation is None:
            self._out_act = ops.ACT_NONE
        else:
            self._out_act = None
        if not (isinstance(activation, nn.ReLU) and not self._skip_connections and self._out_act is not None
                and ops.mlp_shape_supported(in_dim, layer_width, self.out_dim, num_layers)):
            raise NotImplementedError(
                "libtn_b200 fuses the MLP shapes thermal-nerfacto uses: ReLU hidden activations, no skip "
                "connections, 2-3 layers, in_dim<=64, width in {16,64}, out_dim<=16, output None|Sigmoid; got "
                f"in_dim={in_dim} num_layers={num_layers} width={layer_width} out_dim={self.out_dim}")

    def forward(self, in_tensor: Tensor) -> Tensor:
        # rows may arrive padded to a multiple of 4 floats (fused_ops.field_split emits 63 -> 64): the kernel
        # reads them in place and ignores the padding
        cols = in_tensor.shape[-1]
        if cols != self.in_dim and cols != (self.in_dim + 3) // 4 * 4:
            raise ValueError(f"MLP expects {self.in_dim} input features, got {cols}")
        flat = in_tensor.reshape(-1, cols)
        y = ops.mlp(flat, [l.weight for l in self.layers], [l.bias for l in self.layers], self._out_act,
                    sinks=self.grad_sinks)
        return y.view(*in_tensor.shape[:-1], self.out_dim)


class MLPWithHashEncoding(FieldComponent):
    """field_components/mlp.py:186-294: `model = Sequential(HashEncoding, MLP)` (same state_dict keys)."""

    def __init__(self, num_levels: int = 16, min_res: int = 16, max_res: int = 1024, log2_hashmap_size: int = 19,
                 features_per_level: int = 2, hash_init_scale: float = 0.001, interpolation=None, num_layers: int = 2,
                 layer_width: int = 64, out_dim: Optional[int] = None, skip_connections: Optional[Tuple[int]] = None,
                 activation: Optional[nn.Module] = nn.ReLU(), out_activation: Optional[nn.Module] = None,
                 implementation: Literal["tcnn", "torch", "b200"] = "b200") -> None:
        super().__init__()
        self.in_dim = 3
        self.num_levels, self.min_res, self.max_res = num_levels, min_res, max_res
        self.features_per_level, self.hash_init_scale = features_per_level, hash_init_scale
        self.log2_hashmap_size, self.hash_table_size = log2_hashmap_size, 2**log2_hashmap_size
        self.growth_factor = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1
        self.out_dim = out_dim if out_dim is not None else layer_width
        self.num_layers, self.layer_width = num_layers, layer_width
        encoder = HashEncoding(num_levels=num_levels, min_res=min_res, max_res=max_res,
                               log2_hashmap_size=log2_hashmap_size, features_per_level=features_per_level,
                               hash_init_scale=hash_init_scale, implementation=implementation,
                               interpolation=interpolation)
        mlp = MLP(in_dim=encoder.get_out_dim(), num_layers=num_layers, layer_width=layer_width, out_dim=self.out_dim,
                  skip_connections=skip_connections, activation=activation, out_activation=out_activation,
                  implementation=implementation)
        self.model = nn.Sequential(encoder, mlp)

    def forward(self, in_tensor: Tensor) -> Tensor:
        return self.model(in_tensor)


class _TruncExp(torch.autograd.Function):
    """field_components/activations.py:28-42."""

    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _TruncExp.apply


class SpatialDistortion(nn.Module):
    pass


class SceneContraction(SpatialDistortion):
    """field_components/spatial_distortions.py:42-90 for point inputs.  order=inf is the kernel path used
    by the fields (fused with the grid normalisation there); other orders use the torch expression."""

    def __init__(self, order=None) -> None:
        super().__init__()
        self.order = order

    def forward(self, positions: Tensor) -> Tensor:
        mag = torch.linalg.norm(positions, ord=self.order, dim=-1)[..., None]
        return torch.where(mag < 1, positions, (2 - (1 / mag)) * (positions / mag))


class Embedding(FieldComponent):
    """field_components/embedding.py:27-55."""

    def __init__(self, in_dim: int, out_dim: int) -> None:
        super().__init__()
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.embedding = nn.Embedding(in_dim, out_dim)
        # accumulation target of d(embedding.weight) for the fused head backward (FlatGradBuffer.attach_sinks)
        self.grad_sink: Optional[Tensor] = None

    def mean(self, dim=0):
        return self.embedding.weight.mean(dim)

    def forward(self, in_tensor: Tensor) -> Tensor:
        return self.embedding(in_tensor)
