"""Oracle: multiresolution hash grid (torch branch of the reference).  TEST INFRASTRUCTURE ONLY.

Follows field_components/encodings.py:324-379 (constructor: scalings, offsets) and :401-461
(`hash_fn`, `pytorch_fwd`).  Note this is *not* tiny-cuda-nn's grid: every level is hashed, every
level has 2^log2T rows, scales are floor(min_res * g^l) in float32, corners are ceil/floor and the
interpolation weight `offset` sits on the CEIL corner.
"""
import numpy as np
import torch

# hash primes, encodings.py:413
_PRIMES = (1, 2654435761, 805459861)

# corner c uses ceil (1) or floor (0) on each axis; order = hashed_0..hashed_7, encodings.py:431-438
CORNER_USES_CEIL = (
    (1, 1, 1),  # 0
    (1, 0, 1),  # 1
    (0, 0, 1),  # 2
    (0, 1, 1),  # 3
    (1, 1, 0),  # 4
    (1, 0, 0),  # 5
    (0, 0, 0),  # 6
    (0, 1, 0),  # 7
)


def hash_scalings(num_levels: int, min_res: int, max_res: int) -> torch.Tensor:
    """Per-level scale, encodings.py:343-345 (np.float64 growth factor, float32 floor)."""
    levels = torch.arange(num_levels)
    growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1
    return torch.floor(min_res * growth**levels)


def _hash(coords: torch.Tensor, table_size: int, level_offset: torch.Tensor) -> torch.Tensor:
    """encodings.py:401-418: int32 coords * int64 primes, xor, mod T, + l*T."""
    prod = coords * torch.tensor(_PRIMES, device=coords.device)
    h = torch.bitwise_xor(torch.bitwise_xor(prod[..., 0], prod[..., 1]), prod[..., 2])
    h = h % table_size
    return h + level_offset


def hash_corner_indices(x: torch.Tensor, scalings: torch.Tensor, log2_table_size: int):
    """Returns (idx[N,L,8] int64 rows into the [L*T,F] table, offset[N,L,3] float32).

    encodings.py:424-438.
    """
    num_levels = scalings.numel()
    table_size = 2**log2_table_size
    scaled = x[..., None, :] * scalings.view(-1, 1)
    hi = torch.ceil(scaled).type(torch.int32)
    lo = torch.floor(scaled).type(torch.int32)
    offset = scaled - lo
    level_offset = torch.arange(num_levels, device=x.device) * table_size
    idx = []
    for use_ceil in CORNER_USES_CEIL:
        coords = torch.stack([hi[..., a] if use_ceil[a] else lo[..., a] for a in range(3)], dim=-1)
        idx.append(_hash(coords, table_size, level_offset))
    return torch.stack(idx, dim=-1), offset


def hash_encode(x: torch.Tensor, table: torch.Tensor, scalings: torch.Tensor, log2_table_size: int) -> torch.Tensor:
    """x[N,3] in [0,1] -> [N, L*F], level-major.  encodings.py:440-461 (same op order)."""
    idx, offset = hash_corner_indices(x, scalings, log2_table_size)
    f = [table[idx[..., c]] for c in range(8)]  # each [N, L, F]
    ox, oy, oz = offset[..., 0:1], offset[..., 1:2], offset[..., 2:3]
    f03 = f[0] * ox + f[3] * (1 - ox)
    f12 = f[1] * ox + f[2] * (1 - ox)
    f56 = f[5] * ox + f[6] * (1 - ox)
    f47 = f[4] * ox + f[7] * (1 - ox)
    f0312 = f03 * oy + f12 * (1 - oy)
    f4756 = f47 * oy + f56 * (1 - oy)
    enc = f0312 * oz + f4756 * (1 - oz)
    return torch.flatten(enc, start_dim=-2, end_dim=-1)
