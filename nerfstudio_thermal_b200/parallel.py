"""Multi-GPU plumbing for the hot path: one process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch).

Replaces, for this path, the reference's only parallel strategy: `DistributedDataParallel(model,
find_unused_parameters=True)` in `nerfstudio/pipelines/base_pipeline.py:280-283` (process groups are created in
`nerfstudio/scripts/train.py:138-157`).

* Training: rays are independent, every rank holds full replicas; the single exchange step is the average of
  all parameter gradients.  DDP does it with ~7 buckets of 25 MB plus an autograd-graph walk for unused
  parameters; here all gradients live in ONE flat, zero-initialised fp32 buffer (parameters' `.grad` are views
  into it) and one all-reduce moves it.  Unused parameters (proposal networks on non-"updated" steps,
  ray_samplers.py:604-611) simply contribute zeros.
* Rendering: rows of the frame are sharded on chunk boundaries with no communication
  (`shard_chunks`); chunk boundaries are the reference's (`eval_num_rays_per_chunk`), so the chunk-global
  expected-depth clip (renderers.py:574) sees the same chunks.
"""
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor, nn


class FlatGradBuffer:
    """All gradients of a module in one contiguous fp32 buffer, laid out group by group."""

    def __init__(self, params: Iterable[nn.Parameter], device=None, group_of: Optional[Dict[int, str]] = None,
                 symmetric: bool = False):
        seen, self.params = set(), []
        for p in params:
            if p.requires_grad and p.numel() > 0 and id(p) not in seen:
                seen.add(id(p))
                self.params.append(p)
        device = device if device is not None else (self.params[0].device if self.params else "cpu")
        # 16-byte aligned offsets: the scatter kernels use vector reductions on table gradients
        self.offsets, total = [], 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        if symmetric:
            # symmetric (peer-mappable) allocation: the other ranks' copy engines read it directly (PeerExchange)
            import torch.distributed._symmetric_memory as symm_mem
            self.flat = symm_mem.empty(total, dtype=torch.float32, device=torch.device(device))
            self.flat.zero_()
        else:
            self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.symmetric = symmetric
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)
        # named groups = contiguous element ranges [begin, end) (alignment padding inside a group included)
        self.group_ranges: Dict[str, Tuple[int, int]] = {}
        self._group_members: Dict[str, List[Tuple[nn.Parameter, int]]] = {}
        if group_of:
            for p, off in zip(self.params, self.offsets):
                name = group_of[id(p)]
                b, _ = self.group_ranges.get(name, (off, off))
                self.group_ranges[name] = (b, off + (p.numel() + 3) // 4 * 4)
                self._group_members.setdefault(name, []).append((p, off))
        self.flat_params: Optional[Tensor] = None

    @classmethod
    def from_param_groups(cls, groups: Dict[str, List[nn.Parameter]], order: Optional[List[str]] = None, device=None,
                          symmetric: bool = False):
        names = order if order is not None else list(groups)
        group_of: Dict[int, str] = {}
        for n in names:
            for p in groups[n]:
                group_of.setdefault(id(p), n)  # a parameter listed twice belongs to its first group
        return cls([p for n in names for p in groups[n]], device=device, group_of=group_of, symmetric=symmetric)

    def group_params(self, name: str) -> List[Tuple[nn.Parameter, int]]:
        """(parameter, element offset) of every parameter of a named group, in buffer order."""
        return self._group_members.get(name, [])

    def flatten_params(self) -> Tensor:
        """Move every parameter into one flat fp32 buffer with the layout of the gradient buffer (`param.data`
        become views of it; values are kept).  Element i of the parameter, gradient and optimiser-state buffers
        then belongs to the same scalar, which is what the fused optimiser kernel needs."""
        if self.flat_params is None:
            flat = torch.zeros_like(self.flat)
            for p, off in zip(self.params, self.offsets):
                view = flat[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
            self.flat_params = flat
        return self.flat_params

    def attach_sinks(self, module: nn.Module) -> int:
        """Let every HashEncoding of `module` scatter its table gradient directly into this buffer (the
        backward kernel accumulates into the parameter's slice; no temporary, no extra add pass).  The caller
        must `zero_()` the buffer once per step.  Returns the number of tables attached."""
        from .field_components import MLP, Embedding, HashEncoding

        n = 0
        for m in module.modules():
            if isinstance(m, HashEncoding) and m.hash_table.grad is not None:
                m.grad_sink = m.hash_table.grad
                n += 1
            elif isinstance(m, MLP) and all(l.weight.grad is not None and l.bias.grad is not None for l in m.layers):
                # MLP backward kernels add dW/db straight into the buffer too (no temporaries, no accumulate pass)
                m.grad_sinks = [(l.weight.grad, l.bias.grad) for l in m.layers]
                n += 1
            elif isinstance(m, Embedding) and m.embedding.weight.grad is not None:
                m.grad_sink = m.embedding.weight.grad  # fused colour-head backward (fused_ops._FieldHeadFn)
                n += 1
            elif hasattr(m, "pose_adjustment") and hasattr(m, "grad_sink") and m.pose_adjustment.grad is not None:
                m.grad_sink = m.pose_adjustment.grad  # CameraOptimizer: ray-bundle and regulariser backward kernels
                n += 1
        return n

    def zero_(self) -> None:
        self.flat.zero_()

    def numel(self) -> int:
        return self.flat.numel()

    def check_views(self) -> bool:
        """True when every parameter's .grad still aliases the flat buffer (optimizers that set grads to None
        break the aliasing; call `zero_()` instead of `zero_grad(set_to_none=True)`)."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)

    def all_reduce_mean(self, group=None, async_op: bool = False, begin: int = 0, end: Optional[int] = None):
        """Average the gradients (elements [begin, end) of the buffer; default all) over the ranks of `group`
        (DDP semantics, base_pipeline.py:282)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        world = dist.get_world_size(group)
        part = self.flat if (begin == 0 and end is None) else self.flat[begin:end]
        if part.numel() == 0:
            return None
        if dist.get_backend(group) == "nccl":
            return dist.all_reduce(part, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        work = dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group, async_op=False)
        part.div_(world)
        return work


class PeerExchange:
    """Average elements [begin, end) of a SYMMETRIC flat gradient buffer over the ranks of one NVLink / NVSwitch
    domain without NCCL and (almost) without SMs:

        barrier | every rank PULLS its own shard of the segment from each peer (copy engines, peer memory)
                | one reduction kernel: shard = (own + pulled copies) / world            (tn_shard_mean)
        barrier | every rank pulls the other ranks' reduced shards into its buffer        (copy engines)
        barrier   (nobody may rewrite its buffer while a peer still reads it)

    An NCCL all-reduce of the same 134 MB keeps 16-32 CTAs busy for its whole duration and slows the issue-bound
    proposal kernels it is supposed to hide behind by 50-60 % (profiles/r02_exchange.md); here the bulk moves on the
    copy engines and the only kernel streams 1/world of the segment once.  Everything is enqueued on the current
    stream (barriers included: device-side, through the symmetric memory's signal pads), nothing blocks the host.
    Same result as DistributedDataParallel's averaged gradients (pipelines/base_pipeline.py:280-283) up to the
    summation order."""

    def __init__(self, flat: Tensor, begin: int, end: int, group=None, max_ctas: int = 32):
        import torch.distributed._symmetric_memory as symm_mem

        from ._lib import call, ptr, stream  # noqa: F401  (fails loudly without the CUDA library)

        group = group if group is not None else dist.group.WORLD
        self.handle = symm_mem.rendezvous(flat, group)  # collective
        self.flat, self.begin, self.end = flat, begin, end
        self.world, self.rank = self.handle.world_size, self.handle.rank
        n = end - begin
        assert begin % 4 == 0 and n % 4 == 0, "segment must be made of whole 16-byte units"
        self.shard = ((n + self.world - 1) // self.world + 3) // 4 * 4
        self.max_ctas = max_ctas
        self.staging = torch.empty((max(self.world - 1, 1), self.shard), dtype=torch.float32, device=flat.device)

    def _span(self, r: int) -> Tuple[int, int]:
        b = min(self.begin + r * self.shard, self.end)
        return b, min(b + self.shard, self.end)

    def all_reduce_mean(self) -> None:
        from ._lib import call, ptr, stream

        h, w, me = self.handle, self.world, self.rank
        b, e = self._span(me)
        h.barrier(channel=0)  # every rank's gradients of the segment are final
        for step in range(1, w):
            peer = (me - step) % w
            if e > b:
                self.staging[step - 1, :e - b].copy_(h.get_buffer(peer, (e - b,), torch.float32, b))
        if e > b:
            call("tn_shard_mean", ptr(self.flat) + 4 * b, ptr(self.staging), w - 1, self.shard, e - b, 1.0 / w,
                 self.max_ctas, stream(), tag="[exchange]", units=e - b)
        h.barrier(channel=1)  # every shard is reduced
        for step in range(1, w):
            peer = (me - step) % w
            pb, pe = self._span(peer)
            if pe > pb:
                self.flat[pb:pe].copy_(h.get_buffer(peer, (pe - pb,), torch.float32, pb))
        h.barrier(channel=2)  # every rank has read what it needs: buffers may be rewritten


def shard_chunks(num_rays: int, chunk: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """[start, end) ray ranges (whole chunks of the reference's chunk loop, models/base_model.py:187-190) that
    `rank` renders; contiguous blocks of chunks per rank, earlier ranks take the remainder."""
    n_chunks = (num_rays + chunk - 1) // chunk
    base, rem = divmod(n_chunks, world)
    first = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return [(c * chunk, min((c + 1) * chunk, num_rays)) for c in range(first, first + count)]


def rank_seed(base_seed: int, rank: int) -> int:
    """scripts/train.py:82-97: every rank seeds with config.machine.seed + global_rank."""
    return base_seed + rank
