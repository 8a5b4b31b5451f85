"""Step runners for the hot path: a CUDA-graph-captured train step and a sharded full-frame renderer.

The per-step work of thermal-nerfacto at 4096 rays is 76 launches (63 kernels of this library plus 13 small torch
kernels); issued eagerly the GPU idles between them.  On B200 the idiomatic fix is a
CUDA graph: the whole iteration is captured once on static buffers and replayed with one launch per step.
Everything on the path is capture-safe by construction -- the C ABI never allocates or synchronises, jitter comes
from torch's graph-registered Philox generator, the loss assembly avoids boolean indexing, the optimiser reads its
step counter from the device.

What one replay contains (`GraphedTrainStep._eager` is the captured program):

    [source: pixel sampling, collation, ray generation]            optional, `source=DeviceBatchSource(...)`
    gradient-buffer clear                                          on its own stream, beside the forward
    forward: RGB branch | thermal branch on a second stream        (model.sample_phase / render_phase), cross terms, losses
    backward: autograd replays every node on its forward stream
      '- when both main fields' gradients are final (hook): all-reduce of their 134 MB segment [N>1] and Adam on
         that segment [optimizer=], on an auxiliary stream beside the proposal networks' backward
    remaining all-reduce [N>1], Adam on the remaining segment [optimizer=]

Data parallel on several GPUs ("pipeline", the default with NCCL): the step is captured as THREE graphs cut where the
reference's dependencies allow an exchange to hide --

    A  pose corrections + proposal sampling of both branches        reads only proposal-network / camera parameters
    B  main fields, renderers, cross terms, losses, backward down to A's outputs     -> main fields' gradients final
       '- all-reduce of the main fields' 134 MB on the communication stream, started here ...
    C  backward of A (proposal networks, camera optimizers), all-reduce of the remaining 21 MB
    A' of the NEXT step                                              ... and hidden behind C and A'
    (wait for the exchange) B' ...

-- so the 134 MB exchange overlaps ~0.6 ms of work that does not depend on it, with exactly DDP's semantics: every
gradient is averaged before anything reads it, and a parameter is never read between its gradient's production and
its update (`optimizer=` steps the main fields on the communication stream right after their all-reduce).

Mirrors what `Trainer.train_iteration` + `VanillaPipeline.get_train_loss_dict` drive in the reference
(engine/trainer.py:456-500, pipelines/base_pipeline.py:291-304; DDP of base_pipeline.py:280-283; the optimiser and
scheduler steps of engine/optimizers.py:150-192).  Without `optimizer=` any optimiser can read the gradients from
`grads.flat` / `param.grad`.
"""
import os
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import parallel
from .model import ThermalNerfactoModel
from .optim import AdamGroupConfig, FusedAdam
from .rays import RayBundle

BATCH_KEYS = ("origins", "directions", "pixel_area", "camera_indices", "image", "is_thermal")


class DeviceBatchSource:
    """A train batch made on the device: PatchPixelSampler over the cached images in HBM + RayGenerator.
    Given to GraphedTrainStep as `source=`, the sampling and the ray generation are captured in the step's CUDA
    graph (torch.rand draws from the graph-registered Philox state), so a step has no host input at all.
    Mirrors VanillaDataManager.next_train (data/datamanagers/base_datamanager.py:  pixel sampler -> ray generator)."""

    def __init__(self, sampler, ray_generator, image_batch: Dict[str, Tensor]):
        self.sampler, self.ray_generator, self.image_batch = sampler, ray_generator, image_batch

    def next(self) -> Dict[str, Tensor]:
        b = self.sampler.sample(self.image_batch)
        rays = self.ray_generator(b["indices"])
        return {"origins": rays.origins, "directions": rays.directions, "pixel_area": rays.pixel_area,
                "camera_indices": rays.camera_indices, "image": b["image"], "is_thermal": b["is_thermal"]}


_WORK_STREAMS: Dict[str, "torch.cuda.Stream"] = {}


def _work_stream(device) -> "torch.cuda.Stream":
    """The one stream per device on which every GraphedTrainStep warms up and captures (see GraphedTrainStep.__init__)."""
    k = str(device)
    if k not in _WORK_STREAMS:
        _WORK_STREAMS[k] = torch.cuda.Stream(device=device)
    return _WORK_STREAMS[k]


class GraphedTrainStep:
    """forward + metrics + loss dict + backward (+ flat gradient all-reduce) for a fixed batch shape.

    step(batch) copies the batch (host or device tensors) into static device buffers, replays the captured
    graph and returns the scalar total loss tensor (device).  `losses` holds the per-term loss tensors of the
    last step; gradients are in `grads.flat` (parameters' .grad are views into it).
    """

    def __init__(self, model: ThermalNerfactoModel, example_batch: Dict[str, Tensor], use_graph: bool = True,
                 warmup: int = 3, group=None, optimizer: Optional[Dict[str, AdamGroupConfig]] = None,
                 overlap_comm: Optional[bool] = None, source: Optional[DeviceBatchSource] = None,
                 pipeline: Optional[bool] = None):
        self.model = model
        self.source = source  # batches come from the device-side sampler / ray generator instead of step(batch)
        self.device = model.device
        assert self.device.type == "cuda", "the hot path runs on CUDA only (no CPU fallback)"
        self.group = group
        world = torch.distributed.get_world_size(group) if torch.distributed.is_initialized() else 1
        groups = model.get_param_groups()
        # Data parallel.  "after" (TN_COMM=after): ONE all-reduce of the flat buffer after the graph replay.
        # "overlap": the main fields' groups (2 x 64 MB tables + their MLPs, 83 % of the buffer)
        # sit at the FRONT of the flat buffer; their gradients are final as soon as both fields' encode backward
        # kernels have run, well before the proposal networks' backward (issue-bound kernels that leave HBM and
        # NVLink idle).  A backward hook per field records an event; when both have fired a communication stream
        # waits for them and starts the all-reduce of that segment, the rest follows after the last backward kernel,
        # and both collectives (and the optimiser) are captured inside the CUDA graph.
        early = [n for n in ("fields", "fields_thermal") if n in groups]
        order = early + [n for n in groups if n not in early]
        nccl = world > 1 and torch.distributed.get_backend(group) == "nccl"
        # measured at the end of round 2 (profiles/r02_exchange/summary.txt, ms/step at 2 / 4 / 8 GPUs; one GPU: 2.10):
        # "after" 2.42 / - / -, "overlap" 2.35 / 2.39 / 4.60, "pipeline" 2.30 / 2.39 / 2.43 -- the three-phase pipeline
        # with the peer-memory exchange is never behind (its end-to-end time is the lowest at every size)
        default = "pipeline"
        mode = os.environ.get("TN_COMM", default if overlap_comm is None else ("overlap" if overlap_comm else "after"))
        # the main fields' exchange over NVLink peer memory (parallel.PeerExchange) needs a peer-mappable buffer
        want_peer = nccl and mode == "pipeline" and use_graph and os.environ.get("TN_PEER_EXCHANGE", "1") == "1"
        self.grads = parallel.FlatGradBuffer.from_param_groups(groups, order=order, device=self.device,
                                                               symmetric=want_peer)
        self.grads.attach_sinks(model)
        self._early_end = max((self.grads.group_ranges[n][1] for n in early), default=0)
        if mode not in ("pipeline", "overlap", "after"):
            raise ValueError(f"TN_COMM={mode!r}: expected pipeline | overlap | after")
        # "pipeline": three graphs per step, the main fields' all-reduce hidden behind the proposal backward and the
        # NEXT step's proposal forward (module docstring).  "overlap" (round 1): one graph, the early all-reduce
        # captured inside it -- wins on 2 and 4 GPUs, loses on 8, where the captured collective's spin-waits on late
        # peers contend with the proposal backward (3.83 ms vs 3.20 with the trailing all-reduce, "after").
        # (`pipeline=True` runs the three-phase schedule on one GPU as well: the tests compare it with the single graph)
        want_pipeline = (nccl and mode == "pipeline") if pipeline is None else pipeline
        self._pipeline = (want_pipeline and use_graph and self._early_end > 0 and getattr(model, "fuse_levels", False)
                          and source is None)
        self._comm_in_graph = nccl and mode == "overlap"
        self._comm_event: Optional[torch.cuda.Event] = None
        self._early_done = False
        self._adam_now = False
        self._zero_in_adam = False
        self._zero_stream: Optional[torch.cuda.Stream] = None
        self._ready_events: List[torch.cuda.Event] = []
        self._fields = [f for f in (getattr(model, "field", None), getattr(model, "field_thermal", None))
                        if f is not None and any(p.requires_grad for p in f.parameters())]
        # parameters move into their flat buffer BEFORE anything is captured (the graph bakes in addresses)
        self.optimizer = FusedAdam(self.grads, optimizer) if optimizer is not None else None
        self._adam_in_graph = self.optimizer is not None and (world == 1 or self._comm_in_graph)
        # the fields' segment is exchanged AND stepped early (HBM/NVLink-bound work beside the issue-bound proposal
        # backward): both hang off the same backward hook and run on the auxiliary stream
        if self._pipeline:
            self._adam_in_graph = False
        self._early_active = self._early_end > 0 and (self._comm_in_graph or self._adam_in_graph) and not self._pipeline
        if self._early_active or self._pipeline:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        # "pipeline": the main fields' all-reduce gets its OWN communicator.  A process group issues all its
        # collectives on one internal stream, in order: on a shared communicator the small trailing all-reduce (which
        # the main stream waits for) would queue behind the big one and put it back on the critical path.
        self._main_group = group
        self._peer: Optional[parallel.PeerExchange] = None
        if self._pipeline and nccl and self.grads.symmetric:
            # copy-engine pulls over peer memory + one small reduction kernel instead of an NCCL all-reduce whose
            # CTAs would compete with the proposal kernels it hides behind
            self._peer = parallel.PeerExchange(self.grads.flat, 0, self._early_end, group)
        elif self._pipeline and nccl:
            ranks = torch.distributed.get_process_group_ranks(group) if group is not None else None
            self._main_group = torch.distributed.new_group(ranks=ranks, backend="nccl")
        for f in self._fields:  # (also detaches the hooks of an earlier runner of the same model)
            f.grads_ready_callback = self._field_ready if self._early_active else None
        # the step's inputs live in ONE device buffer (16-byte aligned slices): a host batch laid out the same way
        # (`host_batch()`) reaches the device with a single copy instead of one per tensor
        self._static_layout, nbytes = [], 0
        for k in BATCH_KEYS:
            t = example_batch[k]
            self._static_layout.append((k, nbytes, t.dtype, tuple(t.shape)))
            nbytes += (t.numel() * t.element_size() + 15) // 16 * 16
        self._static_bytes = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.static = self._views(self._static_bytes)
        self._seed = torch.ones((), device=self.device)
        # Warm-up passes and every capture run on ONE stream.  autograd's AccumulateGrad nodes remember the stream they
        # were created on and outlive a step (the model and this runner keep parts of the last graph alive); a
        # parameter whose gradient goes to a sink hands its node an undefined gradient, which synchronises nothing, yet
        # the engine still joins that node's stream at the end of the backward -- on a stream that is not part of the
        # capture that is cudaErrorStreamCaptureIsolation.
        self._work_stream = _work_stream(self.device) if use_graph else None
        self._load(example_batch)
        self.losses: Dict[str, Tensor] = {}
        self.total: Optional[Tensor] = None
        self.use_graph = use_graph
        # Host-side schedule state that a captured graph would otherwise freeze (ray_samplers.py:591, 604-613): whether
        # each proposal sampler's networks are "updated" (trained) this iteration.  One graph VARIANT is captured per
        # decision tuple, lazily; step() picks the variant on the host and applies the samplers' bookkeeping after the
        # replay.  (The annealing exponent is not frozen either: the PDF kernel reads it from a device scalar that
        # ProposalNetworkSampler.set_anneal rewrites.)
        self._samplers = [model.proposal_sampler]
        self._sampler_groups = ["proposal_networks"]
        if model.config.density_mode == "separate":
            self._samplers.append(model.proposal_sampler_thermal)
            self._sampler_groups.append("proposal_networks_thermal")
        self._variants: Dict[Tuple[bool, ...], Tuple[torch.cuda.CUDAGraph, Tensor, Dict[str, Tensor]]] = {}
        self.graph: Optional[torch.cuda.CUDAGraph] = None  # the variant replayed last
        if self.optimizer is not None or use_graph:
            from .field_components import HashEncoding
            if any(isinstance(m, HashEncoding) and m.use_half_table for m in model.modules()):
                # the fp16 gather mirrors are refreshed from Python when the parameter's version changes; neither the
                # fused optimiser (raw pointers) nor a graph replay would ever trigger that
                raise NotImplementedError("fp16 gather tables (use_half_table) are an inference option: not supported "
                                          "together with the fused optimiser or a captured train step")
        if world > 1:
            self.sync_replicas()
        model.train()
        if use_graph:
            key = self._variant_key()
            side = self._work_stream
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(max(warmup, 2) - 1):  # allocator / cudaFuncSetAttribute / host-constant caches warm
                    if self._pipeline:
                        self._phases(key)
                    else:
                        self._eager(apply_optimizer=False, key=key)  # warm-up must not move the parameters
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            self._capture(key)

    def _drop_old_graphs(self) -> None:
        """Release what this runner and the model still hold of earlier autograd graphs (and with them AccumulateGrad
        nodes made on other streams, see `_work_stream`)."""
        self.total, self.losses = None, {}
        for m in self.model.modules():
            for name in ("_sample_locations", "_density_before_activation", "_reg_cache"):
                if getattr(m, name, None) is not None:
                    setattr(m, name, None)

    def sync_replicas(self) -> None:
        """Data-parallel replicas must start identical: DistributedDataParallel broadcasts rank 0's parameters and
        buffers when it wraps the model (pipelines/base_pipeline.py:280-283); this is the same step for the flat
        buffers (call it again after loading a checkpoint on one rank)."""
        if not torch.distributed.is_initialized() or torch.distributed.get_world_size(self.group) == 1:
            return
        bc = lambda t: torch.distributed.broadcast(t, src=torch.distributed.get_global_rank(self.group, 0)  # noqa: E731
                                                   if self.group is not None else 0, group=self.group)
        with torch.no_grad():
            if self.grads.flat_params is not None:
                bc(self.grads.flat_params)
            else:
                for p in self.grads.params:
                    bc(p.data)
            for b in self.model.buffers():
                if b.numel() > 0:
                    bc(b)
            if self.optimizer is not None:
                for t in (self.optimizer.exp_avg, self.optimizer.exp_avg_sq, self.optimizer.step_dev):
                    bc(t)

    def _variant_key(self) -> Tuple[bool, ...]:
        return tuple(s.will_update() for s in self._samplers)

    def _capture(self, key: Tuple[bool, ...]):
        """Capture the train step for one tuple of "updated" decisions (one extra eager pass first: a variant met for
        the first time mid-training may run kernels that have not been launched yet)."""
        side = self._work_stream
        self._drop_old_graphs()
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            if self._pipeline:
                self._phases(key)
            else:
                self._eager(apply_optimizer=False, key=key)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.grads.zero_()
        first = next(iter(self._variants.values()))[0] if self._variants else None
        pool = (first[0] if isinstance(first, tuple) else first).pool() if first is not None else None  # serial replays
        if self._pipeline:
            graphs = []
            for phase in (self._phase_a, self._phase_b, self._phase_c):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool, stream=side):
                    phase(key)
                pool = g.pool()
                graphs.append(g)
            graph = tuple(graphs)
        else:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, pool=pool, stream=side):
                self._eager(apply_optimizer=True, captured=True, key=key)
        torch.cuda.synchronize(self.device)
        self.graph = graph
        self._variants[key] = (graph, self.total, self.losses)
        return self._variants[key]

    # ---- "pipeline": the step in three phases (module docstring)
    def _phase_a(self, key: Tuple[bool, ...]) -> None:
        """Pose corrections + proposal sampling of both branches; its differentiable outputs become leaves."""
        s = self.static
        bundle = self.model.collider(RayBundle(origins=s["origins"], directions=s["directions"],
                                               pixel_area=s["pixel_area"], camera_indices=s["camera_indices"]))
        for smp, upd in zip(self._samplers, key):
            smp._forced_updated = upd
        try:
            self._state = self.model.sample_phase(bundle)
        finally:
            for smp in self._samplers:
                smp._forced_updated = None
        self._pairs = self._state.cut()
        if self._state.streams:  # a phase ends with every stream joined (graph capture requires it)
            torch.cuda.current_stream(self.device).wait_stream(self.model._side_stream)

    def _phase_b(self, key: Tuple[bool, ...]) -> None:
        """Gradient clear, main fields + renderers + cross terms + losses, backward down to phase A's outputs."""
        s = self.static
        cur = torch.cuda.current_stream(self.device)
        if self._zero_stream is None:
            self._zero_stream = torch.cuda.Stream(device=self.device)
        self._zero_stream.wait_stream(cur)
        with torch.cuda.stream(self._zero_stream):
            self.grads.zero_()
        batch = {"image": s["image"], "is_thermal": s["is_thermal"]}
        outputs = self.model.render_phase(self._state)
        metrics = self.model.get_metrics_dict(outputs, batch)
        self.losses = self.model.get_loss_dict(outputs, batch, metrics)
        total = getattr(self.losses, "total", None)
        self.total = total if total is not None else sum(self.losses.values())
        cur.wait_stream(self._zero_stream)
        self.total.backward(gradient=self._seed)  # cached 1.0: no ones_like launch per step

    def _phase_c(self, key: Tuple[bool, ...]) -> None:
        """Backward of phase A from the gradients phase B left on its leaves: proposal networks, camera optimizers."""
        live = [(t, leaf.grad) for t, leaf in self._pairs if leaf.grad is not None]
        if live:
            torch.autograd.backward([t for t, _ in live], [g for _, g in live])

    def _phases(self, key: Tuple[bool, ...]) -> None:
        self._phase_a(key)
        self._phase_b(key)
        self._phase_c(key)

    def finish_exchange(self) -> None:
        """Make the current stream wait for the gradient exchange of the last step ("pipeline" leaves the main
        fields' all-reduce running on the communication stream when step() returns: call this before reading
        `grads.flat` / `param.grad` on the current stream; the next step() does it at the right place itself)."""
        if self._comm_event is not None:
            torch.cuda.current_stream(self.device).wait_event(self._comm_event)

    def _field_ready(self) -> None:
        """Backward hook of a main field (fires on the stream that ran its encode backward).  Once every field has
        reported, the communication stream waits for those points of both streams and all-reduces the fields'
        segment of the buffer, asynchronously w.r.t. the rest of the backward."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._ready_events.append(ev)
        if len(self._ready_events) < len(self._fields):
            return
        for e in self._ready_events:
            self._comm_stream.wait_event(e)
        with torch.cuda.stream(self._comm_stream):
            if self._comm_in_graph:
                work = self.grads.all_reduce_mean(self.group, async_op=True, begin=0, end=self._early_end)
                if work is not None:
                    work.wait()  # the auxiliary stream waits for the collective
            if self._adam_now:
                self.optimizer.step_range(0, self._early_end, zero_grads=self._zero_in_adam)
        self._early_done = True

    def _views(self, buf: Tensor) -> Dict[str, Tensor]:
        out = {}
        for k, off, dtype, shape in self._static_layout:
            n = int(torch.tensor(shape).prod().item()) if len(shape) else 1
            out[k] = buf[off:off + n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(shape)
        return out

    def host_batch(self) -> Dict[str, Tensor]:
        """Pinned host tensors (one per batch key, same shapes and dtypes as the example batch) that are views of ONE
        pinned buffer laid out like the device-side input buffer: fill them (a data loader collates into them) and
        pass the dict to `step()` -- the whole batch then moves with a single host-to-device copy."""
        buf = torch.empty(self._static_bytes.numel(), dtype=torch.uint8).pin_memory()
        views = self._views(buf)
        views["_packed"] = buf
        return views

    def _load(self, batch: Dict[str, Tensor]) -> None:
        packed = batch.get("_packed")
        if packed is not None and packed.numel() == self._static_bytes.numel() and packed.dtype == torch.uint8:
            self._static_bytes.copy_(packed, non_blocking=True)
            return
        for k in BATCH_KEYS:
            self.static[k].copy_(batch[k], non_blocking=True)

    def _eager(self, apply_optimizer: bool = True, captured: bool = False,
               key: Optional[Tuple[bool, ...]] = None) -> None:
        if key is None:
            key = self._variant_key()
        s = self.static if self.source is None else self.source.next()
        cur = torch.cuda.current_stream(self.device)
        zeroed = None
        if not (captured and self._adam_in_graph):  # the in-graph Adam pass leaves the gradients cleared
            # the 155 MB clear is only needed by the backward: it runs beside the forward on its own stream
            if self._zero_stream is None:
                self._zero_stream = torch.cuda.Stream(device=self.device)
            zeroed = self._zero_stream
            zeroed.wait_stream(cur)
            with torch.cuda.stream(zeroed):
                self.grads.zero_()
        bundle = RayBundle(origins=s["origins"], directions=s["directions"], pixel_area=s["pixel_area"],
                           camera_indices=s["camera_indices"])
        for smp, upd in zip(self._samplers, key):
            smp._forced_updated = upd
        try:
            _, self.losses, _ = self.model.get_train_loss_dict(bundle, {"image": s["image"],
                                                                        "is_thermal": s["is_thermal"]})
        finally:
            for smp in self._samplers:
                smp._forced_updated = None
        total = getattr(self.losses, "total", None)
        self.total = total if total is not None else sum(self.losses.values())
        self._early_done, self._ready_events = False, []
        self._adam_now = apply_optimizer and self._adam_in_graph
        self._zero_in_adam = captured
        if self._adam_now:
            # proposal networks that are not updated this iteration have no gradient in the reference: their
            # optimiser is skipped (engine/optimizers.py:165-170)
            self.optimizer.tick(inactive=[g for g, upd in zip(self._sampler_groups, key) if not upd])
        if zeroed is not None:
            cur.wait_stream(zeroed)
        self.total.backward(gradient=self._seed)  # cached 1.0: no ones_like launch per step
        begin = self._early_end if self._early_done else 0
        if self._comm_in_graph:
            self.grads.all_reduce_mean(self.group, begin=begin)
        if self._adam_now:
            self.optimizer.step_range(begin, self.grads.flat.numel(), zero_grads=captured)
        if self._early_done:
            cur.wait_stream(self._comm_stream)

    def step(self, batch: Optional[Dict[str, Tensor]] = None) -> Tensor:
        if batch is not None:
            self._load(batch)
        key = self._variant_key()
        if self._pipeline:
            self._step_pipeline(key)
        elif self.use_graph:
            variant = self._variants.get(key)
            if variant is None:
                variant = self._capture(key)
            self.graph, self.total, self.losses = variant
            self.graph.replay()
        else:
            self._eager(key=key)
        for smp, upd in zip(self._samplers, key):  # what generate_ray_samples does at its end
            smp.mark_sampled(upd)
        if self._pipeline:
            return self.total
        if not self._comm_in_graph:
            self.grads.all_reduce_mean(self.group)
        if self.optimizer is not None and not self._adam_in_graph:
            self.optimizer.step(inactive=[g for g, upd in zip(self._sampler_groups, key) if not upd])
        return self.total

    def _step_pipeline(self, key: Tuple[bool, ...]) -> None:
        variant = self._variants.get(key)
        if variant is None:
            self.finish_exchange()  # a capture synchronises the device anyway
            variant = self._capture(key)
        (g_a, g_b, g_c), self.total, self.losses = variant
        self.graph = variant[0]
        main, comm = torch.cuda.current_stream(self.device), self._comm_stream
        inactive = [g for g, upd in zip(self._sampler_groups, key) if not upd]
        if self.optimizer is not None:
            self.optimizer.tick(inactive=inactive)
        g_a.replay()
        self.finish_exchange()  # B clears and rewrites the gradient buffer (and reads the main fields' parameters)
        g_b.replay()
        ready = torch.cuda.Event()
        ready.record(main)
        comm.wait_event(ready)
        with torch.cuda.stream(comm):  # the main fields' segment: hidden behind C and the next step's A
            if self._peer is not None:
                self._peer.all_reduce_mean()
            else:
                self.grads.all_reduce_mean(self._main_group, begin=0, end=self._early_end)
            if self.optimizer is not None:
                self.optimizer.step_range(0, self._early_end, zero_grads=False)
            self._comm_event = torch.cuda.Event()
            self._comm_event.record(comm)
        g_c.replay()
        self.grads.all_reduce_mean(self.group, begin=self._early_end)
        if self.optimizer is not None:
            self.optimizer.step_range(self._early_end, self.grads.flat.numel(), zero_grads=False)

    def train_iteration(self, step: int, batch: Optional[Dict[str, Tensor]] = None) -> Tensor:
        """Trainer.train_iteration with its callbacks (engine/trainer.py:262-275, 456-500): the
        BEFORE_TRAIN_ITERATION hooks (proposal-weight annealing), the step, the AFTER_TRAIN_ITERATION hooks (the
        samplers' update counters)."""
        self.model.set_anneal_step(step)
        total = self.step(batch)
        self.model.step_cb(step)
        return total


class GraphedRenderChunk:
    """The eval forward of one full chunk of rays (`eval_num_rays_per_chunk`, models/base_model.py:187-190) captured
    as a CUDA graph: ~100 launches become one, and the thermal branch runs on a second stream beside the RGB one
    (`model.eval_branch_streams`, as in training).  `render(bundle)` copies a chunk into the static input buffers,
    replays, and returns the output dict (views of static buffers: consume or clone before the next call).  A
    shorter last chunk is padded with copies of its first ray (duplicates leave the chunk-global extrema of
    renderers.py:574 unchanged) and the outputs are sliced."""

    def __init__(self, model: ThermalNerfactoModel, chunk: Optional[int] = None, warmup: int = 2):
        self.model, self.device = model, model.device
        assert self.device.type == "cuda", "the hot path runs on CUDA only (no CPU fallback)"
        self.chunk = chunk or model.config.eval_num_rays_per_chunk
        c = self.chunk
        self.static = RayBundle(origins=torch.zeros((c, 3), device=self.device),
                                directions=torch.zeros((c, 3), device=self.device),
                                pixel_area=torch.ones((c, 1), device=self.device),
                                camera_indices=torch.zeros((c, 1), dtype=torch.long, device=self.device))
        self.static.directions[:, 2] = 1.0
        model.eval()
        streams, model.eval_branch_streams = model.eval_branch_streams, True
        try:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(max(warmup, 1)):
                    model(self._bundle())
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph), torch.no_grad():
                self.outputs = {k: v for k, v in model(self._bundle()).items() if torch.is_tensor(v)}
            torch.cuda.synchronize(self.device)
        finally:
            model.eval_branch_streams = streams

    def _bundle(self) -> RayBundle:
        s = self.static  # a fresh bundle object per call: the model rewrites origins/directions/nears/fars on it
        return RayBundle(origins=s.origins, directions=s.directions, pixel_area=s.pixel_area,
                         camera_indices=s.camera_indices)

    @torch.no_grad()
    def render(self, bundle: RayBundle) -> Dict[str, Tensor]:
        n = bundle.origins.shape[0]
        if n > self.chunk or n == 0:
            raise ValueError(f"chunk of {n} rays (captured for up to {self.chunk})")
        s = self.static
        for dst, src in ((s.origins, bundle.origins), (s.directions, bundle.directions),
                         (s.pixel_area, bundle.pixel_area), (s.camera_indices, bundle.camera_indices)):
            dst[:n].copy_(src, non_blocking=True)
            if n < self.chunk:
                dst[n:].copy_(dst[:1].expand(self.chunk - n, -1))
        self.graph.replay()
        return self.outputs if n == self.chunk else {k: v[:n] for k, v in self.outputs.items()}


@torch.no_grad()
def render_rays_sharded(model: ThermalNerfactoModel, bundle: RayBundle, rank: int = 0, world: int = 1,
                        keys: Optional[List[str]] = None, use_graph: bool = False) -> Dict[str, Tensor]:
    """Zero-communication full-frame render: this rank evaluates its contiguous block of the reference's chunks
    (models/base_model.py:177-206) of a flattened ray bundle and returns the outputs for those rays only
    (the caller concatenates rank outputs in rank order).  use_graph: chunks replay one captured graph
    (GraphedRenderChunk, cached on the model) -- pays off for small chunks; at the reference's 32768 rays per chunk the
    kernels already fill the GPU and the eager loop is as fast."""
    model.eval()
    flat = bundle.flatten()
    chunk = model.config.eval_num_rays_per_chunk
    runner = None
    if use_graph and model.device.type == "cuda":
        runner = getattr(model, "_render_chunk_runner", None)
        if runner is None or runner.chunk != chunk:
            runner = model._render_chunk_runner = GraphedRenderChunk(model, chunk)
    outs: Dict[str, List[Tensor]] = {}
    for start, end in parallel.shard_chunks(len(flat), chunk, rank, world):
        part = flat[start:end].to(model.device)
        res = runner.render(part) if runner is not None else model(part)
        for k, v in res.items():
            if torch.is_tensor(v) and (keys is None or k in keys):
                outs.setdefault(k, []).append(v.clone() if runner is not None else v)
    return {k: torch.cat(v) for k, v in outs.items()}
