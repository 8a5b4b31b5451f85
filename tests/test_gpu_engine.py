"""GPU tests of the step runner (engine.GraphedTrainStep) and of the component boundary.

* a captured train step follows the reference's host-side schedules (proposal-weight annealing, the
  "updated"/no_grad schedule of the proposal networks, per-group Adam skipping) -- ADVICE r1 (high);
* reference-built (`_layout`-less, duck-typed) RaySamples run through the components;
* world_size-2 NCCL: averaged gradients of the overlapped and the trailing exchange equal the single-GPU gradients
  of the concatenated batch (pipelines/base_pipeline.py:280-283 semantics) -- needs two GPUs.
"""
import os
import socket

import numpy as np
import pytest
import torch

import oracle

import parity_utils as pu

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import nerfstudio_thermal_b200 as tn
    from nerfstudio_thermal_b200 import engine, optim as poptim

DEV = "cuda"
SMALL_PROPS = [{"hidden_dim": 16, "log2_hashmap_size": 8, "num_levels": 5, "max_res": m, "use_linear": False}
               for m in (128, 256)]


def _small(seed=3, rays=64, **over):
    torch.manual_seed(seed)
    cfg = tn.ThermalNerfactoModelConfig(density_mode="separate", log2_hashmap_size=9, proposal_net_args_list=SMALL_PROPS,
                                        **over)
    model = cfg.setup(num_train_data=8, metadata={"is_thermal": [0] * 4 + [1] * 4})
    with torch.no_grad():
        for k, p in model.named_parameters():
            if k.endswith("hash_table"):
                p.mul_(100.0)
            if "pose_adjustment" in k:
                p.normal_(0, 1e-3)
    import bench
    batch = {k: v.to(DEV) for k, v in bench.make_batch(rays, 5, num_cams=8).items()}
    ocfg = oracle.OracleConfig(density_mode="separate", log2_hashmap_size=9, is_thermal_cameras=(0, 0, 0, 0, 1, 1, 1, 1),
                               proposal_net_args_list=[{k: v for k, v in a.items() if k != "use_linear"}
                                                       for a in SMALL_PROPS])
    return model.to(DEV).train(), batch, ocfg


def _no_jitter(model):
    """deterministic sampling in train mode (bin centres), so graph replays are comparable with the oracle"""
    for s in (model.proposal_sampler, model.proposal_sampler_thermal):
        s.initial_sampler.train_stratified = False
        s.pdf_sampler.train_stratified = False


@pytest.mark.parametrize("use_graph", [True, False])
def test_captured_step_follows_anneal_and_update_schedules(use_graph):
    """Trainer.train_iteration semantics through a replayed graph: at every iteration the loss dict equals the
    oracle's for THAT iteration's annealing exponent (models/nerfacto.py:271-281) and "updated" decision
    (ray_samplers.py:591); proposal networks do not move on iterations where they are not updated, and their Adam
    step count lags accordingly (torch.optim.Adam skips grad=None parameters)."""
    model, batch, ocfg = _small(proposal_warmup=40, proposal_update_every=3, proposal_weights_anneal_max_num_iters=20)
    _no_jitter(model)
    runner = engine.GraphedTrainStep(model, batch, use_graph=use_graph, optimizer=poptim.thermal_nerfacto_optimizers())
    c = model.config
    ssu, expected_updates, seen_not_updated = 0, 0, 0
    prop_params = [p for p in model.proposal_networks.parameters()]
    none_jit = [None] * 6
    for step in range(18):
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        before = [p.detach().clone() for p in prop_params]
        # the reference's schedule, restated on the host
        frac = np.clip(step / c.proposal_weights_anneal_max_num_iters, 0, 1)
        b = c.proposal_weights_anneal_slope
        anneal = b * frac / ((b - 1) * frac + 1)
        # (step_cb runs AFTER the iteration: during iteration `step` the sampler still holds the previous step)
        seen = max(step - 1, 0)
        sched = np.clip(np.interp(seen, [0, c.proposal_warmup], [0, c.proposal_update_every]), 1, c.proposal_update_every)
        updated = ssu > sched or seen < 10
        total = runner.train_iteration(step, batch)
        with torch.no_grad():
            # thermal sampler: never driven by callbacks in the default config -> anneal 1, always updated
            ref = oracle.thermal_nerfacto_forward(sd, ocfg, batch["origins"], batch["directions"],
                                                  batch["camera_indices"], training=True, jitters=none_jit[:3],
                                                  jitters_thermal=none_jit[3:], anneal=float(anneal), updated=updated)
            # (the oracle applies one anneal to both samplers; the thermal one must stay at 1.0)
            ref_t = oracle.thermal_nerfacto_forward(sd, ocfg, batch["origins"], batch["directions"],
                                                    batch["camera_indices"], training=True, jitters=none_jit[:3],
                                                    jitters_thermal=none_jit[3:], anneal=1.0, updated=True)
        got = runner.losses
        # rgb branch outputs depend on the annealed RGB sampler only; thermal branch on the un-annealed thermal one
        ref_losses = oracle.thermal_nerfacto_losses(sd, ocfg, _merge(ref, ref_t), batch["image"], batch["is_thermal"])
        for k in ref_losses:
            a, bb = float(got[k]), float(ref_losses[k])
            assert abs(a - bb) <= 2e-3 * max(abs(bb), 1e-6), (step, k, a, bb, anneal, updated)
        assert abs(float(total) - float(sum(ref_losses.values()))) <= 2e-3 * abs(float(sum(ref_losses.values())))
        moved = any(not torch.equal(p.detach(), q) for p, q in zip(prop_params, before))
        assert moved == updated, (step, moved, updated)
        expected_updates += int(updated)
        seen_not_updated += int(not updated)
        # ProposalNetworkSampler bookkeeping (ray_samplers.py:612-613, step_cb)
        if updated:
            ssu = 0
        ssu += 1
    assert seen_not_updated >= 2
    counts = runner.optimizer.group_step_counts()
    assert counts["proposal_networks"] == expected_updates
    assert counts["proposal_networks_thermal"] == 18 and counts["fields"] == 18
    assert runner.optimizer.step_count == 18
    if use_graph:
        assert len(runner._variants) == 2


def _merge(ref_rgb, ref_thermal):
    """RGB-branch entries from the run with the annealed sampler, thermal-branch entries from the un-annealed one.
    The cross-field densities follow the samples they are evaluated ON: density2 (RGB field on thermal samples) from
    the thermal run, density2_thermal (thermal field on RGB samples) from the RGB run."""
    out = dict(ref_rgb)
    for k, v in ref_thermal.items():
        if k.endswith("_thermal") and k != "density2_thermal":
            out[k] = v
    out["density2"] = ref_thermal["density2"]
    return out


class _RefFrustums:
    """shaped like the reference's Frustums TensorDataclass (cameras/rays.py:32-103): no private layout"""

    def __init__(self, origins, directions, starts, ends, pixel_area):
        self.origins, self.directions, self.starts, self.ends, self.pixel_area = origins, directions, starts, ends, pixel_area
        self.offsets = None

    @property
    def shape(self):
        return self.origins.shape[:-1]

    def get_positions(self):
        return self.origins + self.directions * (self.starts + self.ends) / 2


class _RefRaySamples:
    def __init__(self, rs):
        f = rs.frustums
        self.frustums = _RefFrustums(f.origins.contiguous(), f.directions.contiguous(), f.starts.contiguous(),
                                     f.ends.contiguous(), f.pixel_area)
        self.camera_indices = rs.camera_indices
        self.deltas = rs.deltas
        self.spacing_starts, self.spacing_ends = rs.spacing_starts, rs.spacing_ends
        self.spacing_to_euclidean_fn = rs.spacing_to_euclidean_fn
        self.metadata, self.times = None, None

    @property
    def shape(self):
        return self.frustums.shape


def test_reference_shaped_ray_samples_run_through_the_components():
    """A RaySamples/RayBundle built by the reference (cameras/rays.py:251-295) has no `_layout`: the components must
    take the generic path and return what the layout fast path returns (missing 7)."""
    model, batch, _ = _small()
    _no_jitter(model)
    rb = model.collider(tn.RayBundle(origins=batch["origins"], directions=batch["directions"],
                                     pixel_area=batch["pixel_area"], camera_indices=batch["camera_indices"]))
    with torch.no_grad():
        rs = model.proposal_sampler.initial_sampler(rb, num_samples=64)
        ref_rs = _RefRaySamples(rs)
        assert not hasattr(ref_rs, "_layout")
        prop = model.proposal_networks[0]
        d_fast, d_ref = prop.get_density(rs)[0], prop.get_density(ref_rs)[0]
        torch.testing.assert_close(d_ref, d_fast, rtol=1e-4, atol=1e-6)
        w = rs.get_weights(d_fast)
        new_fast = model.proposal_sampler.pdf_sampler(rb, rs, w, num_samples=24)
        new_ref = model.proposal_sampler.pdf_sampler(rb, ref_rs, w, num_samples=24)
        torch.testing.assert_close(new_ref.frustums.starts, new_fast.frustums.starts, rtol=1e-5, atol=1e-6)
        ref_new = _RefRaySamples(new_fast)
        f = model.field
        dens_fast, emb_fast = f.get_density(new_fast)
        dens_ref, emb_ref = f.get_density(ref_new)
        torch.testing.assert_close(dens_ref, dens_fast, rtol=1e-4, atol=1e-6)
        out_fast = f.get_outputs(new_fast, density_embedding=emb_fast)[tn.FieldHeadNames.RGB]
        out_ref = f.get_outputs(ref_new, density_embedding=emb_ref)[tn.FieldHeadNames.RGB]
        torch.testing.assert_close(out_ref, out_fast, rtol=1e-4, atol=1e-5)
        full_fast, full_ref = f(new_fast), f(ref_new)
        for k in full_fast:
            torch.testing.assert_close(full_ref[k], full_fast[k], rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------------------------------- NCCL, world_size 2
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ddp_worker(rank, world, port, mode, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      TN_COMM=mode)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import bench
        import nerfstudio_thermal_b200 as tn_  # noqa: F401
        from nerfstudio_thermal_b200 import engine as eng
        # different seeds per rank on purpose: sync_replicas() must make the replicas identical (DDP's broadcast)
        model, _ = pu.bench_like_model(tn_, "separate", 15, "trained", num_cams=8, seed=100 + rank)
        model = model.to(dev).train()
        for s in (model.proposal_sampler, model.proposal_sampler_thermal):
            s.initial_sampler.train_stratified = False
            s.pdf_sampler.train_stratified = False
        rays = 1024
        batch = {k: v.to(dev) for k, v in bench.make_batch(rays, 42 + rank, num_cams=8).items()}
        runner = eng.GraphedTrainStep(model, batch, use_graph=True)
        assert runner._comm_in_graph == (mode == "overlap") and runner._pipeline == (mode == "pipeline")
        for _ in range(2):
            runner.step(batch)
        torch.cuda.synchronize()
        grads = runner.grads.flat.clone()
        gathered = [torch.empty_like(grads) for _ in range(world)]
        dist.all_gather(gathered, grads)
        assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks disagree on the averaged gradients"
        for f in runner._fields:  # the comparison run below is local: no collective may start from a hook
            f.grads_ready_callback = None
        if rank == 0:
            # single-GPU gradients of the concatenated batch on the same (rank-0) weights
            batches = [bench.make_batch(rays, 42 + r, num_cams=8) for r in range(world)]
            cat = {k: torch.cat([b[k] for b in batches]).to(dev) for k in batches[0]}
            runner.grads.zero_()
            rb = tn_.RayBundle(origins=cat["origins"], directions=cat["directions"], pixel_area=cat["pixel_area"],
                               camera_indices=cat["camera_indices"])
            _, losses, _ = model.get_train_loss_dict(rb, {"image": cat["image"], "is_thermal": cat["is_thermal"]})
            losses.total.backward()
            torch.cuda.synchronize()
            single = runner.grads.flat
            res = {}
            for name, (b, e) in runner.grads.group_ranges.items():
                res[name] = pu.rel_l2(grads[b:e], single[b:e])
            torch.save(res, os.path.join(out_dir, f"ddp_{mode}.pt"))
        # leave without destroying the group: tearing NCCL down while captured graphs still hold its kernels can stall
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)
    except BaseException:
        import traceback
        traceback.print_exc()
        os._exit(1)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("mode", ["pipeline", "overlap", "after"])
def test_nccl_world2_gradients_equal_single_gpu_on_concatenated_batch(tmp_path, mode):
    """SURVEY 8e parity test: N-GPU averaged gradients == 1-GPU gradients of the concatenated batch (every loss term
    is a mean over rays, patches or samples, and the ranks hold equal shares), rel 1e-3, for the overlapped two-segment
    exchange captured in the graph and for the single trailing all-reduce."""
    import torch.multiprocessing as mp
    mp.spawn(_ddp_worker, args=(2, _free_port(), mode, str(tmp_path)), nprocs=2, join=True)
    res = torch.load(os.path.join(str(tmp_path), f"ddp_{mode}.pt"))
    pu.report(f"nccl_world2_{mode}", res)
    assert len(res) >= 6
    for k, v in res.items():
        assert v <= 1e-3, (mode, k, v)


def test_three_phase_schedule_equals_the_single_graph():
    """engine "pipeline" (sample phase | render phase + backward to the cut | backward of the sample phase, three
    graphs) gives the losses and gradients of the one-graph step: the cut only re-orders independent work."""
    from nerfstudio_thermal_b200 import engine as eng
    import bench

    model, _ = pu.bench_like_model(tn, "separate", 15, "trained", num_cams=8, seed=7)
    model = model.to(DEV).train()
    for s in (model.proposal_sampler, model.proposal_sampler_thermal):
        s.initial_sampler.train_stratified = False  # no jitter: both runners see the same samples
        s.pdf_sampler.train_stratified = False
    batch = {k: v.to(DEV) for k, v in bench.make_batch(1024, 3, num_cams=8).items()}
    res = []
    for pipeline in (False, True):
        runner = eng.GraphedTrainStep(model, batch, use_graph=True, pipeline=pipeline)
        assert runner._pipeline == pipeline
        for _ in range(2):
            total = runner.step(batch)
        runner.finish_exchange()
        torch.cuda.synchronize()
        res.append((float(total), {k: float(v) for k, v in runner.losses.items()}, runner.grads.flat.clone(),
                    dict(runner.grads.group_ranges)))
    (t0, l0, g0, ranges), (t1, l1, g1, _) = res
    assert abs(t0 - t1) <= 1e-6 * abs(t0)
    for k in l0:
        assert abs(l0[k] - l1[k]) <= 2e-6 * max(abs(l0[k]), 1e-6), k
    for name, (b, e) in ranges.items():
        assert pu.rel_l2(g1[b:e], g0[b:e]) <= 2e-5, name


def test_packed_host_batch_moves_with_one_copy_and_equals_the_per_tensor_load():
    """`GraphedTrainStep.host_batch()`: pinned views of one buffer laid out like the device-side inputs; a step fed
    with it gives the loss of a step fed with the separate tensors."""
    model, batch, _ = _small()
    _no_jitter(model)
    runner = engine.GraphedTrainStep(model, batch, use_graph=True)
    cpu = {k: v.cpu() for k, v in batch.items()}
    ref = float(runner.step(cpu))
    host = runner.host_batch()
    assert host["_packed"].is_pinned() and set(cpu) <= set(host)
    for k, v in cpu.items():
        assert host[k].shape == v.shape and host[k].dtype == v.dtype
        host[k].copy_(v)
    for k in cpu:  # scribble over the device inputs: the packed copy must restore every one of them
        runner.static[k].zero_()
    got = float(runner.step(host))
    assert abs(got - ref) <= 1e-6 * abs(ref)  # (the ray-level loss sums are atomics: last-bit differences between replays)
    for k, v in cpu.items():
        assert torch.equal(runner.static[k].cpu(), v), k
