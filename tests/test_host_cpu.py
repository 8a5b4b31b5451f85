"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the host-side mirror keeps the
reference's API / state_dict contract, the product path refuses to run without CUDA (no CPU fallback), and
the multi-rank plumbing works over gloo with world_size 2.  No kernel is launched here.
"""
import ctypes
import os
import re
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import nerfstudio_thermal_b200 as tn
from nerfstudio_thermal_b200 import _lib, parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tn_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared  # the ctypes table binds exactly the header
    loaded = _lib.load()
    assert loaded.tn_version() >= 100
    assert loaded.tn_build_arch() == b"sm_100a"


def test_no_cpu_fallback():
    enc = tn.HashEncoding(num_levels=2, log2_hashmap_size=4)
    with pytest.raises(tn.TnKernelError):
        enc(torch.rand(4, 3))
    mlp = tn.MLP(in_dim=32, num_layers=2, layer_width=64, out_dim=16)
    with pytest.raises(tn.TnKernelError):
        mlp(torch.rand(4, 32))
    rs = tn.RaySamples(frustums=tn.Frustums(torch.zeros(2, 3, 3), torch.ones(2, 3, 3), torch.zeros(2, 3, 1),
                                            torch.ones(2, 3, 1), torch.ones(2, 3, 1)), deltas=torch.ones(2, 3, 1))
    with pytest.raises(tn.TnKernelError):
        rs.get_weights(torch.ones(2, 3, 1))


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "nerfstudio_thermal_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_error_conventions_match_reference():
    with pytest.raises(ValueError):  # encodings.py:765-766
        tn.SHEncoding(levels=5)
    with pytest.raises(ValueError):  # ray_samplers.py:551-552
        tn.ProposalNetworkSampler(num_proposal_network_iterations=0)
    with pytest.raises(ValueError):  # ray_samplers.py:297-299
        tn.PDFSampler()(None, None, torch.ones(1, 1, 1))
    with pytest.raises(AssertionError):  # encodings.py:370-373
        tn.HashEncoding(num_levels=2, log2_hashmap_size=4, interpolation="Smoothstep")


def test_hash_encoding_attributes_match_reference_probe():
    enc = tn.HashEncoding(num_levels=16, min_res=16, max_res=2048, log2_hashmap_size=19)
    assert enc.scalings.tolist() == [16, 22, 30, 42, 58, 80, 111, 153, 212, 294, 406, 561, 776, 1072, 1482, 2047]
    assert enc.hash_table.shape == (16 * 2**19, 2) and enc.get_out_dim() == 32
    assert enc.hash_offset.tolist() == [i * 2**19 for i in range(16)]
    assert enc.hash_table.abs().max() <= 1e-3  # U(-1,1) * hash_init_scale
    pts = torch.tensor([[[1, 2, 3]], [[5, 0, 9]]], dtype=torch.int32).expand(2, 16, 3)
    h = enc.hash_fn(pts)
    assert h.shape == (2, 16) and h.dtype == torch.int64
    assert int(h[0, 0]) == ((1 * 1) ^ (2 * 2654435761) ^ (3 * 805459861)) % 2**19


def test_state_dict_contract_and_param_groups(golden):
    g = golden("model_separate.npz")
    cfg = tn.ThermalNerfactoModelConfig(
        density_mode="separate", log2_hashmap_size=9,
        proposal_net_args_list=[{"hidden_dim": 16, "log2_hashmap_size": 8, "num_levels": 5, "max_res": m,
                                 "use_linear": False} for m in (128, 256)])
    model = cfg.setup(num_train_data=8, metadata={"is_thermal": [0] * 4 + [1] * 4})
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd/")}
    sd["device_indicator_param"] = torch.empty(0)
    assert sorted(model.state_dict()) == sorted(sd)  # identical key set, incl. the aliased proposal tables
    model.load_state_dict(sd, strict=True)
    pn = model.proposal_networks[0]
    assert pn.encoding.hash_table is pn.mlp_base[0].hash_table
    groups = model.get_param_groups()
    assert sorted(groups) == ["camera_opt", "camera_opt_thermal", "fields", "fields_thermal", "proposal_networks",
                              "proposal_networks_thermal"]
    # default sizes: parameter counts probed from the reference
    big = tn.ThermalNerfactoModelConfig(density_mode="separate").setup(
        num_train_data=8, metadata={"is_thermal": [0] * 4 + [1] * 4})
    counts = {k: sum(p.numel() for p in v) for k, v in big.get_param_groups().items()}
    assert counts == {"proposal_networks": 2621826, "fields": 16789075, "camera_opt": 48,
                      "proposal_networks_thermal": 2621826, "fields_thermal": 16788945, "camera_opt_thermal": 48}
    for mode, ch in (("shared", 4), ("rgb_only", 3)):
        m = tn.ThermalNerfactoModelConfig(density_mode=mode, log2_hashmap_size=6).setup(
            num_train_data=4, metadata={"is_thermal": [0, 0, 1, 1]})
        assert m.field.mlp_head.layers[-1].weight.shape == (ch, 64)
        assert not hasattr(m, "field_thermal")
        assert sorted(m.get_param_groups()) == ["camera_opt", "fields", "proposal_networks"]


def test_camera_optimizer_matches_oracle(golden):
    from oracle import model as om
    g = golden("model_separate.npz")
    pose = g["sd/camera_optimizer.pose_adjustment"]
    opt = tn.CameraOptimizer(tn.CameraOptimizerConfig(mode="SO3xR3"), 8,
                             non_trainable_camera_indices=torch.tensor([4, 5, 6, 7]))
    with torch.no_grad():
        opt.pose_adjustment.copy_(pose)
    rb = tn.RayBundle(origins=g["origins"].clone(), directions=g["directions"].clone(),
                      pixel_area=torch.ones(32, 1), camera_indices=g["camera_indices"])
    opt.apply_to_raybundle(rb)
    frozen = torch.tensor([False] * 4 + [True] * 4)
    o, d = om.apply_camera_optimizer(pose, frozen, g["camera_indices"], g["origins"], g["directions"])
    torch.testing.assert_close(rb.origins, o, atol=1e-7, rtol=1e-6)
    torch.testing.assert_close(rb.directions, d, atol=1e-7, rtol=1e-6)
    off = tn.CameraOptimizer(tn.CameraOptimizerConfig(mode="shared_SO3xR3", penalty_scale=-1), 8)
    assert off.config.mode == "off" and len(list(off.parameters())) == 0


def test_ray_bundle_api():
    rb = tn.RayBundle(origins=torch.rand(4, 6, 3), directions=torch.rand(4, 6, 3), pixel_area=torch.ones(4, 6, 1),
                      camera_indices=torch.zeros(4, 6, 1, dtype=torch.long))
    assert len(rb) == 24 and rb.shape == (4, 6)
    sl = rb.get_row_major_sliced_ray_bundle(5, 13)
    assert sl.origins.shape == (8, 3) and torch.equal(sl.origins, rb.origins.reshape(-1, 3)[5:13])
    flat = rb.flatten()
    rs = flat.get_ray_samples(bin_starts=torch.zeros(24, 5, 1), bin_ends=torch.ones(24, 5, 1))
    assert rs.frustums.origins.shape == (24, 5, 3) and rs.deltas.shape == (24, 5, 1)
    assert rs.camera_indices.shape == (24, 5, 1) and rs.shape == (24, 5)
    fr = tn.Frustums(origins=torch.ones((5, 3)), directions=torch.tensor([[0.0, 1.0, 0.5]]).expand(5, 3),
                     starts=torch.ones((5, 1)) * 2, ends=torch.ones((5, 1)) * 3, pixel_area=torch.ones((5, 1)))
    # the reference's own golden vector, tests/cameras/test_rays.py:11-31
    torch.testing.assert_close(fr.get_positions(), torch.tensor([[1.0, 3.5, 2.25]]).expand(5, 3), atol=1e-6, rtol=0)


def test_patch_losses_equal_oracle_formulation():
    from oracle import model as om
    from nerfstudio_thermal_b200 import losses
    torch.manual_seed(0)
    pred, gt = torch.rand(32, 1), torch.rand(32, 3)
    is_thermal = torch.tensor([0.0] * 16 + [1.0] * 16)
    torch.testing.assert_close(losses.tv_pixel_loss(pred, is_thermal), om.tv_pixel_loss(pred, is_thermal))
    torch.testing.assert_close(losses.cross_channel_loss(pred, gt, is_thermal),
                               om.cross_channel_loss(pred, gt, is_thermal))


def test_shard_chunks_cover_frame_exactly():
    for num_rays, chunk in ((640 * 512, 1 << 15), (1920 * 1080, 1 << 15), (100, 7), (5, 8)):
        for world in (1, 2, 4, 8):
            got = [r for rank in range(world) for r in parallel.shard_chunks(num_rays, chunk, rank, world)]
            assert got == [(s, min(s + chunk, num_rays)) for s in range(0, num_rays, chunk)]
            sizes = [len(parallel.shard_chunks(num_rays, chunk, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ddp_worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)  # identical replicas
        cfg = tn.ThermalNerfactoModelConfig(
            density_mode="separate", log2_hashmap_size=6,
            proposal_net_args_list=[{"hidden_dim": 16, "log2_hashmap_size": 5, "num_levels": 5, "max_res": 128,
                                     "use_linear": False}] * 2)
        model = cfg.setup(num_train_data=4, metadata={"is_thermal": [0, 0, 1, 1]})
        groups = model.get_param_groups()
        buf = parallel.FlatGradBuffer.from_param_groups(groups)
        assert buf.check_views()
        n_params = sum(p.numel() for g in groups.values() for p in g)
        assert n_params <= buf.numel() < n_params + 4 * len(buf.params)
        # rank-dependent "gradients" written through the parameter views (as the scatter kernels would)
        gen = torch.Generator().manual_seed(100 + rank)
        local = []
        for p in buf.params:
            v = torch.randn(p.shape, generator=gen)
            p.grad.copy_(v)
            local.append(v)
        # a parameter left untouched on this rank (unused proposal network) just contributes zeros
        if rank == 1:
            buf.params[0].grad.zero_()
            local[0] = torch.zeros_like(local[0])
        buf.all_reduce_mean()
        gathered = [None] * world
        dist.all_gather_object(gathered, [t.numpy() for t in local])
        for i, p in enumerate(buf.params):
            want = sum(torch.from_numpy(g[i]) for g in gathered) / world
            torch.testing.assert_close(p.grad, want, atol=1e-6, rtol=1e-6)
        # the two-segment exchange of the train step (fields first, the rest later): group order puts the early
        # groups at the front of the buffer, each segment is averaged by its own collective
        early = [n for n in ("fields", "fields_thermal") if n in groups]
        order = early + [n for n in groups if n not in early]
        buf2 = parallel.FlatGradBuffer.from_param_groups(groups, order=order)
        cut = max(buf2.group_ranges[n][1] for n in early)
        assert buf2.group_ranges[early[0]][0] == 0 and 0 < cut < buf2.numel() and cut % 4 == 0
        buf2.flat.copy_(torch.arange(buf2.numel(), dtype=torch.float32) * (rank + 1))
        buf2.all_reduce_mean(begin=0, end=cut)
        want = torch.arange(buf2.numel(), dtype=torch.float32)
        torch.testing.assert_close(buf2.flat[:cut], want[:cut] * 1.5)
        torch.testing.assert_close(buf2.flat[cut:], want[cut:] * (rank + 1))  # untouched so far
        buf2.all_reduce_mean(begin=cut)
        torch.testing.assert_close(buf2.flat, want * 1.5)
        buf2.zero_()
        assert all(float(p.grad.abs().sum()) == 0 for p in buf2.params) and buf2.check_views()
        assert parallel.rank_seed(42, rank) == 42 + rank
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_ddp_worker, args=(world, port, results), nprocs=world, join=True)
        assert dict(results) == {0: "ok", 1: "ok"}


def test_reference_checkpoint_round_trip(golden):
    """a dict in the format of Trainer.save_checkpoint (with the pipeline's "_model." prefix, and DDP's
    "module.") loads strict=True; what save_checkpoint writes loads back and carries the reference's key set."""
    import nerfstudio_thermal_b200 as tn
    from nerfstudio_thermal_b200 import checkpoint
    g = golden("model_separate.npz")
    props = [{"hidden_dim": 16, "log2_hashmap_size": 8, "num_levels": 5, "max_res": 128, "use_linear": False},
             {"hidden_dim": 16, "log2_hashmap_size": 8, "num_levels": 5, "max_res": 256, "use_linear": False}]
    cfg = tn.ThermalNerfactoModelConfig(density_mode="separate", log2_hashmap_size=9, proposal_net_args_list=props)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd/")}
    sd["device_indicator_param"] = torch.empty(0)
    for prefix in ("_model.", "_model.module.", "module._model."):
        model = cfg.setup(num_train_data=8, metadata={"is_thermal": [0] * 4 + [1] * 4})
        ckpt = {"step": 1234, "pipeline": {prefix + k: v for k, v in sd.items()}, "optimizers": {}, "scalers": {}}
        ckpt["pipeline"]["datamanager.train_camera_optimizer.pose_adjustment"] = torch.zeros(8, 6)  # non-model key
        # a real step-*.ckpt carries the LPIPS network of NerfactoModel (models/nerfacto.py:253): ignored, not an error
        ckpt["pipeline"][prefix + "lpips.net.lin0.model.1.weight"] = torch.zeros(1, 64, 1, 1)
        ckpt["pipeline"][prefix + "lpips.net.scaling_layer.shift"] = torch.zeros(1, 3, 1, 1)
        assert checkpoint.load_checkpoint(ckpt, model) == 1235
        assert model.step == 1234
        for k, v in model.state_dict().items():
            assert torch.equal(v, sd[k]), k
    out = checkpoint.save_checkpoint(1234, model)
    assert set(out) == {"step", "pipeline", "optimizers", "schedulers", "scalers"}
    assert set(out["pipeline"]) == {"_model." + k for k in sd}
    model2 = cfg.setup(num_train_data=8, metadata={"is_thermal": [0] * 4 + [1] * 4})
    checkpoint.load_checkpoint(out, model2)
    for k, v in model2.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_ctypes_table_matches_header_prototypes():
    """Every prototype of include/tn_b200.h and its ctypes binding agree on the number and the KIND of the arguments
    (pointer / int64 / int / float / double), so a changed C signature cannot silently shift the Python call."""
    import ctypes as C
    text = open(os.path.join(ROOT, "include", "tn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = dict(re.findall(r"\b(?:int|const char\*)\s+(tn_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S))
    assert set(protos) == set(_lib._SIGNATURES)

    def kind_c(arg):
        arg = " ".join(arg.split())
        if "*" in arg:
            return "ptr"
        for t, k in (("int64_t", "i64"), ("uint32_t", "u32"), ("double", "f64"), ("float", "f32"), ("int", "i32")):
            if re.search(rf"\b{t}\b", arg):
                return k
        raise AssertionError(f"unparsed argument {arg!r}")

    def kind_py(t):
        if t in (C.c_int64,):
            return "i64"
        if t in (C.c_int, C.c_int32):
            return "i32"
        if t in (C.c_uint, C.c_uint32):
            return "u32"
        if t is C.c_float:
            return "f32"
        if t is C.c_double:
            return "f64"
        return "ptr"  # c_void_p and POINTER(...) tables

    for name, args in protos.items():
        cargs = [a for a in (s.strip() for s in args.split(",")) if a and a != "void"]
        want = [kind_c(a) for a in cargs]
        got = [kind_py(t) for t in _lib._SIGNATURES[name]]
        assert got == want, (name, got, want)


def test_draw_jitters_reproduces_the_samplers_random_stream():
    """ProposalNetworkSampler.draw_jitters makes exactly the torch.rand calls the sampler would make level by level
    (one [R,1] draw per level with use_single_jitter, ray_samplers.py:96-106, 311-320), and none in eval."""
    cfg = tn.ThermalNerfactoModelConfig(density_mode="separate", log2_hashmap_size=4,
                                        proposal_net_args_list=[{"hidden_dim": 16, "log2_hashmap_size": 4,
                                                                 "num_levels": 2, "max_res": 32, "use_linear": False}] * 2)
    model = cfg.setup(num_train_data=4, metadata={"is_thermal": [0, 0, 1, 1]})
    model.train()
    torch.manual_seed(9)
    got = model.proposal_sampler.draw_jitters(10, "cpu") + model.proposal_sampler_thermal.draw_jitters(10, "cpu")
    torch.manual_seed(9)
    want = [torch.rand((10, 1)) for _ in range(6)]
    assert len(got) == 6 and all(torch.equal(a, b) for a, b in zip(got, want))
    model.eval()
    assert model.proposal_sampler.draw_jitters(10, "cpu") is None
    model.train()
    model.proposal_sampler.initial_sampler.single_jitter = False
    j = model.proposal_sampler.draw_jitters(10, "cpu")
    assert j[0].shape == (10, 257) and j[1].shape == (10, 1)


def test_gradient_sinks_cover_tables_mlps_embeddings_and_poses():
    """FlatGradBuffer.attach_sinks hands every fused backward its accumulation target: hash tables, MLP layers, the
    appearance embeddings and the camera optimizers' pose tables all alias their slice of the ONE flat buffer."""
    cfg = tn.ThermalNerfactoModelConfig(density_mode="separate", log2_hashmap_size=8, proposal_net_args_list=[
        {"hidden_dim": 16, "log2_hashmap_size": 8, "num_levels": 5, "max_res": m, "use_linear": False} for m in (64, 128)])
    model = cfg.setup(num_train_data=6, metadata={"is_thermal": [0, 0, 0, 1, 1, 1]})
    grads = parallel.FlatGradBuffer.from_param_groups(model.get_param_groups())
    n = grads.attach_sinks(model)
    base = grads.flat.untyped_storage().data_ptr()
    sinks = []
    for m in model.modules():
        if isinstance(m, tn.field_components.HashEncoding):
            sinks.append(m.grad_sink)
        elif isinstance(m, tn.field_components.MLP):
            sinks += [t for pair in m.grad_sinks for t in pair]
        elif isinstance(m, tn.field_components.Embedding):
            sinks.append(m.grad_sink)
        elif isinstance(m, tn.CameraOptimizer) and m.config.mode != "off":
            sinks.append(m.grad_sink)
    assert n >= 12 and len(sinks) >= 12 and all(s is not None for s in sinks)
    assert all(s.untyped_storage().data_ptr() == base for s in sinks)
    emb = model.field.embedding_appearance
    assert emb.grad_sink.data_ptr() == emb.embedding.weight.grad.data_ptr()
    cam = model.camera_optimizer
    assert cam.grad_sink.shape == cam.pose_adjustment.shape
    assert cam.grad_sink.data_ptr() == cam.pose_adjustment.grad.data_ptr()
    # "off" optimizers (shared pose tables are disabled by default) have no parameter and no sink
    assert model.shared_camera_optimizer.config.mode == "off" and model.shared_camera_optimizer.grad_sink is None


def test_launch_scratch_and_chain_defaults():
    from nerfstudio_thermal_b200 import fused_ops
    from nerfstudio_thermal_b200.rays import RayLayout
    keep = fused_ops.LaunchScratch()
    made = []
    a = keep.get("k", lambda: made.append(1) or torch.zeros(3))
    b = keep.get("k", lambda: made.append(1) or torch.zeros(3))
    assert a is b and len(made) == 1  # made once, then the same buffer
    with torch.inference_mode():
        assert keep.get("other", lambda: torch.zeros(1)) is None  # never created where it could not be reused
    lay = RayLayout(torch.zeros(2, 3), torch.zeros(2, 3), torch.zeros(2, 5), torch.zeros(2, 5), torch.zeros(2),
                    torch.ones(2))
    assert lay.chain is False  # gradient chaining is opt-in (the model sets it for the layouts of one forward)
