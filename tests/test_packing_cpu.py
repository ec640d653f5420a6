"""CPU-only host logic: operand packing (block images), layout constants vs the C-ABI library, gradient-blob
unpacking, IPE frequency table, and the data-parallel flat-gradient flush over gloo (world_size 2)."""
import ctypes
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reflect_sampling_nerf_b200 import _lib, ops, packing
from reflect_sampling_nerf_b200.blocks import pack_blocks, unpack_blocks
from reflect_sampling_nerf_b200.field import ReflectSamplingNeRFNerfField
from reflect_sampling_nerf_b200.plugin_field_init import random_field_state


def test_block_image_roundtrip_and_swizzle():
    m = torch.randn(128, 256).bfloat16()
    img = pack_blocks(m)
    assert img.shape == (4, 128, 128) and img.dtype == torch.uint8
    assert torch.equal(unpack_blocks(img, 128), m)
    # element (r, k) of block kb lives at r*128 + (((k>>3) ^ (r&7)) << 4) + (k&7)*2   (csrc/umma.cuh block_off)
    r, k = 13, 37
    off = r * 128 + ((((k >> 3) ^ (r & 7)) << 4) | ((k & 7) << 1))
    raw = img[0].reshape(-1)[off:off + 2].view(torch.bfloat16)
    assert raw.item() == m[r, k].item()


def test_chunk_major_image_roundtrip_and_offsets():
    """csrc/field_layout.cuh stash_chunk_off: chunk c (8 features) of point r at (r // 64) * 8192 + c * 1024 + (r % 64) * 16."""
    from reflect_sampling_nerf_b200.blocks import pack_blocks_cm, unpack_blocks_cm
    m = torch.randn(256, 128).bfloat16()
    img = pack_blocks_cm(m)
    assert img.shape == (2, 2, 16384) and img.dtype == torch.uint8
    assert torch.equal(unpack_blocks_cm(img), m)
    for t, kb, r, c in ((0, 0, 0, 0), (1, 1, 77, 5), (0, 1, 127, 7), (1, 0, 64, 3)):
        off = (r // 64) * 8192 + c * 1024 + (r % 64) * 16
        assert torch.equal(img[t, kb, off:off + 16].view(torch.bfloat16), m[128 * t + r, kb * 64 + 8 * c: kb * 64 + 8 * c + 8])


def test_layout_constants_match_library():
    lib = _lib.lib()
    assert lib.rsn_field_blob_bytes() == packing.FWD_BLOB_BYTES
    assert lib.rsn_field_blob_t_bytes() == packing.BWD_BLOB_BYTES
    assert lib.rsn_field_bias_count() == packing.N_BIAS
    assert lib.rsn_field_stash_bytes(129) == 2 * (41 * 16384 + 9 * 4 * 128 * 8)
    assert lib.rsn_field_dy_stash_bytes(128) == 39 * 16384
    offs, shapes, total = ops.wgrad_layout()
    assert len(shapes) == 14 and total == sum(m * n for m, n in shapes) + sum(m for (m, n), o in zip(shapes, offs[1::2]) if o >= 0)


def test_ipe_frequency_table_is_torch_s():
    buf = (ctypes.c_float * 16)()
    assert _lib.lib().rsn_ipe_freqs(buf) == 0
    ref = 2 ** torch.linspace(0.0, 16.0, 16)
    assert torch.equal(torch.tensor(list(buf)), ref)


def test_pack_field_places_weights_where_the_kernels_read_them():
    sd = random_field_state()
    blob, bias = packing.pack_field(sd)
    assert blob.numel() == packing.FWD_BLOB_BYTES and bias.numel() == packing.N_BIAS
    w0 = unpack_blocks(blob[:2 * 32768].view(2, 256, 128), 256)            # layer 0: [256, 128] (99 used)
    assert torch.equal(w0[:, :99], sd["mlp_base.layers.0.weight"].bfloat16())
    assert float(w0[:, 99:].abs().max()) == 0.0
    off = (2 + 12) * 32768                                                   # layer 4: enc part then hidden part
    w4 = unpack_blocks(blob[off:off + 6 * 32768].view(6, 256, 128), 256)
    assert torch.equal(w4[:, :99], sd["mlp_base.layers.4.weight"][:, :99].bfloat16())
    assert torch.equal(w4[:, 128:], sd["mlp_base.layers.4.weight"][:, 99:].bfloat16())
    assert torch.equal(bias[256:512], sd["mlp_base.layers.1.bias"])
    assert torch.equal(bias[packing.BIAS_HEAD + 5:packing.BIAS_HEAD + 8], sd["field_output_diff.net.bias"])
    blob_t, wd = packing.pack_field_t(sd)
    assert blob_t.numel() == packing.BWD_BLOB_BYTES
    assert torch.equal(wd, sd["field_output_density.net.weight"].reshape(256).bfloat16())
    w7t = unpack_blocks(blob_t[245760 + 6 * 131072: 245760 + 7 * 131072].view(4, 256, 128), 256)   # BT_L(7)
    assert torch.equal(w7t, sd["mlp_base.layers.7.weight"].T.bfloat16())


def test_unpack_grads_maps_every_region_to_its_parameter():
    offs, shapes, total = ops.wgrad_layout()
    blob = torch.arange(total, dtype=torch.float32)
    g = packing.unpack_grads(blob, offs, shapes)
    sd = random_field_state()
    for name, p in sd.items():
        if "field_output_low" in name:
            assert name not in g
            continue
        assert g[name].shape == p.shape, name
    # spot checks against the job table of csrc/field_wgrad.cu
    assert g["mlp_base.layers.0.weight"][3, 5] == offs[0] + 3 * 128 + 5
    assert g["mlp_base.layers.4.weight"][7, 99 + 11] == offs[2 * 5] + 7 * 256 + 11
    assert g["mlp_base.layers.4.weight"][7, 11] == offs[2 * 4] + 7 * 128 + 11
    assert g["field_output_tint.net.weight"][2, 9] == offs[2 * 10] + (16 + 10) * 256 + 9
    assert g["field_output_mid.net.bias"][1] == offs[2 * 10 + 1] + 1
    assert g["mlp_mid.layers.0.weight"][4, 2] == offs[2 * 13] + 4 * 64 + 2
    assert g["mlp_mid.layers.0.weight"][4, 34 + 2] == offs[2 * 12] + 4 * 256 + 2


def test_field_state_dict_and_no_cpu_fallback():
    f = ReflectSamplingNeRFNerfField()
    assert sum(p.numel() for p in f.parameters()) == 618513
    assert "mlp_base.layers.4.weight" in f.state_dict() and f.state_dict()["mlp_base.layers.4.weight"].shape == (256, 355)
    with pytest.raises(ValueError):
        ReflectSamplingNeRFNerfField(base_mlp_num_layers=4)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            f.evaluate_samples(torch.zeros(2, 3), torch.zeros(2, 3), torch.ones(2, 1), torch.zeros(2, 9))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _flush_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from reflect_sampling_nerf_b200.train_path import _flush_grads
    torch.manual_seed(0)
    field = ReflectSamplingNeRFNerfField()
    field.dp_world_size = world
    _, _, total = ops.wgrad_layout()
    field._grad_blob = torch.full((total,), float(rank + 1))        # rank r contributed (r + 1) everywhere
    _flush_grads(field)
    g = field.mlp_base.layers[3].weight.grad
    ok = bool(torch.allclose(g, torch.full_like(g, (1 + world) / 2.0)))   # mean over ranks
    ok &= field.field_output_low.net.weight.grad is None
    ok &= field._grad_blob is None
    torch.save(ok, os.path.join(out_dir, f"ok{rank}.pt"))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world_size_2(tmp_path):
    """The DDP replacement (pipeline.py:73-77): one all-reduce of the flat gradient blob, averaged, then unpacked."""
    mp.spawn(_flush_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert all(torch.load(os.path.join(tmp_path, f"ok{r}.pt")) for r in range(2))


def _split_flush_worker(rank, world, port, out_dir):
    """The two halves TrainStep's split CUDA graphs are built from: reduce the blob, then flush WITHOUT a collective."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from reflect_sampling_nerf_b200.train_path import _allreduce_blob, _flush_grads
    torch.manual_seed(0)
    _, _, total = ops.wgrad_layout()
    base = torch.randn(total)
    grads = []
    for split in (False, True):
        torch.manual_seed(1)                                  # same parameters: the bottleneck layer's gradients depend on them
        field = ReflectSamplingNeRFNerfField()
        field.dp_world_size = world
        field._grad_blob = base * float(rank + 1)
        if split:
            _allreduce_blob(field, field._grad_blob)          # (gloo: SUM + scale; NCCL averages inside the collective)
            _flush_grads(field, allreduce=False)
        else:
            _flush_grads(field)
        grads.append(torch.cat([p.grad.reshape(-1) for p in field.parameters() if p.grad is not None]))
    # a flush with allreduce=False on an un-reduced blob must NOT communicate: rank-local values stay rank-local
    field = ReflectSamplingNeRFNerfField()
    field.dp_world_size = world
    field._grad_blob = torch.full((total,), float(rank + 1))
    _flush_grads(field, allreduce=False)
    g = field.mlp_base.layers[3].weight.grad
    torch.save({"same": bool(torch.equal(grads[0], grads[1])), "local": bool(torch.allclose(g, torch.full_like(g, float(rank + 1))))},
               os.path.join(out_dir, f"s{rank}.pt"))
    dist.destroy_process_group()


def test_split_flush_equals_the_one_call_flush_world_size_2(tmp_path):
    mp.spawn(_split_flush_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    res = [torch.load(os.path.join(tmp_path, f"s{r}.pt")) for r in range(2)]
    assert all(r["same"] and r["local"] for r in res), res


def _bcast_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from reflect_sampling_nerf_b200.train_path import _flush_grads, sync_parameters
    torch.manual_seed(100 + rank)                     # nerfstudio seeds every rank differently (machine.seed + global_rank)
    field = ReflectSamplingNeRFNerfField()
    before = field.mlp_base.layers[2].weight.detach().clone()
    sync_parameters(field)
    # one "step": identical averaged gradients applied to (now) identical parameters
    field.dp_world_size = world
    _, _, total = ops.wgrad_layout()
    field._grad_blob = torch.full((total,), float(rank + 1))
    _flush_grads(field)
    with torch.no_grad():
        for p in field.parameters():
            if p.grad is not None:
                p.add_(p.grad, alpha=-1e-3)
    flat = torch.cat([p.detach().reshape(-1) for p in field.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    torch.save({"same": all(torch.equal(gathered[0], t) for t in gathered),
                "changed": bool(rank == 0 or not torch.equal(before, field.mlp_base.layers[2].weight.detach() + 1e-3 * field.mlp_base.layers[2].weight.grad))},
               os.path.join(out_dir, f"b{rank}.pt"))
    dist.destroy_process_group()


def test_parameters_are_broadcast_from_rank_0_world_size_2(tmp_path):
    """without DDP's construction-time broadcast, differently seeded ranks would train different
    replicas.  After sync_parameters + one averaged-gradient step every rank holds bit-identical parameters."""
    mp.spawn(_bcast_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    res = [torch.load(os.path.join(tmp_path, f"b{r}.pt")) for r in range(2)]
    assert all(r["same"] for r in res)
    assert res[1]["changed"]                          # rank 1's own initialisation was replaced


def test_loss_coefficient_warmup_matches_reference_pipeline():
    """reflect_sampling_nerf_pipeline.py:79-91 (also restated in oracle.refpath.warmup_coefficients)."""
    from oracle.refpath import LOSS_COEFFICIENTS, warmup_coefficients
    from reflect_sampling_nerf_b200.model import LOSS_COEFFICIENTS as MINE
    from reflect_sampling_nerf_b200.pipeline import warmup_loss_coefficients
    assert MINE == LOSS_COEFFICIENTS
    for step in (0, 49, 50, 1000):
        assert warmup_loss_coefficients(step, dict(MINE)) == warmup_coefficients(step, dict(LOSS_COEFFICIENTS))


def test_model_contract_without_gpu():
    from reflect_sampling_nerf_b200.model import ReflectSamplingNeRFModel, ReflectSamplingNeRFModelConfig
    cfg = ReflectSamplingNeRFModelConfig()
    assert (cfg.num_coarse_samples, cfg.num_importance_samples, cfg.num_reflect_coarse_samples,
            cfg.num_reflect_importance_samples) == (128, 128, 64, 64)
    m = cfg.setup()
    assert isinstance(m, ReflectSamplingNeRFModel)
    assert list(m.get_param_groups()) == ["fields"] and len(m.get_param_groups()["fields"]) == 34
    assert (m.near, m.far) == (1.0 / 16, 256)
    for name in ("sampler_uniform", "sampler_pdf", "sampler_reciprocal", "sampler_reflect_pdf", "rgb_loss", "field",
                 "renderer_rgb", "renderer_accumulation", "renderer_depth", "renderer_normals", "renderer_roughness",
                 "renderer_factor", "renderer_reflect", "psnr", "ssim"):
        assert hasattr(m, name), name
    # component classes under the reference's names and constructor contracts (components.py:14-36, 38-140; model.py:109-124)
    from reflect_sampling_nerf_b200 import components as C
    assert isinstance(m.sampler_reciprocal, C.ReciprocalSampler) and m.sampler_reciprocal.tan == 0.25
    assert isinstance(m.sampler_uniform, C.UniformSampler) and isinstance(m.sampler_pdf, C.PDFSampler)
    assert m.sampler_pdf.include_original is False and m.sampler_pdf.histogram_padding == 0.01
    assert isinstance(m.field.direction_encoding, C.IntegratedSHEncoding) and m.field.direction_encoding.get_out_dim() == 34
    assert isinstance(m.field.position_encoding, C.NeRFEncoding) and m.field.position_encoding.get_out_dim() == 99
    assert m.renderer_rgb.background_color.tolist() == [1.0, 1.0, 1.0] and m.renderer_reflect.background_color == "random"
    for meth in ("get_blob", "contract", "get_density", "get_pred_normals", "get_normals", "get_roughness", "get_low",
                 "get_mid", "get_diff", "get_tint", "get_inf_color", "get_reflection"):
        assert callable(getattr(m.field, meth)), meth                       # field.py:90-207
    m.field = None
    with pytest.raises(ValueError):
        m.get_param_groups()


def test_wgrad_finish_host_mirror_matches_the_algebra():
    """ops.wgrad_finish on a CPU blob (the path the gloo flush test takes): dW_bott = Wmb^T G, db_bott = Wmb^T db_mid,
    dW_mid[:, bott] = G Wb^T + db_mid bb^T with G = job 12's region; everything else untouched."""
    g = torch.Generator().manual_seed(5)
    offs, shapes, total = ops.wgrad_layout()
    assert shapes[9] == (256, 256) and shapes[10] == (256, 256) and shapes[12] == (128, 256)
    blob = torch.randn(total, generator=g)
    w_b, b_b, w_m = torch.randn(256, 256, generator=g), torch.randn(256, generator=g), torch.randn(128, 290, generator=g)
    G = blob[offs[20] + 64 * 256: offs[20] + 192 * 256].view(128, 256).clone()       # rows 64-191 of job 10's region
    dbm = blob[offs[21] + 64: offs[21] + 192].clone()
    out = blob.clone()
    ops.wgrad_finish(out, w_b, b_b, w_m)
    torch.testing.assert_close(out[offs[18]: offs[18] + 65536].view(256, 256), w_m[:, 34:].T @ G)
    torch.testing.assert_close(out[offs[19]: offs[19] + 256], w_m[:, 34:].T @ dbm)
    torch.testing.assert_close(out[offs[24]: offs[24] + 32768].view(128, 256), G @ w_b.T + torch.outer(dbm, b_b))
    keep = torch.ones(total, dtype=torch.bool)
    keep[offs[18]: offs[19] + 256] = False
    keep[offs[24]: offs[25] + 128] = False
    assert torch.equal(out[offs[25]: offs[25] + 128], dbm)
    assert torch.equal(out[keep], blob[keep])
