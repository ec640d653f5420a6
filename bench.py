#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native reflect-sampling-nerf hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|render] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic rays of BASELINE.json configs[1]
(16,384 rays per GPU, 128 coarse + 128 fine samples, 64 + 64 reflected samples, random-init field):
  train : get_outputs + get_loss_dict + backward + gradient all-reduce (N>1) + optimizer step
  render: get_outputs in eval mode (BASELINE.json configs[2] chunk form)
Rays shard across ranks with no data-path collective (weak scaling: 16,384 rays per GPU); the only
collective is the per-step gradient all-reduce of the training workload.

Prints ONE JSON line (rank 0).  `value` = whole-job rays/s with the batch resident in HBM; `e2e` = the same
through the public API with HOST (pinned) ray/pixel buffers, H2D + D2H copies inside the timed region.
`--impl reference` times the oracle restatement of the reference's PyTorch path on the host CPU cores
(the reference itself cannot be imported: nerfstudio is absent).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import torch  # noqa: E402

CFG = dict(num_coarse_samples=128, num_importance_samples=128,
           num_reflect_coarse_samples=64, num_reflect_importance_samples=64)
RAYS_PER_GPU = 16384
PIXEL_AREA = 3.2e-6          # 400x400 Blender camera, f = 555.6
FLOP_PRIMARY, FLOP_REFLECT, FLOP_INF = 1230592, 1229056, 1225472   # forward, per point


def synthetic_batch(n: int, seed: int):
    """directions ~ normalised N(0,I), origins = -4 d + 0.3 N(0,I); target pixels ~ U(0,1)."""
    g = torch.Generator().manual_seed(seed)
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    o = -4.0 * d + 0.3 * torch.randn(n, 3, generator=g)
    area = torch.full((n, 1), PIXEL_AREA)
    image = torch.rand(n, 3, generator=g)
    return o, d, area, image


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if len(s) >= 6 and s[0].isdigit())
        reasons = set()
        for s in self.samples:
            if len(s) >= 6:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = [int(s[1]) for s in self.samples if len(s) >= 6 and s[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ reference arm
def oracle_step_rays_per_s(n_rays: int, workload: str, threads: int, steps: int, warmup: int):
    """The oracle restatement of the reference's PyTorch path on the host CPU (bounded sample of the workload)."""
    from oracle import upstream as U
    from oracle.refpath import OracleModel
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = OracleModel(**CFG)
    model.train(workload == "train")
    opt = torch.optim.RAdam(model.field.parameters(), lr=1e-3, eps=1e-15)
    o, d, area, image = synthetic_batch(n_rays, 0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        bundle = U.RayBundle(origins=o, directions=d, pixel_area=area)
        if workload == "train":
            opt.zero_grad(set_to_none=True)
            out = model(bundle)
            loss = sum(model.get_loss_dict(out, {"image": image}).values())
            loss.backward()
            opt.step()
        else:
            with torch.no_grad():
                model(bundle)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return n_rays / sec, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.ref_rays
    rps, sec = oracle_step_rays_per_s(n, args.workload, threads, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "rays_per_sec", "value": rps, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C2 {args.workload} step (128+128 samples, 64+64 reflected), bounded sample of "
                               f"{n} of {RAYS_PER_GPU} rays per step", "rays_per_step": n, **CFG},
        "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port",
                         "sample": f"{n} rays/step x {args.steps} steps, oracle restatement (PyTorch fp32 CPU)"},
        "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch.distributed as dist
    from reflect_sampling_nerf_b200 import _lib, ops
    from reflect_sampling_nerf_b200.model import ReflectSamplingNeRFModel, ReflectSamplingNeRFModelConfig
    from reflect_sampling_nerf_b200.rays import RayBundle

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the rsn_b200 kernels have no CPU fallback")
    # NCCL prints its version banner on stdout at communicator creation: keep stdout clean for the ONE JSON line by
    # pointing fd 1 at stderr for the duration of the run and writing the result to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.call("rsn_device_ok")

    torch.manual_seed(0)                       # identical random-init field on every rank
    model = ReflectSamplingNeRFModel(ReflectSamplingNeRFModelConfig(**CFG)).to(dev)
    train = args.workload == "train"
    model.train(train)
    n = args.rays
    o, d, area, image = synthetic_batch(n, 1000 + rank)
    host = [t.pin_memory() for t in (o, d, area, image)]
    resident = [t.to(dev) for t in host]
    opt = None
    if train:
        from reflect_sampling_nerf_b200.train_path import TrainStep
        stepper = TrainStep(model, world_size=world)

    counters = {"launches": 0}
    orig_call = _lib.call

    def counting_call(name, *a):
        counters["launches"] += 1
        return orig_call(name, *a)
    _lib.call = counting_call
    ops._lib.call = counting_call

    loss_ring = {"i": 0, "ev": [None, None], "seen": [],
                 "buf": [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]}

    def one_step(bufs, e2e: bool):
        if e2e:
            bufs = [t.to(dev, non_blocking=True) for t in host]
        bo, bd, ba, bi = bufs
        bundle = RayBundle(origins=bo, directions=bd, pixel_area=ba)
        if train:
            loss = stepper.step(bundle, bi)
            if e2e:
                # D2H read of the step's loss, consumed by the host one step late (pinned double buffer + event), the way
                # a training loop logs it: the copy is inside this step's timed region, the host never stalls the queue
                slot = loss_ring["i"] & 1
                if loss_ring["ev"][slot] is not None:
                    loss_ring["ev"][slot].synchronize()
                    loss_ring["seen"].append(float(loss_ring["buf"][slot]))
                loss_ring["buf"][slot].copy_(loss.detach().reshape(()), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                loss_ring["ev"][slot] = ev
                loss_ring["i"] += 1
            return None
        out = model(bundle)
        if e2e:
            return out["mid_rgb_fine"].cpu(), out["mid_reflect_fine"].cpu()
        return out

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2 (126 MB)

    def timed(e2e: bool, steps: int, warmup: int):
        for _ in range(warmup):
            one_step(resident, e2e)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ops.PROFILE = [] if not e2e else None
        counters["launches"] = 0
        evs = []
        for _ in range(steps):
            flush.zero_()                       # L2 flush between timed iterations (outside the event pairs)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            one_step(resident, e2e)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        per_step = [a.elapsed_time(b) for a, b in evs]
        ms = sum(per_step)
        timed.last_per_step = per_step
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        prof, ops.PROFILE = ops.PROFILE, None
        return float(t.item()) / steps, counters["launches"], prof

    sampler = ClockSampler(local)
    sampler.start()
    ms_step, launches, prof = timed(False, args.steps, args.warmup)
    step_ms = [round(x, 2) for x in timed.last_per_step]
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    ms_e2e, _, _ = timed(True, max(2, args.steps // 2), 1)

    # roofline: every field kernel from its own CUDA-event launch durations; `roofline` = the one with the largest
    # share of the step (tensor-bound kernels against the sustained cuBLAS bf16 peak, HBM-bound against the copy peak)
    roof, roof_all = None, []
    if prof:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) \
            if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else None
        tf_peak = peaks["bf16_tflops_sustained"] if peaks else 1400.0
        bw_peak = peaks["hbm_gbs"] if peaks else 6650.0
        src = "MEASURED_PEAKS.json (bf16_tflops_sustained: kernels timed inside a long step; hbm_gbs)" if peaks \
            else "B200_PROFILING.md fallback"
        # ncu --set full dram__bytes_read + dram__bytes_write per launch of a C2 primary pass (2.1 M points;
        # profiles/r01_field_all_v5_ncu.txt)
        ncu_traffic = {"field_fwd_kernel": 0.0988e9, "field_fwd_kernel[train]": 10.74e9,
                       "field_chain_kernel<normals>": 1.10e9, "field_chain_kernel<backward>": 10.26e9,
                       "field_wgrad_kernel": 20.72e9}
        by = {}
        for (name, a, b, flop, nbytes) in prof:
            d = by.setdefault(name, [0.0, 0.0, 0.0, 0])
            d[0] += a.elapsed_time(b); d[1] += flop; d[2] += nbytes; d[3] += 1
        for name, (ms, flop, nbytes, cnt) in by.items():
            tf, gb = flop / ms / 1e9, nbytes / ms / 1e6
            hbm_bound = (gb / bw_peak) > (tf / tf_peak)
            roof_all.append({"kernel": name, "bound": "hbm" if hbm_bound else "tensor",
                             "achieved": gb if hbm_bound else tf, "peak": bw_peak if hbm_bound else tf_peak,
                             "unit": "GB/s" if hbm_bound else "TFLOP/s",
                             "frac": (gb / bw_peak) if hbm_bound else (tf / tf_peak),
                             "tflops": tf, "gbs": gb, "traffic": ncu_traffic.get(name),
                             "launches": cnt, "avg_launch_ms": ms / cnt, "share_of_step": ms / (ms_step * args.steps)})
        roof_all.sort(key=lambda r: -r["share_of_step"])
        roof = dict(roof_all[0], peak_source=src)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rps, sec = oracle_step_rays_per_s(args.ref_rays, args.workload, threads, 1, 1)
        cpu_base = {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port",
                    "sample": f"{args.ref_rays} of {n} rays, 1 warm-up + 1 timed step ({sec:.1f} s), oracle restatement "
                              f"(PyTorch fp32 CPU) of the same {args.workload} step"}

    if rank == 0:
        rays_total = n * world
        h2d = sum(t.numel() * t.element_size() for t in host)
        d2h = 4 if train else 2 * n * 3 * 4
        line = {
            "metric": "rays_per_sec", "value": rays_total / (ms_step * 1e-3), "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"C2 {args.workload} step: {n} rays/GPU x (128 coarse + 128 fine) samples + reflected "
                                   f"(64 + 64) per masked ray, random-init field", "rays_per_gpu": n, **CFG,
                       "l2": "256 MB buffer written between timed iterations; per-pass field outputs (134 MB) exceed L2",
                       "parallelism": f"dp{world} (rays sharded, no data-path collective)"},
            "e2e": {"value": rays_total / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "step_ms_rank0": step_ms, "clocks": sampler.summary(), "roofline": roof, "roofline_all": roof_all, "cpu_baseline": cpu_base,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "render"])
    ap.add_argument("--rays", type=int, default=RAYS_PER_GPU)
    ap.add_argument("--ref-rays", type=int, default=1024, help="bounded CPU sample (rays per oracle step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
