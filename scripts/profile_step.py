"""Kernel-time breakdown of one C2 training step with torch.profiler (CUPTI), no replay: python scripts/profile_step.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from reflect_sampling_nerf_b200.model import ReflectSamplingNeRFModel, ReflectSamplingNeRFModelConfig
from reflect_sampling_nerf_b200.rays import RayBundle
from reflect_sampling_nerf_b200.train_path import TrainStep
torch.manual_seed(0)
model = ReflectSamplingNeRFModel(ReflectSamplingNeRFModelConfig(**bench.CFG)).cuda().train()
o, d, a, img = [t.cuda() for t in bench.synthetic_batch(bench.RAYS_PER_GPU, 1000)]
st = TrainStep(model)
for _ in range(4):
    st.step(RayBundle(origins=o, directions=d, pixel_area=a), img)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        st.step(RayBundle(origins=o, directions=d, pixel_area=a), img)
    torch.cuda.synchronize()
from torch.autograd import DeviceType
kern = [e for e in prof.events() if e.device_type == DeviceType.CUDA]
busy = sum(e.device_time_total for e in kern) / 3 / 1e3
t_first, t_last = min(e.time_range.start for e in kern), max(e.time_range.end for e in kern)
print(f"GPU kernel time per step: {busy:.2f} ms; wall per step (first kernel start -> last kernel end)/3: {(t_last - t_first) / 3 / 1e3:.2f} ms; "
      f"kernels per step: {len(kern) // 3}")
# largest idle gaps between consecutive kernels
ks = sorted(kern, key=lambda e: e.time_range.start)
gaps = sorted(((b.time_range.start - a.time_range.end, a.name[:50], b.name[:50]) for a, b in zip(ks, ks[1:])), reverse=True)[:12]
for g, a, b in gaps:
    print(f"  idle {g/1e3:7.3f} ms between {a} -> {b}")
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.device_time_total for e in ev)
print(f"total device time per step: {tot/3/1e3:.2f} ms")
for e in sorted(ev, key=lambda e: -e.device_time_total)[:32]:
    print(f"{e.device_time_total/3/1e3:8.3f} ms {e.count//3:5d}x {100*e.device_time_total/tot:5.1f}%  {e.key[:90]}")
