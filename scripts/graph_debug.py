"""Debug: which stream do the backward kernels of the training step run on, with and without side-stream warm-up."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reflect_sampling_nerf_b200 import _lib
from reflect_sampling_nerf_b200.model import ReflectSamplingNeRFModel, ReflectSamplingNeRFModelConfig
from reflect_sampling_nerf_b200.rays import RayBundle
from reflect_sampling_nerf_b200.train_path import TrainStep
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import synthetic_rays

SIZES = dict(num_coarse_samples=32, num_importance_samples=32, num_reflect_coarse_samples=16, num_reflect_importance_samples=16)
mode = sys.argv[1] if len(sys.argv) > 1 else "default"
torch.manual_seed(0)
model = ReflectSamplingNeRFModel(ReflectSamplingNeRFModelConfig(**SIZES)).cuda().train()
o, d, pa, img = [t.cuda() for t in synthetic_rays(512, 31, pixel_area=3.2e-6)]
stepper = TrainStep(model, graph=False)
log = []
orig = _lib.call
def call(name, *a):
    import threading
    log.append((name, threading.current_thread().name, torch.cuda.current_stream().cuda_stream, a[-1]))
    return orig(name, *a)
_lib.call = call
import reflect_sampling_nerf_b200.ops as ops
ops._lib.call = call
b = lambda: RayBundle(origins=o, directions=d, pixel_area=pa)
if mode == "side":
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            stepper._eager(b(), img)
    torch.cuda.current_stream().wait_stream(s)
else:
    for _ in range(3):
        stepper._eager(b(), img)
torch.cuda.synchronize()
log.clear()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        loss = stepper._eager(b(), img)
    print("capture ok", mode)
    g.replay(); torch.cuda.synchronize(); print("loss", float(loss))
except Exception as e:
    print("capture FAILED", mode, repr(e)[:300])
streams = {}
for name, th, cur, passed in log:
    streams.setdefault((th, cur, passed), []).append(name)
for k, v in streams.items():
    print(k, len(v), v[:4], "...", v[-2:])
