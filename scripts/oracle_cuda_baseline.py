"""Informational baseline (NOT part of bench.py's contract): the oracle restatement of the reference's eager
PyTorch path run on the GPU -- the "reference PyTorch-CUDA rays/sec" that north_star's >= 10x target is quoted
against (the reference itself cannot be imported here: nerfstudio is absent).
usage: python scripts/oracle_cuda_baseline.py [n_rays] [train|render] [fp32|bf16|fp16]"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import upstream as U
from oracle.refpath import OracleModel

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
workload = sys.argv[2] if len(sys.argv) > 2 else "train"
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
torch.manual_seed(0)
model = OracleModel(**bench.CFG).cuda()
model.train(workload == "train")
opt = torch.optim.RAdam(model.field.parameters(), lr=1e-3, eps=1e-15)
o, d, a, img = [t.cuda() for t in bench.synthetic_batch(n, 0)]
ctx = torch.autocast("cuda", dtype={"bf16": torch.bfloat16, "fp16": torch.float16}.get(prec)) if prec != "fp32" else torch.autocast("cuda", enabled=False)


def step():
    b = U.RayBundle(origins=o, directions=d, pixel_area=a)
    if workload == "train":
        opt.zero_grad(set_to_none=True)
        with ctx:
            out = model(b)
            loss = sum(model.get_loss_dict(out, {"image": img}).values())
        loss.backward()
        opt.step()
    else:
        with torch.no_grad(), ctx:
            model(b)


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
k = 5
for _ in range(k):
    step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / k
print(f"oracle on cuda ({prec}, {workload}): {n} rays/step, {dt*1e3:.1f} ms/step, {n/dt:.0f} rays/s, "
      f"peak memory {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
