import sys, os, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch.nn.functional as F
from oracle import refpath as R
from test_field_train_gpu import _run_mine, _setup, _cos
for scale in (0.02, 0.5, 3.0):
    n, s = 64, 64
    field, o, d, pa, bins, g = _setup(n, s, 31, "reciprocal", 1.0)
    pa = (torch.rand(n, 1, generator=g) * scale + 1e-4).requires_grad_(True)
    ex = lambda x: x[:, None, :].expand(n, s, x.shape[-1])
    mean, cov = R.frustum_gaussian(ex(o), ex(d), bins[:, :-1, None], bins[:, 1:, None], ex(pa))
    mean, cov = R.contract(mean, cov)
    ref = field.point_heads(mean, cov, ex(d), primary=False)
    g_feat = torch.zeros(n, s, 16)
    g_feat[..., 0:3] = torch.randn(n, s, 3, generator=g) * 0.1
    (g_feat[..., 0:3] * ref["rgb"]).sum().backward()
    _, _, _, _, grads, g_area = _run_mine(field, 0, o, d, pa.detach(), bins, torch.zeros(n, s), g_feat, True)
    got = g_area.cpu().sum(-1); refg = pa.grad[:, 0]
    print("refl scale", scale, "cos", _cos(got, refg), "ratio", float(got.norm() / refg.norm()), "sum ratio", float(got.sum()/refg.sum()))
    # per-sample-index breakdown
    field.zero_grad()
for scale in (0.02, 0.5):
    torch.manual_seed(5)
    field = R.OracleField().train()
    m = 512
    w = F.normalize(torch.randn(m, 3), dim=-1)
    sq = (torch.rand(m, 1) * scale + 1e-5).requires_grad_(True)
    g_rgb = torch.randn(m, 3) * 0.1
    (g_rgb * field.inf_color(w, sq)).sum().backward()
    g_feat = torch.zeros(m, 1, 16); g_feat[:, 0, 0:3] = g_rgb
    _, _, _, _, grads, g_area = _run_mine(field, 1, None, w, sq.detach(), None, None, g_feat, True)
    got, refg = g_area.cpu()[:, 0], sq.grad[:, 0]
    print("inf scale", scale, "cos", _cos(got, refg), "ratio", float(got.norm() / refg.norm()))
