"""The oracle restatement (oracle/refpath.py) against vectors produced by the UNMODIFIED reference
files (tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from helpers import GOLDEN_SIZES, load_golden, oracle_model
from oracle import upstream as U


def _run(mode):
    g = load_golden(mode)
    model = oracle_model()
    model.train(mode == "train")
    checksum = float(sum(p.detach().double().abs().sum() for p in model.field.parameters()))
    assert abs(checksum - float(g["weight_checksum"])) < 1e-6 * checksum, \
        "seeded nn.Linear init no longer reproduces the golden weights (torch RNG changed?)"
    m = int(g["num_masked"])
    model.set_jitter(uniform=g["jit_uniform"], pdf=g["jit_pdf"],
                     reciprocal=g["jit_reciprocal"][:m], reflect_pdf=g["jit_reflect_pdf"][:m])
    bundle = U.RayBundle(origins=g["in_origins"], directions=g["in_directions"], pixel_area=g["in_pixel_area"])
    if mode == "train":
        out = model(bundle)
        loss = model.get_loss_dict(out, {"image": g["in_image"]})
        sum(loss.values()).backward()
    else:
        with torch.no_grad():
            out = model(bundle)
            loss = model.get_loss_dict(out, {"image": g["in_image"]})
    return g, model, out, loss


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_outputs_match_reference(mode):
    g, model, out, loss = _run(mode)
    keys = [k[4:] for k in g if k.startswith("out_")]
    assert sorted(keys) == sorted(out.keys())
    for k in keys:
        ref = g["out_" + k]
        got = out[k].detach()
        assert got.shape == ref.shape, k
        if ref.dtype == torch.bool:
            assert torch.equal(got, ref), k
        else:
            torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6, msg=lambda s, k=k: f"{k}: {s}")
    for k in [k for k in g if k.startswith("loss_")]:
        torch.testing.assert_close(loss[k[5:]].detach(), g[k], rtol=1e-5, atol=1e-7)


def test_gradients_match_reference():
    g, model, out, loss = _run("train")
    params = dict(model.field.named_parameters())
    for name, p in params.items():
        grad = p.grad if p.grad is not None else torch.zeros_like(p)
        ref_abs = float(g["gabs_" + name])
        assert abs(float(grad.double().abs().sum()) - ref_abs) <= 1e-4 * ref_abs + 1e-9, name
        if "grad_" + name in g:
            torch.testing.assert_close(grad, g["grad_" + name], rtol=1e-4, atol=1e-7, msg=lambda s, n=name: f"{n}: {s}")
    torch.testing.assert_close(params["mlp_base.layers.4.weight"].grad[::16, ::16],
                               g["grad_mlp_base.layers.4.weight[::16,::16]"], rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(params["mlp_mid.layers.0.weight"].grad[::8, ::8],
                               g["grad_mlp_mid.layers.0.weight[::8,::8]"], rtol=1e-4, atol=1e-7)
    # Q18: field_output_low never receives a gradient
    assert params["field_output_low.net.weight"].grad is None
