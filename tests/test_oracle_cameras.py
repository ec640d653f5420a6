"""CPU: the camera / pixel-sampling oracle (oracle/cameras.py).  nerfstudio is absent (parity unpinned):
the checks are analytic properties of Cameras.generate_rays plus the distance between the literal tensor form and the
explicit-operation-order form that csrc/raygen.cu reproduces bit for bit."""
import math

import torch

from oracle import cameras as C
from reflect_sampling_nerf_b200.data import orbit_cameras


def _cams(v=7, h=40, w=50):
    return orbit_cameras(v, 4.0, h, w, 0.6911112070083618, seed=3)


def test_pixel_sampling_floor_of_scaled_uniforms():
    g = torch.Generator().manual_seed(0)
    rand = torch.rand(1000, 3, generator=g)
    pix = C.sample_pixels(rand, 7, 40, 50)
    assert pix.dtype == torch.int64
    assert int(pix[:, 0].max()) < 7 and int(pix[:, 1].max()) < 40 and int(pix[:, 2].max()) < 50
    assert torch.equal(pix[:, 1], torch.floor(rand[:, 1] * 40).long())


def test_rays_are_unit_look_at_the_origin_and_carry_the_pixel_footprint():
    cams = _cams()
    h, w = cams.height, cams.width
    centre = torch.tensor([[v, h // 2, w // 2] for v in range(len(cams))])
    o, d, area = C.generate_rays(cams.camera_to_worlds, cams.fx, cams.fy, cams.cx, cams.cy, centre)
    torch.testing.assert_close(torch.linalg.norm(d, dim=-1), torch.ones(len(cams)), rtol=0, atol=1e-6)
    torch.testing.assert_close(torch.linalg.norm(o, dim=-1), torch.full((len(cams),), 4.0), rtol=0, atol=1e-5)
    # the central pixel looks (half a pixel off) at the origin: direction ~ -origin / 4
    assert float((d * (-o / 4.0)).sum(-1).min()) > 1 - 1e-3
    # footprint of a central pixel of a pinhole camera: (1 / f)^2
    f = float(cams.fx[0])
    torch.testing.assert_close(area[:, 0], torch.full((len(cams),), 1.0 / f ** 2), rtol=2e-2, atol=0)
    assert abs(f - 0.5 * w / math.tan(0.5 * 0.6911112070083618)) < 1e-3


def test_explicit_order_form_is_within_two_ulp_of_the_literal_form():
    cams = _cams()
    g = torch.Generator().manual_seed(1)
    pix = C.sample_pixels(torch.rand(4096, 3, generator=g), len(cams), cams.height, cams.width)
    a = C.generate_rays_upstream(cams.camera_to_worlds, cams.fx, cams.fy, cams.cx, cams.cy, pix)
    b = C.generate_rays(cams.camera_to_worlds, cams.fx, cams.fy, cams.cx, cams.cy, pix)
    assert torch.equal(a[0], b[0])
    assert float((a[1] - b[1]).abs().max()) <= 2 ** -22           # unit vectors: reduction order + ATen's 1-ulp CPU sqrt
    torch.testing.assert_close(a[2], b[2], rtol=1e-4, atol=0)     # |d - d_x| |d - d_y| amplifies the ulp of d by ~1 / pixel size
    # the explicit form is plain IEEE fp32 arithmetic: numpy reproduces it bit for bit
    import numpy as np
    w = b[1].numpy()
    assert np.abs(np.sqrt((w * w).sum(-1, dtype=np.float64)) - 1).max() < 1e-6


def test_target_gather_blends_rgba_onto_white():
    imgs = torch.zeros(2, 4, 4, 4, dtype=torch.uint8)
    imgs[0, 1, 2] = torch.tensor([255, 0, 0, 255], dtype=torch.uint8)      # opaque red
    imgs[1, 3, 0] = torch.tensor([0, 0, 255, 128], dtype=torch.uint8)      # half-transparent blue
    pix = torch.tensor([[0, 1, 2], [1, 3, 0], [0, 0, 0]])
    out = C.gather_targets(imgs, pix)
    torch.testing.assert_close(out[0], torch.tensor([1.0, 0.0, 0.0]))
    a = 128 / 255
    torch.testing.assert_close(out[1], torch.tensor([1 - a, 1 - a, a + (1 - a)]), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(out[2], torch.ones(3))                     # transparent -> white background
