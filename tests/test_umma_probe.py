"""tcgen05 building blocks (csrc/umma.cuh) through single-tile GEMM probes vs torch matmul on the same
bf16 operands.  Validates: 128B-swizzle block images, K-major and MN-major smem descriptors, the
instruction descriptor, TMEM lane/column mapping, bulk-copy staging, N-split at a TMEM column offset."""
import pytest
import torch

from reflect_sampling_nerf_b200 import _lib
from reflect_sampling_nerf_b200.blocks import pack_blocks

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _test_build():
    """The probes live in the test build (librsn_b200_dbg.so, include/rsn_b200_test.h) only."""
    _lib.use_dbg(True)
    yield
    _lib.use_dbg(False)


@pytest.mark.parametrize("N,KB,split", [(256, 4, 1), (256, 4, 2), (128, 2, 1), (16, 4, 1), (256, 1, 1), (64, 3, 2)])
def test_kmajor_tile_gemm(N, KB, split):
    g = torch.Generator().manual_seed(N + KB)
    x = torch.randn(128, KB * 64, generator=g).bfloat16()
    w = torch.randn(N, KB * 64, generator=g).bfloat16()
    xb, wb = pack_blocks(x).cuda(), pack_blocks(w).cuda()
    out = torch.full((128, N), float("nan"), device="cuda")
    _lib.call("rsn_probe_umma_kmajor", xb.data_ptr(), wb.data_ptr(), N, KB, split, out.data_ptr(), _lib.stream())
    torch.cuda.synchronize()
    ref = x.float() @ w.float().T
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("NB", [1, 2, 4])
def test_mnmajor_tile_gemm(NB):
    g = torch.Generator().manual_seed(NB)
    u = torch.randn(128, 128, generator=g).bfloat16()       # [points, M]
    v = torch.randn(128, NB * 64, generator=g).bfloat16()   # [points, N]
    ub, vb = pack_blocks(u).cuda(), pack_blocks(v).cuda()
    out = torch.full((128, NB * 64), float("nan"), device="cuda")
    _lib.call("rsn_probe_umma_mnmajor", ub.data_ptr(), vb.data_ptr(), 2, NB, out.data_ptr(), _lib.stream())
    torch.cuda.synchronize()
    ref = u.float().T @ v.float()
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("NB", [1, 2, 4])
def test_mnmajor_chunk_major_tile_gemm(NB):
    """The wgrad's read of chunk-major stash blocks: NO-swizzle MN-major descriptors, LBO = 128 (next 8 points), SBO = 1024
    (next 8 features); the swapped pair must NOT give the product (pins which field carries which stride)."""
    from reflect_sampling_nerf_b200.blocks import pack_blocks_cm
    g = torch.Generator().manual_seed(10 + NB)
    u = torch.randn(128, 128, generator=g).bfloat16()       # [points, M]
    v = torch.randn(128, NB * 64, generator=g).bfloat16()   # [points, N]
    ub, vb = pack_blocks_cm(u)[0].cuda(), pack_blocks_cm(v)[0].cuda()
    ref = u.float().T @ v.float()
    out = torch.full((128, NB * 64), float("nan"), device="cuda")
    _lib.call("rsn_probe_umma_mnmajor_cm", ub.data_ptr(), vb.data_ptr(), NB, 128, 1024, out.data_ptr(), 0, None, _lib.stream())
    torch.cuda.synchronize()
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-3, atol=1e-2)
    _lib.call("rsn_probe_umma_mnmajor_cm", ub.data_ptr(), vb.data_ptr(), NB, 1024, 128, out.data_ptr(), 0, None, _lib.stream())
    torch.cuda.synchronize()
    assert (out.cpu() - ref).abs().max() > 1.0


@pytest.mark.parametrize("N,KB", [(256, 4), (256, 1), (128, 2), (16, 4), (64, 3)])
def test_cta_pair_tile_gemm(N, KB):
    """tcgen05.mma.cta_group::2: M = 256 over a CTA pair, each CTA staging its 128 rows of X and N/2 rows of W."""
    g = torch.Generator().manual_seed(N * 7 + KB)
    x = torch.randn(256, KB * 64, generator=g).bfloat16()
    w = torch.randn(N, KB * 64, generator=g).bfloat16()
    xb = torch.stack([pack_blocks(x[:128]), pack_blocks(x[128:])]).cuda()      # [2 tiles][KB][128][128 B]
    wb = pack_blocks(w).cuda()
    out = torch.full((256, N), float("nan"), device="cuda")
    _lib.call("rsn_probe_umma_2cta", xb.data_ptr(), wb.data_ptr(), N, KB, out.data_ptr(), _lib.stream())
    torch.cuda.synchronize()
    ref = x.float() @ w.float().T
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("N,KB", [(256, 4), (256, 1), (128, 2), (16, 4), (64, 3)])
def test_a_from_tmem_tile_gemm(N, KB):
    """tcgen05.mma with the A operand in TMEM (written by tcgen05.st as packed bf16 pairs): the form the fused field
    kernels use to hand one layer's activations to the next without a shared-memory round trip."""
    g = torch.Generator().manual_seed(N * 3 + KB)
    x = torch.randn(128, KB * 64, generator=g).bfloat16()
    w = torch.randn(N, KB * 64, generator=g).bfloat16()
    xb, wb = pack_blocks(x).cuda(), pack_blocks(w).cuda()
    out = torch.full((128, N), float("nan"), device="cuda")
    cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
    _lib.call("rsn_probe_umma_ts", xb.data_ptr(), wb.data_ptr(), N, KB, out.data_ptr(), 0, cyc.data_ptr(), _lib.stream())
    torch.cuda.synchronize()
    ref = x.float() @ w.float().T
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-3, atol=1e-2)
