"""tcgen05 building blocks on the device: the three MMA forms used by the edge tiles (forward, data-gradient,
weight-gradient) with split-bf16 operands and fp32 accumulation in TMEM, against an fp64 matmul."""
import ctypes

import pytest
import torch

import no_node_comparison_b200 as nb

pytestmark = pytest.mark.gpu


def _run(mode, A, W):
    lib = nb.load_library()
    d = torch.device("cuda:0")
    Ad, Wd = A.to(d).contiguous(), W.to(d).contiguous()
    out = torch.zeros(2 * 128 * 64, device=d)  # modes >= 4 append cycle counts (mode 10: columns 64..71) behind the dump
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.nb_tc_selftest(mode, P(Ad), P(Wd), P(out), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.nb_last_error()
    torch.cuda.synchronize()
    if mode == 10:
        return out[:128 * 64].reshape(128, 64).cpu(), out[128 * 64:].reshape(128, 64)[:, :8].cpu()
    if mode >= 4:
        print(f"tcgen05 selftest mode {mode}: one 12-MMA group {out[8192].item():.0f} cycles, four groups {out[8193].item():.0f}")
    return out[:128 * 64].reshape(128, 64).cpu()


def _relerr(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


def test_tcgen05_forward_form():
    g = torch.Generator().manual_seed(0)
    A, W = torch.randn(128, 64, generator=g), torch.randn(64, 64, generator=g)
    out = _run(0, A, W)
    ref = A.double() @ W.double().t()
    assert _relerr(out, ref) < 3e-5, _relerr(out, ref)


def test_tcgen05_dgrad_form():
    g = torch.Generator().manual_seed(1)
    A, W = torch.randn(128, 64, generator=g), torch.randn(64, 64, generator=g)
    out = _run(1, A, W)
    ref = A.double() @ W.double()
    assert _relerr(out, ref) < 3e-5, _relerr(out, ref)


def test_tcgen05_wgrad_form_and_m64_lane_layout():
    g = torch.Generator().manual_seed(2)
    A, G = torch.randn(128, 64, generator=g), torch.randn(128, 64, generator=g)
    out = _run(2, A, G)
    ref = A.double().t() @ G.double()          # [64, 64]
    # M = 64 accumulators live in TMEM lanes (i % 16) + 32 * (i // 16)   (cute tmem_frg, M_MMA == 64)
    lanes = torch.tensor([(i % 16) + 32 * (i // 16) for i in range(64)])
    assert _relerr(out[lanes], ref) < 3e-5, _relerr(out[lanes], ref)


def test_tcgen05_column_sum_form_noswizzle_operand():
    g = torch.Generator().manual_seed(3)
    A = torch.randn(128, 64, generator=g)
    out = _run(3, A, torch.zeros(64, 64))
    ref = A.double().sum(0)                    # [64]
    lanes = torch.tensor([(i % 16) + 32 * (i // 16) for i in range(64)])
    for col in range(8):
        assert _relerr(out[lanes, col], ref) < 3e-5, (col, _relerr(out[lanes, col], ref))


@pytest.mark.parametrize("mode", [4, 5, 6, 7])
def test_tcgen05_a_operand_in_tensor_memory(mode):
    """modes 4 / 5: forward / data-gradient forms with the A operand read from TMEM (packed bf16 pairs, one row per
    lane); modes 6 / 7: the shared-memory-A forms through the same timed path."""
    g = torch.Generator().manual_seed(10 + mode)
    A, W = torch.randn(128, 64, generator=g), torch.randn(64, 64, generator=g)
    out = _run(mode, A, W)
    ref = A.double() @ (W.double().t() if mode in (4, 6) else W.double())
    assert _relerr(out, ref) < 3e-5, _relerr(out, ref)


def test_tcgen05_timed_wgrad_and_column_sum_forms():
    """modes 8 / 9 = modes 2 / 3 through the timed path (prints the cycles of the 24- and 16-MMA groups)."""
    g = torch.Generator().manual_seed(3)
    A, G = torch.randn(128, 64, generator=g), torch.randn(128, 64, generator=g)
    out = _run(8, A, G)
    ref = A.double().t() @ G.double()
    rows = torch.tensor([(o % 16) + 32 * (o // 16) for o in range(64)])
    assert _relerr(out[rows], ref) < 3e-5
    out = _run(9, A, G)
    assert _relerr(out[rows, 0], A.double().sum(0)) < 3e-5


def test_tcgen05_wgrad_form_with_folded_column_sums():
    """mode 10: D[64 x 72] = A^T [G | 1] — an N = 72 MN-major B operand whose second block (LBO) is an all-ones tile, so
    the bias gradient (column sums of A) rides in the weight-gradient MMAs."""
    g = torch.Generator().manual_seed(4)
    A, G = torch.randn(128, 64, generator=g), torch.randn(128, 64, generator=g)
    out, tail = _run(10, A, G)
    rows = torch.tensor([(o % 16) + 32 * (o // 16) for o in range(64)])
    assert _relerr(out[rows], A.double().t() @ G.double()) < 3e-5
    for c in range(8):
        assert _relerr(tail[rows, c], A.double().sum(0)) < 3e-5, c


def test_silu_with_the_reciprocal_on_the_fma_pipe_matches_the_mufu_variant():
    """nb_silu_fma (ex2.approx + bit-trick seed + three Newton steps; the forward edge tile uses it to halve its MUFU load)
    against float64 SiLU and against the two-MUFU variant, over the whole range incl. the saturating ends."""
    import ctypes

    import no_node_comparison_b200 as nb
    lib = nb.load_library()
    d = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    x = torch.cat([torch.randn(1 << 20, generator=g) * 4, torch.linspace(-120, 120, 100001), torch.tensor([0.0, -0.0, 88.0, -88.0, -104.0, 1e4, -1e4])]).to(d)
    a, b = torch.empty_like(x), torch.empty_like(x)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    assert lib.nb_silu_selftest(x.numel(), P(x), P(a), P(b), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    ref = (x.double() * torch.sigmoid(x.double()))
    scale = ref.abs().clamp_min(1e-30)
    err_mufu = ((a.double() - ref).abs() / torch.maximum(scale, torch.full_like(scale, 1e-3))).max().item()
    err_fma = ((b.double() - ref).abs() / torch.maximum(scale, torch.full_like(scale, 1e-3))).max().item()
    print(f"SiLU max rel. error vs float64: two-MUFU {err_mufu:.2e}, FMA-pipe reciprocal {err_fma:.2e}")
    assert torch.isfinite(b).all()
    assert err_fma < 1e-6 and err_fma < 2 * err_mufu + 2e-7
