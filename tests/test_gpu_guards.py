"""Out-of-bounds evidence without compute-sanitizer (closed on this pool): every caller-owned buffer of the C ABI is carved
out of one arena with sentinel guard bands on both sides; after a forward + backward through the raw entry points (the
product kernel sequence: fused node kernels, selector edge kernels, SEGNO chain, reductions) every guard word must be
untouched and every output fully written (no sentinel left inside)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 1024          # floats on each side of every buffer
SENT = -1.2345678e30  # sentinel


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


class Arena:
    def __init__(self, sizes, dev):
        self.off, n = {}, GUARD
        for k, s in sizes.items():
            s = (int(s) + 63) // 64 * 64
            self.off[k] = (n, int(sizes[k]))
            n += s + GUARD
        self.buf = torch.full((n,), SENT, device=dev, dtype=torch.float32)
        self.mask = torch.ones(n, dtype=torch.bool, device=dev)      # True = guard word
        for k, (o, s) in self.off.items():
            self.mask[o:o + s] = False

    def view(self, k):
        o, s = self.off[k]
        return self.buf[o:o + s]

    def ptr(self, k):
        return ctypes.c_void_p(self.view(k).data_ptr())

    def guards_intact(self):
        return bool((self.buf[self.mask] == SENT).all())


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize("B,N,T,L,modes,nin", [(3, 20, 10, 2, 2, 1), (2, 37, 4, 1, 2, 1), (5, 5, 8, 2, 2, 1), (2, 5, 6, 2, 4, 1),
                                                (3, 5, 10, 2, 2, 3), (1, 100, 2, 1, 1, 1), (2, 20, 16, 1, 2, 2)])
def test_egno_c_abi_never_writes_outside_its_buffers(B, N, T, L, modes, nin):
    """whole-graph and blocked selector walks, the fused (modes <= 2) and the general temporal convolution, several input
    frames, the longest supported horizon"""
    import no_node_comparison_b200 as nb
    from no_node_comparison_b200 import _cabi, synth
    d = _dev()
    lib = nb.load_library()
    row, col = synth.canonical_edges(B, N)
    fr = []
    for i in range(nin):
        s = synth.sample_state("charged", B, N, seed=7 + i)
        fr.append([t.to(d) for t in synth.egno_features(s["loc"], s["vel"], s["charges"], row, col)])
    x, nodes, ea, v, lm = [(torch.stack([f[k] for f in fr]) if nin > 1 else fr[0][k]).contiguous() for k in range(5)]
    cfg = _cabi.NbEgnoConfig(B, N, T, L, modes, 2, 2, 32, 1, nin)
    npar = lib.nb_egno_param_count(ctypes.byref(cfg))
    assert npar > 0, lib.nb_last_error()
    params = (0.1 * torch.randn(npar, generator=torch.Generator().manual_seed(1))).to(d)
    ts = torch.arange(1, T + 1, device=d)[None].repeat(B, 1).contiguous()
    ts_in = torch.arange(-nin + 1, 1, device=d)[None].repeat(B, 1).contiguous() if nin > 1 else None
    Nn = T * B * N
    st = ctypes.c_void_p(torch.cuda.current_stream(d).cuda_stream)
    A = Arena(dict(x_out=Nn * 3, v_out=Nn * 3, h_out=Nn * 64, saved=lib.nb_egno_saved_floats(ctypes.byref(cfg)),
                   ws_f=lib.nb_egno_workspace_floats(ctypes.byref(cfg), 2), ws_b=lib.nb_egno_workspace_floats(ctypes.byref(cfg), 1),
                   grad=npar, gx_in=nin * B * N * 3, gv_in=nin * B * N * 3), d)
    rc = lib.nb_egno_forward(ctypes.byref(cfg), _p(params), _p(x), _p(nodes), _p(ea), _p(v), _p(lm), _p(ts), _p(ts_in), A.ptr("x_out"),
                             A.ptr("v_out"), A.ptr("h_out"), A.ptr("saved"), A.ptr("ws_f"), st)
    assert rc == 0, lib.nb_last_error()
    torch.cuda.synchronize()
    assert A.guards_intact()
    for k in ("x_out", "v_out", "h_out"):
        assert not bool((A.view(k) == SENT).any()), k
        assert bool(torch.isfinite(A.view(k)).all()), k
    g = torch.Generator().manual_seed(2)
    Gx, Gv, Gh = [torch.randn(Nn, c, generator=g).to(d) for c in (3, 3, 64)]
    rc = lib.nb_egno_backward(ctypes.byref(cfg), _p(params), _p(nodes), _p(ea), _p(lm), _p(ts), _p(ts_in), A.ptr("saved"), _p(Gx), _p(Gv),
                              _p(Gh), A.ptr("grad"), A.ptr("gx_in"), A.ptr("gv_in"), A.ptr("ws_b"), st)
    assert rc == 0, lib.nb_last_error()
    torch.cuda.synchronize()
    assert A.guards_intact()
    for k in ("grad", "gx_in", "gv_in"):
        assert not bool((A.view(k) == SENT).any()), k
        assert bool(torch.isfinite(A.view(k)).all()), k


@pytest.mark.parametrize("B,N,T,h_given", [(3, 20, 10, 0), (4, 5, 6, 1), (2, 27, 3, 0)])
def test_segno_c_abi_never_writes_outside_its_buffers(B, N, T, h_given):
    import no_node_comparison_b200 as nb
    from no_node_comparison_b200 import _cabi, synth
    d = _dev()
    lib = nb.load_library()
    s = synth.sample_state("gravity", B, N, seed=9)
    row, col = synth.canonical_edges(B, N)
    his, x, v, ea = [t.to(d).contiguous() for t in synth.segno_features(s["loc"], s["vel"], s["charges"], row, col)]
    if h_given:
        his = torch.randn(B * N, 64, generator=torch.Generator().manual_seed(4)).to(d)
    cfg = _cabi.NbSegnoConfig(B, N, T, 1, 2, 1, 1.0, h_given)
    npar = lib.nb_segno_param_count(ctypes.byref(cfg))
    params = (0.1 * torch.randn(npar, generator=torch.Generator().manual_seed(1))).to(d)
    Nn = B * N
    st = ctypes.c_void_p(torch.cuda.current_stream(d).cuda_stream)
    A = Arena(dict(x_out=Nn * 3, h_out=Nn * 64, v_out=Nn * 3, saved=lib.nb_segno_saved_floats(ctypes.byref(cfg)),
                   ws_f=lib.nb_segno_workspace_floats(ctypes.byref(cfg), 2), ws_b=lib.nb_segno_workspace_floats(ctypes.byref(cfg), 1),
                   grad=npar, gx_in=Nn * 3, gv_in=Nn * 3, gh_in=Nn * 64), d)
    rc = lib.nb_segno_forward(ctypes.byref(cfg), _p(params), _p(his), _p(x), _p(v), _p(ea), A.ptr("x_out"), A.ptr("h_out"),
                              A.ptr("v_out"), A.ptr("saved"), A.ptr("ws_f"), st)
    assert rc == 0, lib.nb_last_error()
    torch.cuda.synchronize()
    assert A.guards_intact()
    for k in ("x_out", "v_out", "h_out"):
        assert not bool((A.view(k) == SENT).any()), k
    g = torch.Generator().manual_seed(2)
    Gx, Gh, Gv = [torch.randn(Nn, c, generator=g).to(d) for c in (3, 64, 3)]
    rc = lib.nb_segno_backward(ctypes.byref(cfg), _p(params), _p(his), _p(ea), A.ptr("saved"), _p(Gx), _p(Gh), _p(Gv), A.ptr("grad"),
                               A.ptr("gx_in"), A.ptr("gv_in"), A.ptr("gh_in") if h_given else None, A.ptr("ws_b"), st)
    assert rc == 0, lib.nb_last_error()
    torch.cuda.synchronize()
    assert A.guards_intact()
    for k in ("grad", "gx_in", "gv_in") + (("gh_in",) if h_given else ()):
        assert not bool((A.view(k) == SENT).any()), k
        assert bool(torch.isfinite(A.view(k)).all()), k
