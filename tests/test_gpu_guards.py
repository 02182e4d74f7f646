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


@pytest.mark.parametrize("mode,n,nf", [(0, 37, 2), (1, 100, 2), (2, 333, 3), (2, 5, 2)])
def test_segno_merge_and_embedding_entry_points_stay_inside_their_buffers(mode, n, nf):
    import no_node_comparison_b200 as nb
    from no_node_comparison_b200 import _cabi
    d = _dev()
    lib = nb.load_library()
    g = torch.Generator().manual_seed(5)
    rnd = lambda *s: torch.randn(*s, generator=g).to(d).contiguous()
    frame = nf - 1
    h_all, x_all, v_all = rnd(n, nf, 64), rnd(n, nf, 3), rnd(n, nf, 3)
    h_int, x_int, v_int = rnd(n, 64), rnd(n, 3), rnd(n, 3)
    ap = (0.2 * torch.randn(64 * 65 + 129, generator=g)).to(d)
    st = ctypes.c_void_p(torch.cuda.current_stream(d).cuda_stream)
    cfg = _cabi.NbSegnoConfig(1, 5, 2, 1, 2, 1, 1.0, 1)      # the embedding entry points take their row count separately
    A = Arena(dict(h_out=n * 64, x_out=n * 3, v_out=n * 3, alpha=n * 2, g_h_all=n * nf * 64, g_x_all=n * nf * 3, g_v_all=n * nf * 3,
                   g_h_int=n * 64, g_x_int=n * 3, g_v_int=n * 3, g_attn=64 * 65 + 129,
                   ws=max(lib.nb_segno_merge_backward_workspace_floats(n), 64), emb=n * nf * 64,
                   ews=lib.nb_segno_embed_backward_workspace_floats(ctypes.byref(cfg), n * nf),
                   gpar=lib.nb_segno_param_count(ctypes.byref(cfg))), d)
    intp = (None, None, None) if mode == 0 else (_p(h_int), _p(x_int), _p(v_int))
    rc = lib.nb_segno_merge_forward(mode, n, nf, frame, _p(h_all), _p(x_all), _p(v_all), *intp, _p(ap) if mode == 2 else None,
                                    A.ptr("h_out"), A.ptr("x_out"), A.ptr("v_out"), A.ptr("alpha") if mode == 2 else None, st)
    assert rc == 0, lib.nb_last_error()
    gint = (None, None, None) if mode == 0 else (A.ptr("g_h_int"), A.ptr("g_x_int"), A.ptr("g_v_int"))
    rc = lib.nb_segno_merge_backward(mode, n, nf, frame, _p(h_all), _p(x_all), _p(v_all), *intp, _p(ap) if mode == 2 else None,
                                     A.ptr("alpha") if mode == 2 else None, _p(rnd(n, 64)), _p(rnd(n, 3)), _p(rnd(n, 3)),
                                     A.ptr("g_h_all"), A.ptr("g_x_all"), A.ptr("g_v_all"), *gint,
                                     A.ptr("g_attn") if mode == 2 else None, 0, A.ptr("ws") if mode == 2 else None, st)
    assert rc == 0, lib.nb_last_error()
    params = (0.1 * torch.randn(A.off["gpar"][1], generator=g)).to(d)
    his = rnd(n * nf, 1)
    rc = lib.nb_segno_embed_forward(ctypes.byref(cfg), _p(params), n * nf, _p(his), A.ptr("emb"), st)
    assert rc == 0, lib.nb_last_error()
    rc = lib.nb_segno_embed_backward(ctypes.byref(cfg), n * nf, _p(his), _p(rnd(n * nf, 64)), A.ptr("gpar"), A.ptr("ews"), st)
    assert rc == 0, lib.nb_last_error()
    torch.cuda.synchronize()
    assert A.guards_intact()
    for k in ("h_out", "x_out", "v_out", "emb"):
        assert not bool((A.view(k) == SENT).any()), k
    # only the observed frame's slice of the *_all gradients is written
    gh = A.view("g_h_all").view(n, nf, 64)
    assert not bool((gh[:, frame] == SENT).any()) and bool((gh[:, :frame] == SENT).all())
    if mode == 2:
        assert not bool((A.view("g_attn") == SENT).any()) and not bool((A.view("alpha") == SENT).any())
