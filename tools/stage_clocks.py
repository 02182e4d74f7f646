#!/usr/bin/env python
"""Per-stage cycle breakdown of k_edge_bwd_sel (CTA 0), from a -DNB_STAGE_CLOCKS profiling build of the same sources.

    python tools/stage_clocks.py build      # here (nvcc): writes no-node-comparison_b200/libnbody_b200_clk.so
    python tools/stage_clocks.py run        # on the GPU box: EGNO N=20, T=10, L=4, B=256 training steps

The clocks are clock64() deltas of thread 0 between the stage boundaries marked NB_CLK(i) in csrc/nb_edge_sel.cuh.
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CLK_LIB = os.path.join(ROOT, "no-node-comparison_b200", "libnbody_b200_clk.so")

NAMES = {
    0: "unit read-out (after its side-MMA wait) / prologue", 1: "unit staging + sync", 2: "T0 geometry (+ ef loads)",
    3: "T0 wait side MMAs of previous tile", 4: "T0 selector row + sync", 5: "T0 issue gather", 6: "T0 wait gather",
    7: "S1 CUDA phase + sync", 8: "S1 issue MMA1", 9: "S1 wait MMA1", 10: "S2 CUDA phase + sync", 11: "S2 issue MMA2",
    12: "S2 wait MMA2", 13: "S3 CUDA phase + syncs", 14: "S3 issue dgrad3 + gM gather", 15: "S3 issue wgrad3 (side)",
    16: "S3 wait dgrad3", 17: "S4 tmem ld + mul", 18: "S4 wait wgrad3", 19: "S4 store + sync", 20: "S4 issue dgrad2",
    21: "S4 issue wgrad2 (side)", 22: "S4 wait dgrad2", 23: "S5 CUDA phase + syncs", 24: "S5 issue scatters (side)",
    25: "unit: wait side MMAs of the last tile", 26: "last unit read-out",
}


def build():
    from no_node_comparison_b200.build import build_library  # noqa
    print(build_library(defines=["NB_STAGE_CLOCKS"], out=CLK_LIB))


def run(B=256, N=20, T=10, L=4, steps=3):
    os.environ["NB_B200_LIBRARY"] = CLK_LIB
    import torch
    import no_node_comparison_b200 as nb
    from no_node_comparison_b200 import synth

    dev = torch.device("cuda:0")
    lib = nb.load_library()
    fn = lib.nb_debug_stage_clocks
    fn.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
    fn.restype = ctypes.c_int
    torch.manual_seed(1)
    m = nb.EGNO(n_layers=L, in_node_nf=2, in_edge_nf=2, hidden_nf=64, with_v=True, num_modes=2, num_timesteps=T, device=dev)
    row, col = synth.canonical_edges(B, N)
    s = synth.sample_state("charged", B, N, seed=0)
    x, nodes, ea, v, lm = synth.egno_features(s["loc"], s["vel"], s["charges"], row, col)
    x, nodes, ea, v, lm, row, col = [t.to(dev) for t in (x, nodes, ea, v, lm, row, col)]
    t_out = torch.arange(1, T + 1, device=dev)[None].repeat(B, 1)

    def step():
        m.zero_grad(set_to_none=True)
        xo, vo, ho = m(x, nodes, [row, col], ea, v=v, loc_mean=lm, timesteps_out=t_out)
        (xo.square().mean() + 1e-3 * ho.square().mean()).backward()

    step()
    buf = (ctypes.c_longlong * 32)()
    fn(buf, 1)
    gfn = lib.nb_debug_gemm_clocks
    gfn.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
    gfn.restype = ctypes.c_int
    gbuf = (ctypes.c_longlong * 16)()
    gfn(gbuf, 1)
    lib.nb_profile_enable(1)
    for _ in range(steps):
        step()
    fn(buf, 0)
    gfn(gbuf, 0)
    gn = max(gbuf[0], 1)
    GN = {1: "first tile's loads issued", 2: "mbarrier init + TMEM alloc", 3: "weight staging + sync", 4: "split + store tile (waits for its loads)",
          5: "fence + sync", 6: "MMA issue + commit", 7: "prefetch next tile / epilogue operands", 8: "wait MMA", 9: "TMEM load", 10: "epilogue + stores",
          11: "final sync", 12: "TMEM dealloc"}
    print(f"k_gemm64_tc, CTA (0,0): {gn} launches, {sum(gbuf[1:13]) / gn:.0f} cycles per launch")
    for i in range(1, 13):
        print(f"  {i:2d} {gbuf[i] / gn:9.0f} cyc/launch  {GN[i]}")
    tot = sum(buf)
    launches = steps * L
    units = -(-T * B // 148)
    tiles = units * -(-N * (N - 1) // 128) if N <= 27 else None
    print(f"k_edge_bwd_sel, CTA 0: {launches} launches, {tot / launches:.0f} cycles per launch"
          + (f", {units} units, {tiles} tiles per launch" if tiles else ""))
    print(f"  whole CTA: slowest CTA of any launch {buf[27]} cycles, mean over CTAs and launches {buf[28] / launches / 148:.0f}"
          f" = {buf[29] / launches / 148 / 1e3:.1f} us (globaltimer): effective SM clock {buf[28] / max(buf[29], 1):.3f} GHz")
    import ctypes as C
    ms = (C.c_double * 8)(); cnt = (C.c_longlong * 8)()
    lib.nb_profile_read(ms, cnt)
    print("  CUDA-event times per category (ms per launch):", [round(ms[i] / max(cnt[i], 1), 4) for i in range(8)], list(cnt))
    for i in range(27):
        if buf[i]:
            per = f"{buf[i] / launches / tiles:8.0f} /tile" if tiles else ""
            print(f"  {i:2d} {100 * buf[i] / tot:5.1f}% {buf[i] / launches:10.0f} cyc/launch {per}  {NAMES.get(i, '')}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        run()
