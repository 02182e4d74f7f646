#!/usr/bin/env python
"""bench.py — throughput of the EGNO / SEGNO hot path on B200 (BASELINE.json metric: trajectories/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU implementation (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU (weak scaling: B per GPU fixed)

A "step" is one training pass (forward + MSE loss + backward + Adam) of the 20-body EGNO (BASELINE.json
configs[2]: N=20, num_timesteps=10, hidden 64, 4 layers) over one batch of B=256 synthetic trajectories
(main.py:32 default batch) per GPU.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

_STDOUT_FD = None


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _STDOUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_STDOUT_FD, data)


METRIC = "train trajectories/s (EGNO 20-body, fwd+bwd+Adam)"
UNIT = "trajectories/s"

# per-edge MACs of the reference formulas: phi_e = 131*64 + 64*64, phi_x = 64*64 + 64
MAC_EDGE_REF = 131 * 64 + 64 * 64 + 64 * 64 + 64
# MACs the fused edge kernel itself executes per edge (first layer's h-part runs per node, outside it)
MAC_EDGE_KERNEL = 64 * 64 + 64 * 64 + 64 + 3 * 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="trajectories per GPU per step")
    ap.add_argument("--n-balls", type=int, default=20)
    ap.add_argument("--timesteps", type=int, default=10)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--cpu-sample", type=int, default=32, help="trajectories per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the SEGNO / inference side numbers")
    ap.add_argument("--quick", action="store_true", help="device-resident timing only (for profiler runs)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying one CUDA "
                                                            "graph per training step")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, window=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = self.lines
        scope = "whole run (sampler start to stop)"
        if window is not None:
            inside = [x for x in lines if window[0] <= x[0] <= window[1] + 0.05]
            if inside:
                lines, scope = inside, "timed region"
            else:
                scope = "warm-up + timed region (timed region shorter than one sampling period)"
        for _, ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_arm(args, steps, warmup, sample):
    """The reference's CPU implementation of the path: the oracle port (oracle/nbody_oracle.py, pinned to the
    reference by golden vectors; /root/reference itself is not on the GPU box) with all host threads."""
    from oracle import nbody_oracle as O
    import no_node_comparison_b200 as nb
    from no_node_comparison_b200 import synth

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    N, T, L = args.n_balls, args.timesteps, args.layers
    torch.manual_seed(1)
    holder = nb.EGNO(n_layers=L, in_node_nf=2, in_edge_nf=2, hidden_nf=64, with_v=True, num_modes=2, num_timesteps=T,
                     device="cpu")
    p = {k: t.detach().clone().requires_grad_(True) for k, t in holder.state_dict().items()}
    opt = torch.optim.Adam(list(p.values()), lr=1e-4, weight_decay=1e-8)
    s = synth.sample_state("charged", sample, N, seed=0)
    row, col = synth.canonical_edges(sample, N)
    x, nodes, ea, v, lm = synth.egno_features(s["loc"], s["vel"], s["charges"], row, col)
    t_out = torch.arange(1, T + 1)[None].repeat(sample, 1)
    target = x.repeat(T, 1) + 0.05 * torch.randn(T * sample * N, 3, generator=torch.Generator().manual_seed(1))

    def step():
        opt.zero_grad()
        xo, _, _ = O.egno_forward(p, x, nodes, row, col, ea, v, lm, t_out, n_layers=L, num_timesteps=T)
        loss = ((xo - target) ** 2).mean()
        loss.backward()
        opt.step()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": sample * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} steps x {sample} trajectories (N={N}, T={T}, L={L}) fwd+bwd+Adam, torch CPU fp32, "
                      f"{cores} threads, {warmup} warm-up", "ms_per_step": 1e3 * dt / steps}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 2))
    cb = cpu_reference_arm(args, steps, warmup, args.cpu_sample)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, per_step=args.cpu_sample, note="CPU reference arm: bounded sample per step"),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_config(args, per_step=None, note=None):
    c = {"workload": f"EGNO charged N-body, {args.n_balls} particles, num_timesteps={args.timesteps}, hidden 64, "
                     f"{args.layers} layers (BASELINE.json configs[2])",
         "batch_per_gpu": per_step if per_step is not None else args.batch, "n_balls": args.n_balls,
         "num_timesteps": args.timesteps, "n_layers": args.layers, "parallelism": f"dp{args.gpus}", "launch": "eager" if args.no_graph else "cuda-graph replay (one graph per step)",
         "cache": "per-step activation working set (~270 MB at B=256) exceeds the 126 MB L2; inputs rotate over "
                  "8 distinct batches"}
    if note:
        c["note"] = note
    return c


# ------------------------------------------------------------------------------------------------ GPU arm
def finish(world, dist):
    """End of a rank: every rank reaches the barrier, then leaves without tearing NCCL down — destroying the process
    group while captured CUDA graphs still hold its collectives can block forever, and the process is exiting anyway."""
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_ours(args):
    import no_node_comparison_b200 as nb
    from no_node_comparison_b200 import synth
    from no_node_comparison_b200.dataparallel import init_from_env, broadcast_parameters
    import torch.distributed as dist
    import ctypes

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    rank, local, world = init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = nb.load_library()
    N, T, L, B = args.n_balls, args.timesteps, args.layers, args.batch
    K, W = args.steps, max(args.warmup, 3)

    torch.manual_seed(1)
    model = nb.EGNO(n_layers=L, in_node_nf=2, in_edge_nf=2, hidden_nf=64, with_v=True, num_modes=2, num_timesteps=T,
                    device=dev)
    if world > 1:
        broadcast_parameters(model)
        model.enable_data_parallel()
    use_graph = not args.no_graph
    opt = nb.FlatAdam(model.parameters(), lr=1e-4, weight_decay=1e-8)   # Adam of model_confs.yaml:15-17, one fused launch

    # ---- synthetic data: NB distinct batches per rank, raw states in pinned host memory
    NBATCH = 8
    row, col = synth.canonical_edges(B, N)
    row_d, col_d = row.to(dev), col.to(dev)
    edges = [row_d, col_d]
    t_out = torch.arange(1, T + 1, device=dev)[None].repeat(B, 1)
    host, resident = [], []
    for i in range(NBATCH):
        s = synth.sample_state("charged", B, N, seed=1000 * rank + i)
        tgt = (s["loc"].reshape(1, B * N, 3) + 0.05 * torch.randn(T, B * N, 3, generator=torch.Generator().manual_seed(i))
               ).reshape(T * B * N, 3)
        hb = {k: v.contiguous().pin_memory() for k, v in dict(loc=s["loc"], vel=s["vel"], charges=s["charges"], target=tgt).items()}
        host.append(hb)
        x, nodes, ea, v, lm = synth.egno_features(s["loc"].to(dev), s["vel"].to(dev), s["charges"].to(dev), row_d, col_d)
        resident.append(dict(x=x, nodes=nodes, ea=ea, v=v, lm=lm, target=tgt.to(dev)))
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0].values())

    def loss_fn(x, nodes, ea, v, lm, target):
        xo, vo, ho = model(x, nodes, edges, ea, v=v, loc_mean=lm, timesteps_out=t_out)
        return nb.trajectory_mse(xo, target, T)[0]      # the callers' MSE (main_simulation_simple_no.py:273-276), fused

    def loss_from_raw(loc, vel, charges, target):
        # prepare_inputs on the device: one featurisation kernel (nb_nbody_features)
        x, v = loc.reshape(-1, 3), vel.reshape(-1, 3)
        nodes, lm, ea = nb.prepare_inputs(x, v, charges, N, with_charge=True)
        return loss_fn(x, nodes, ea, v, lm, target)

    def eager_step(b):
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(b["x"], b["nodes"], b["ea"], b["v"], b["lm"], b["target"])
        loss.backward()
        opt.step()
        return loss

    if use_graph:
        # one CUDA graph per step shape: forward + loss + backward (+ all-reduce) + Adam replayed as a single launch
        g_res = nb.GraphedStep(loss_fn, {k: resident[0][k] for k in ("x", "nodes", "ea", "v", "lm", "target")}, opt)
        g_raw = nb.GraphedStep(loss_from_raw, {k: host[0][k].to(dev) for k in ("loc", "vel", "charges", "target")}, opt)

        def train_step(b):
            return g_res(**{k: b[k] for k in ("x", "nodes", "ea", "v", "lm", "target")})

        def e2e_step(hb):
            return g_raw(**hb).item()      # H2D from pinned memory into the graph's input buffers; D2H read of the loss
    else:
        train_step = eager_step

        def e2e_step(hb):
            d = {k: t.to(dev, non_blocking=True) for k, t in hb.items()}              # H2D from pinned memory
            loss = loss_from_raw(d["loc"], d["vel"], d["charges"], d["target"])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            return loss.item()                                                         # D2H read of the step's loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(W):
        train_step(resident[i % NBATCH])
    l0 = lib.nb_launch_count()
    tw0 = time.time()
    ms = timed(lambda i: train_step(resident[i % NBATCH]), K)
    tw1 = time.time()
    launches = lib.nb_launch_count() - l0
    if use_graph:   # replays do not pass through the library's host-side counter: kernels per captured step x K
        launches = g_res.launches_per_replay * K
    clocks = sampler.stop(window=(tw0, tw1)) if rank == 0 else None
    value = world * B * K / (ms / 1e3)

    if args.quick:
        lib.nb_profile_enable(1)
        ms_q = timed(lambda i: eager_step(resident[i % NBATCH]), K)
        qm, qc = (ctypes.c_double * 8)(), (ctypes.c_longlong * 8)()
        lib.nb_profile_read(qm, qc)
        lib.nb_profile_enable(0)
        per = {c: round(1e3 * qm[i] / max(qc[i], 1), 1) for i, c in enumerate(["edge_fwd", "edge_bwd", "gemm64", "wgrad64", "tconv"])}
        if rank == 0:
            emit(dict({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                              "ms_per_step": ms / K, "gpu_launches": int(launches), "quick": True,
                              "us_per_launch": per, "ms_in_kernels_per_step": round(sum(qm[:5]) / K, 3)}))
        finish(world, dist)
        return

    # ---- end to end: host buffers in, loss out, through the module API
    for i in range(2):
        e2e_step(host[i % NBATCH])
    ms_e2e = timed(lambda i: e2e_step(host[i % NBATCH]), K)
    e2e_value = world * B * K / (ms_e2e / 1e3)

    # ---- per-kernel CUDA-event timing (separate pass so the events do not perturb `value`)
    # (eager launches: the per-kernel events are recorded by the library's host-side launch code, which a graph replay skips)
    lib.nb_profile_enable(1)
    ms_prof = timed(lambda i: eager_step(resident[i % NBATCH]), K)
    pm = (ctypes.c_double * 8)()
    pc = (ctypes.c_longlong * 8)()
    lib.nb_profile_read(pm, pc)
    lib.nb_profile_enable(0)
    cats = ["edge_fwd", "edge_bwd", "gemm64", "wgrad64", "tconv"]
    kern = {c: {"ms_total": pm[i], "launches": int(pc[i]), "ms_per_launch": (pm[i] / pc[i]) if pc[i] else None,
                "share_of_step": pm[i] / ms_prof if ms_prof else None} for i, c in enumerate(cats)}

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        pk = json.load(open(peaks_path))
        peak_tf, peak_src = float(pk["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
        hbm = float(pk["hbm_gbs"])
    else:
        peak_tf, peak_src, hbm = 1400.0, "fallback (B200_PROFILING.md sustained bf16)", 6650.0
    ne = T * B * N * (N - 1)
    t_bwd = kern["edge_bwd"]["ms_per_launch"]
    flop_kernel = 2 * 2 * MAC_EDGE_KERNEL * ne           # backward = 2 x forward; recompute is not credited
    achieved = flop_kernel / (t_bwd * 1e-3) / 1e12 if t_bwd else None
    roofline = {"kernel": "k_edge_bwd_sel (fused E_GCL edge-tile backward: recompute + dgrad + wgrad + node gathers/scatters, "
                          "all on tcgen05 with split-bf16 operands and fp32 TMEM accumulators)",
                "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": (achieved / peak_tf) if achieved else None, "peak_source": peak_src,
                # dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full
                # (profiles/r01b_edge_bwd_sel_digest.txt); valid for the default workload only
                "traffic": 46.1e6 if (N, T, B) == (20, 10, 256) else None,
                "flop_per_launch": flop_kernel, "flop_per_launch_reference_formula": 2 * 2 * MAC_EDGE_REF * ne,
                "tensor_pipe_active_pct_ncu": 28.0 if (N, T, B) == (20, 10, 256) else None,
                "note": "achieved = ALGORITHMIC fp32 FLOPs the edge kernel owns (8448 MAC/edge forward, x2 for backward; the "
                        "recompute is not credited) / its mean launch time (CUDA events on the launch stream).  The reference's "
                        "dense 131-wide first layer would count 16640 MAC/edge (second figure).  Every logical fp32 MMA is "
                        "three bf16 tcgen05 passes (hi*hi + lo*hi + hi*lo) and the backward executes ~2x the credited MACs "
                        "(recompute, weight-gradient and one-hot scatter MMAs), so the tensor pipe is ~8x busier than `frac` "
                        "suggests: ncu reports 28 % tensor-pipe active for this launch.  The kernel is bound by its dependent "
                        "MMA -> TMEM -> SiLU -> smem -> MMA chain (4 round trips per 128-edge tile; per-stage cycles in "
                        "DESIGN.md / tools/stage_clocks.py), not by HBM (46 MB per launch = 105 GB/s).",
                "hbm_peak_gbs": hbm}

    # HBM-bound side of the path: the fused temporal convolution (forward reads h and writes h once; the backward reads h
    # and dL/dh_out, writes dL/dh_in and the 2 x 3 coefficient planes the weight-gradient reduction consumes)
    nn0 = B * N
    tconv_bytes_fwd = 2 * T * nn0 * 64 * 4
    tconv_bytes_bwd = 3 * T * nn0 * 64 * 4 + 2 * 3 * nn0 * 64 * 4
    t_tc = kern["tconv"]["ms_per_launch"]
    tc_ach = 0.5 * (tconv_bytes_fwd + tconv_bytes_bwd) / (t_tc * 1e-3) / 1e9 if t_tc else None
    roofline_hbm = {"kernel": "k_tconv_fwd / k_tconv_bwd (fused DFT + fp32 mode mixing + inverse DFT + LeakyReLU/residual)",
                    "bound": "hbm", "achieved": tc_ach, "peak": hbm, "unit": "GB/s", "frac": (tc_ach / hbm) if tc_ach else None,
                    "bytes_per_launch_fwd": tconv_bytes_fwd, "bytes_per_launch_bwd": tconv_bytes_bwd,
                    "note": "mean over the forward and backward launches of a step (CUDA events); algorithmic bytes = "
                            "2*T*64*4 B per node-trajectory forward, (3*T + 6)*64*4 B backward"}

    extras = {}
    if not args.no_extras:
        # inference throughput (no_grad forward) of the same model
        with torch.no_grad():
            f = lambda i: model(resident[i % NBATCH]["x"], resident[i % NBATCH]["nodes"], edges, resident[i % NBATCH]["ea"],
                                v=resident[i % NBATCH]["v"], loc_mean=resident[i % NBATCH]["lm"], timesteps_out=t_out)
            for i in range(3):
                f(i)
            ms_inf = timed(f, K)
        extras["egno_infer_traj_per_s"] = world * B * K / (ms_inf / 1e3)
        # SEGNO, BASELINE.json configs[3] shape (N=20, T=10, B=256): train step and forward
        torch.manual_seed(1)
        seg = nb.SEGNO(in_node_nf=1, in_edge_nf=2, hidden_nf=64, device=dev, n_layers=8, recurrent=True)
        if world > 1:
            broadcast_parameters(seg)
            seg.enable_data_parallel()
        sopt = nb.FlatAdam(seg.parameters(), lr=5e-3, weight_decay=1e-12)
        sb = []
        for i in range(NBATCH):
            s = synth.sample_state("gravity", B, N, seed=77 + 1000 * rank + i)
            his, x, v, ea = synth.segno_features(s["loc"].to(dev), s["vel"].to(dev), s["charges"].to(dev), row_d, col_d)
            sb.append(dict(his=his, x=x, v=v, ea=ea, target=x + 0.05 * torch.randn_like(x)))

        def seg_loss(his, x, v, ea, target):
            xo, ho, vo = seg(his, x, edges, v, ea, T=T)
            return ((xo - target) ** 2).mean()

        if use_graph:
            g_seg = nb.GraphedStep(seg_loss, sb[0], sopt)

            def seg_step(i):
                g_seg(**sb[i % NBATCH])
        else:
            def seg_step(i):
                sopt.zero_grad(set_to_none=True)
                seg_loss(**sb[i % NBATCH]).backward()
                sopt.step()

        for i in range(3):
            seg_step(i)
        ms_seg = timed(seg_step, K)
        extras["segno_train_traj_per_s"] = world * B * K / (ms_seg / 1e3)
        with torch.no_grad():
            g = lambda i: seg(sb[i % NBATCH]["his"], sb[i % NBATCH]["x"], edges, sb[i % NBATCH]["v"], sb[i % NBATCH]["ea"], T=T)
            for i in range(3):
                g(i)
            ms_sinf = timed(g, K)
        extras["segno_infer_traj_per_s"] = world * B * K / (ms_sinf / 1e3)
        # long-horizon rollouts kept on the device (featurisation + model + energies; BASELINE.json configs[3]:
        # traj_len = 20 calls, main.py:42), one host read at the end
        TL = 20
        s0 = synth.sample_state("gravity", B, N, seed=4242 + rank)
        l0_, v0_, m0_ = (s0[k].reshape(B * N, -1).to(dev) for k in ("loc", "vel", "charges"))
        ro = lambda i: nb.segno_rollout(seg, l0_, v0_, m0_, edges, N, traj_len=TL, num_steps=T, dataset="gravity")[1].sum().item()
        ro(0)
        ms_ro = timed(ro, 3)
        extras["segno_rollout_traj_per_s"] = world * B * 3 / (ms_ro / 1e3)
        extras["segno_rollout_note"] = f"{TL} autoregressive calls x {T} sub-steps + energy of every frame per trajectory"
        c0 = synth.sample_state("charged", B, N, seed=777 + rank)
        cl, cv, cq = (c0[k].reshape(B * N, -1).to(dev) for k in ("loc", "vel", "charges"))
        er = lambda i: nb.egno_rollout(model, cl, cv, cq, edges, N, traj_len=TL, dataset="charged")[2].sum().item()
        er(0)
        ms_er = timed(er, 3)
        extras["egno_rollout_traj_per_s"] = world * B * 3 / (ms_er / 1e3)
        extras["egno_rollout_note"] = f"{TL} autoregressive calls x {T} frames + energy of every frame per trajectory"
        # the two 5-body configurations (BASELINE.json configs[0], configs[1]): launch-bound, so one CUDA graph per step
        if world == 1:
            extras.update(small_configs(nb, synth, dev, K))

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_arm(args, steps=8, warmup=1, sample=args.cpu_sample)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(args), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / K},
                "gpu_launches": int(launches), "roofline": roofline, "roofline_hbm": roofline_hbm,
                "cpu_baseline": cpu_baseline,
                "kernels": kern, "extras": extras}
        emit(line)
    finish(world, dist)


def small_configs(nb, synth, dev, K):
    """Training throughput of the 5-body configurations: EGNO N=5, T=8, L=4, B=100 and SEGNO N=5, T=10, B=100."""
    out = {}
    B, N = 100, 5
    row, col = synth.canonical_edges(B, N)
    edges = [row.to(dev), col.to(dev)]

    def run(step, n):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        return B * n / (e0.elapsed_time(e1) / 1e3)

    T = 8
    torch.manual_seed(1)
    m = nb.EGNO(n_layers=4, in_node_nf=2, in_edge_nf=2, hidden_nf=64, with_v=True, num_modes=2, num_timesteps=T, device=dev)
    opt = nb.FlatAdam(m.parameters(), lr=1e-4, weight_decay=1e-8)
    s = synth.sample_state("charged", B, N, seed=5)
    x, nodes, ea, v, lm = synth.egno_features(s["loc"].to(dev), s["vel"].to(dev), s["charges"].to(dev), edges[0], edges[1])
    tgt = x.repeat(T, 1) + 0.05 * torch.randn(T * B * N, 3, device=dev)
    t_out = torch.arange(1, T + 1, device=dev)[None].repeat(B, 1)

    def f1(x, nodes, ea, v, lm, tgt):
        xo, _, _ = m(x, nodes, edges, ea, v=v, loc_mean=lm, timesteps_out=t_out)
        return ((xo - tgt) ** 2).mean()

    ins = dict(x=x, nodes=nodes, ea=ea, v=v, lm=lm, tgt=tgt)
    g1 = nb.GraphedStep(f1, ins, opt)
    out["egno_n5_t8_b100_train_traj_per_s"] = run(lambda: g1(**ins), 4 * K)
    T2 = 10
    torch.manual_seed(1)
    sg = nb.SEGNO(in_node_nf=1, in_edge_nf=2, hidden_nf=64, device=dev, n_layers=8, recurrent=True)
    sopt = nb.FlatAdam(sg.parameters(), lr=5e-3, weight_decay=1e-12)
    his, x2, v2, ea2 = synth.segno_features(s["loc"].to(dev), s["vel"].to(dev), s["charges"].to(dev), edges[0], edges[1])
    tgt2 = x2 + 0.05 * torch.randn_like(x2)

    def f2(his, x, v, ea, tgt):
        xo, _, _ = sg(his, x, edges, v, ea, T=T2)
        return ((xo - tgt) ** 2).mean()

    ins2 = dict(his=his, x=x2, v=v2, ea=ea2, tgt=tgt2)
    g2 = nb.GraphedStep(f2, ins2, sopt)
    out["segno_n5_t10_b100_train_traj_per_s"] = run(lambda: g2(**ins2), 4 * K)
    out["small_configs_note"] = "BASELINE.json configs[0] / [1] (B=100): launch-bound; whole step replayed as one CUDA graph"
    del g1, g2, m, sg
    # BASELINE.json configs[4] shape: 100-body EGNO, the per-GPU share (64) of the batch of 512 at 8 GPUs
    B5, N5, T5 = 64, 100, 10
    row5, col5 = synth.canonical_edges(B5, N5)
    e5 = [row5.to(dev), col5.to(dev)]
    torch.manual_seed(1)
    m5 = nb.EGNO(n_layers=4, in_node_nf=2, in_edge_nf=2, hidden_nf=64, with_v=True, num_modes=2, num_timesteps=T5, device=dev)
    o5 = nb.FlatAdam(m5.parameters(), lr=1e-4, weight_decay=1e-8)
    s5 = synth.sample_state("charged", B5, N5, seed=9)
    x5, n5, ea5, v5, lm5 = synth.egno_features(s5["loc"].to(dev), s5["vel"].to(dev), s5["charges"].to(dev), e5[0], e5[1])
    tg5 = x5.repeat(T5, 1) + 0.05 * torch.randn(T5 * B5 * N5, 3, device=dev)
    to5 = torch.arange(1, T5 + 1, device=dev)[None].repeat(B5, 1)

    def step5():
        o5.zero_grad(set_to_none=True)
        xo, _, _ = m5(x5, n5, e5, ea5, v=v5, loc_mean=lm5, timesteps_out=to5)
        ((xo - tg5) ** 2).mean().backward()
        o5.step()

    B = B5
    out["egno_n100_b64_train_traj_per_s"] = run(step5, 5)
    out["egno_n100_note"] = "BASELINE.json configs[4] shape per GPU (B=512 over 8 GPUs): N=100, T=10, L=4, 6.3M edges per layer"
    return out


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: libraries that print to file descriptor 1 (NCCL's "NCCL version ..." banner when
    # NCCL_DEBUG is set in the environment) are sent to stderr; emit() writes the line to the real stdout
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
