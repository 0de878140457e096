#!/usr/bin/env python
"""Benchmark of the OR-NOR Gibbs hot path: edge-updates/sec on the STRING-scale network.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one rank per GPU)
  python bench.py --impl reference --steps K --warmup W    the reference's own C++ on the host cores

Workload (BASELINE.json configs[2]): synthetic network following
gbnet.tools.genData's law, 1,500 TFs / 20,000 targets / ~300,000 signed edges, CLI-default
hyper-parameters, 256 chains per GPU (weak scaling: N GPUs run 256 N chains of ONE model, chains
sharded over ranks, Gelman-Rubin moments all-reduced with NCCL).  One step = what one iteration of
ModelBase::sample does (ModelBase.cpp:119-124): a batch of dN = 30 full sweeps X -> T -> S of every
chain, followed by the max Gelman-Rubin evaluation.

Printed JSON (one line, rank 0): see the task contract; `value` is device-timed (CUDA events on the
model's stream, max over ranks) with everything resident in HBM; `e2e` goes through the public API
with host buffers: per step new evidence from host memory -> chains re-initialised -> dN sweeps ->
max R-hat and posterior X means read back to the host.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "edge-updates/sec (all chains, device-timed)"
UNIT = "edge-updates/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--shape", default="C3", choices=["C1", "C2", "C3"])
    ap.add_argument("--chains", type=int, default=256, help="chains per GPU")
    ap.add_argument("--sweeps-per-step", type=int, default=30)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--roofline-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the other BASELINE configs (C5, C4, time to R-hat < 1.1) and the sharding self-check")
    ap.add_argument("--ref-sweeps-per-step", type=int, default=1)
    return ap.parse_args()


def workload(shape):
    from gbnet_b200 import synth
    net = synth.shape(shape)
    hy = synth.cli_default_hyper()
    return net, hy


def config_dict(args, net, world, chains_per_gpu):
    r = net.realised()
    return {"workload": f"{args.shape}: synthetic genData-law network {r['n_tf']} TFs / {r['n_trg']} targets / "
                        f"{r['n_edge']} edges, {chains_per_gpu} chains per GPU, CLI-default hyper-parameters",
            "n_tf": r["n_tf"], "n_targets": r["n_trg"], "n_edges": r["n_edge"],
            "chains_per_gpu": chains_per_gpu, "chains_total": chains_per_gpu * world,
            "sweeps_per_step": args.sweeps_per_step, "step": "dN sweeps of every chain + max Gelman-Rubin",
            "parallelism": f"chains sharded over {world} GPU(s); NCCL all-reduce of R-hat moments",
            "l2": "per-chain state + counters are larger than L2 (no flush needed)"}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU arms
def time_cpu(net, hy, n_graphs, sweeps, warm_sweeps, prefer_ref=True):
    """edge-updates/s of the CPU sampler (the compiled reference when oracle/_ref exists, else
    the C port) with one chain per host thread; construction is not timed."""
    import oracle
    cores = os.cpu_count() or 1
    pb = oracle.Problem(src=net.src, trg=net.trg, mor=net.mor, evid_trg=net.evid_trg, evid_val=net.evid_val,
                        n_graphs=n_graphs, **hy)
    kind = "reference" if (prefer_ref and oracle.have_ref()) else "port"
    M = oracle.RefModel(pb) if kind == "reference" else oracle.PortModel(pb)
    M._f("omp_set_threads")(min(cores, n_graphs))
    if warm_sweeps:
        M.sample_n(warm_sweeps)
    times = [M.time_sample_n(s) for s in sweeps]
    M.close()
    return kind, min(cores, n_graphs), times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    net, hy = workload(args.shape)
    cores = os.cpu_count() or 1
    n_graphs = min(args.chains, cores)
    s = args.ref_sweeps_per_step
    kind, used, times = time_cpu(net, hy, n_graphs, [s] * args.steps, args.warmup * s)
    total = sum(times)
    value = n_graphs * s * args.steps * net.n_edges / total
    cfg = config_dict(args, net, 1, args.chains)
    cfg["sweeps_per_step"] = s
    sample = (f"{n_graphs} of the workload's {args.chains} chains (one per host thread), {s} sweep(s) per step, "
              f"ModelBase::sample_n timed with omp_get_wtime, construction excluded")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------ native arm
class _StdoutToStderr:
    """NCCL prints its version banner on stdout when a communicator is created; the contract wants
    exactly one JSON line there, so file descriptor 1 points at stderr while communicators are set up."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)



# ------------------------------------------------------------------------------------------ secondary configs
def _attach(G, dist, rank, world):
    """library-side NCCL communicator for the ranks of the torch process group (the id travels through torch)"""
    if world == 1:
        return
    import gbnet_b200
    with _StdoutToStderr():
        uid = [gbnet_b200.Sampler.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        G.attach_comm(rank, world, uid[0])
        G.max_gelman_rubin()


def _timed_steps(G, dN, steps, warmup, barrier):
    """device ms of `steps` x (dN sweeps + max R-hat) after `warmup` untimed ones"""
    for _ in range(warmup):
        G.sample_n(dN)
        G.max_gelman_rubin()
    barrier(G)
    G.timer_start()
    gr = None
    for _ in range(steps):
        G.sample_n(dN)
        gr = G.max_gelman_rubin()
    ms = G.timer_stop()
    barrier(G)
    return ms, gr


def run_secondary(args, dist, rank, local_rank, world, reduce_max):
    """The other BASELINE.json configs on the same GPUs, and the sharding self-check:
    C5  4,096 chains of the C3 network in total, strong-scaled over the ranks (chains sharded, NCCL moments)
    C4  C2 network, 512 contrasts x 8 chains, contrasts sharded over the ranks (independent models: no collective)
    time to R-hat < 1.1 (protocol of gbnet/utils.py:93-135) for C1 and C2, CUDA against the reference C++ (N = 1)
    parity_check (N > 1): a small model sharded over the ranks gives the R-hat / posterior means of the same
    model held by one rank."""
    import gbnet_b200
    from gbnet_b200 import synth
    from gbnet_b200.distributed import shard_chains
    hy = synth.cli_default_hyper()
    dN = args.sweeps_per_step
    out = {}

    def barrier(G):
        if world > 1:
            import torch
            dist.barrier()
            torch.cuda.synchronize()
        G.synchronize()

    # ---- C5
    net = synth.shape("C3")
    off, loc = shard_chains(4096, rank, world)
    G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, net.evid_val, n_graphs=4096, chain_offset=off,
                           n_graphs_local=loc, seed=2026, device=local_rank, **hy)
    _attach(G, dist, rank, world)
    steps = 3
    ms, gr = _timed_steps(G, dN, steps, 1, barrier)
    ms = reduce_max(ms)
    out["C5"] = {"workload": "C3 network, 4,096 chains in total (strong scaling), NCCL moment all-reduce every dN sweeps",
                 "chains_total": 4096, "chains_per_gpu": loc, "value": 4096 * dN * steps * net.n_edges / (ms * 1e-3),
                 "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "max_gelman_rubin": gr}
    G.close()

    # ---- C4
    net2 = synth.shape("C2")
    n_ds = 512
    d0, dl = shard_chains(n_ds, rank, world)                         # contrasts are independent models: shard them
    evs = np.stack([synth.redraw_evidence(net2, num_active_tfs=8, seed=100 + d0 + i)[0] for i in range(dl)])
    G = gbnet_b200.Sampler(net2.src, net2.trg, net2.mor, net2.evid_trg, evs, n_graphs=8, seed=13, device=local_rank, **hy)
    steps = 10
    ms, gr = _timed_steps(G, dN, steps, 2, barrier)
    ms = reduce_max(ms)
    out["C4"] = {"workload": "C2 network, 512 contrasts x 8 chains, contrasts sharded over the GPUs (no collective)",
                 "contrasts_per_gpu": dl, "chains_per_gpu": 8 * dl, "path": G.path(),
                 "value": n_ds * 8 * dN * steps * net2.n_edges / (ms * 1e-3), "unit": UNIT,
                 "ms_per_step": ms / steps, "steps": steps, "max_gelman_rubin_rank0": gr}
    G.close()

    # ---- time to R-hat < 1.1 (rank 0; the reference leg only at N = 1, as cpu_baseline)
    if rank == 0:
        ttr = {}
        for shape, chains in (("C1", 3), ("C2", 64)):
            netp = synth.shape(shape)
            G = gbnet_b200.Sampler(netp.src, netp.trg, netp.mor, netp.evid_trg, netp.evid_val, n_graphs=chains, seed=12,
                                   device=local_rank, **hy)
            G.run_protocol(burn_its=1, max_its=2)                   # warm the path (kernel load, first launch)
            G.close()
            G = gbnet_b200.Sampler(netp.src, netp.trg, netp.mor, netp.evid_trg, netp.evid_val, n_graphs=chains, seed=12,
                                   device=local_rank, **hy)
            t0 = time.perf_counter()
            its, mgr = G.run_protocol(burn_its=10, max_its=200)
            tg = time.perf_counter() - t0
            ttr[shape] = {"chains": chains, "edges": netp.n_edges, "gpu_s": tg, "gpu_iterations": its, "gpu_max_rhat": mgr,
                          "path": G.path()}
            G.close()
            if world == 1 and not args.no_cpu_baseline:
                import oracle
                pb = oracle.Problem(src=netp.src, trg=netp.trg, mor=netp.mor, evid_trg=netp.evid_trg,
                                    evid_val=netp.evid_val, n_graphs=chains, **hy)
                R = oracle.RefModel(pb) if oracle.have_ref() else oracle.PortModel(pb)
                R._f("omp_set_threads")(min(os.cpu_count() or 1, chains))
                t0 = time.perf_counter()
                it = 0
                for _ in range(10):
                    it += 1
                    R.sample(20, 30)
                R.burn_stats()
                mgr = float("inf")
                while mgr > 1.1 and it < 200:
                    it += 1
                    R.sample(20, 30)
                    mgr = R.max_gelman_rubin()
                ttr[shape].update({"reference_s": time.perf_counter() - t0, "reference_iterations": it,
                                   "reference_max_rhat": mgr, "reference_cores": min(os.cpu_count() or 1, chains),
                                   "reference_kind": "reference" if oracle.have_ref() else "port"})
                R.close()
        out["time_to_rhat"] = {"protocol": "gbnet/utils.py:93-135: 10 x sample(20), burn_stats, sample(20) until max R-hat <= 1.1",
                               **ttr}

    # ---- sharding self-check
    if world > 1:
        nets = synth.gen_network(NX=40, NY=800, AvgNTF=3.0, num_active_tfs=4, seed=17)
        M = 4 * world
        off, loc = shard_chains(M, rank, world)
        G = gbnet_b200.Sampler(nets.src, nets.trg, nets.mor, nets.evid_trg, nets.evid_val, n_graphs=M, chain_offset=off,
                               n_graphs_local=loc, seed=7, device=local_rank, **hy)
        _attach(G, dist, rank, world)
        G.sample_n(60)
        gr, px, ps, pt = G.max_gelman_rubin(), G.posterior_means("X"), G.posterior_means("S"), G.posterior_means("T")
        grn = np.asarray(G.gelman_rubin())
        G.close()
        if rank == 0:
            F = gbnet_b200.Sampler(nets.src, nets.trg, nets.mor, nets.evid_trg, nets.evid_val, n_graphs=M, seed=7,
                                   device=local_rank, **hy)
            F.sample_n(60)
            fgr, fgrn = F.max_gelman_rubin(), np.asarray(F.gelman_rubin())
            nXs = F.n_tf()
            disc = np.r_[0:nXs, 2 * nXs:len(fgrn)]           # X and S nodes (random_nodes order: X, T, S): integer count sums
            detail = {"max_rhat_sharded": gr, "max_rhat_single": fgr,
                      "rhat_discrete_nodes_identical": bool(np.array_equal(fgrn[disc], grn[disc])),
                      "rhat_t_nodes_max_rel_diff": float(np.max(np.abs(fgrn[nXs:2 * nXs] - grn[nXs:2 * nXs]) / fgrn[nXs:2 * nXs])),
                      "x_means_identical": bool(np.array_equal(F.posterior_means("X"), px)),
                      "s_means_identical": bool(np.array_equal(F.posterior_means("S"), ps)),
                      "t_means_max_abs_diff": float(np.max(np.abs(F.posterior_means("T") - pt)))}
            F.close()
            # X / S statistics cross the ranks as exact integer sums: identical whatever the rank count; the T nodes'
            # fp64 sums are reduced in a different order: 1e-12
            ok = bool(detail["rhat_discrete_nodes_identical"] and detail["x_means_identical"] and detail["s_means_identical"]
                      and detail["rhat_t_nodes_max_rel_diff"] < 1e-12 and detail["t_means_max_abs_diff"] < 1e-12
                      and abs(fgr - gr) <= 1e-12 * fgr)
            out["parity_check"] = ok
            out["parity_check_detail"] = detail
            out["parity_check_what"] = (f"{M} chains of a 40-TF network sharded over {world} ranks vs the same model on one "
                                        "rank after 60 sweeps: R-hat of every X / S node and posterior X / S means identical, "
                                        "R-hat and means of the T nodes to 1e-12")
    return out


def run_native(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        with _StdoutToStderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()

    import gbnet_b200
    from gbnet_b200 import synth
    if gbnet_b200.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device (gbnet_b200 has no CPU path)")
    net, hy = workload(args.shape)
    E = net.n_edges
    C = args.chains
    dN = args.sweeps_per_step

    G = gbnet_b200.Sampler(net.src, net.trg, net.mor, net.evid_trg, net.evid_val, n_graphs=C * world,
                           chain_offset=rank * C, n_graphs_local=C, seed=2026, device=local_rank, **hy)
    assert G.kernel() == gbnet_b200.KERNEL_FAST
    if world > 1:
        import torch
        with _StdoutToStderr():
            uid = [gbnet_b200.Sampler.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            G.attach_comm(rank, world, uid[0])
            G.max_gelman_rubin()                   # first all-reduce (lazy NCCL connection set-up)

    def barrier():
        if world > 1:
            import torch
            dist.barrier()
            torch.cuda.synchronize()
        G.synchronize()

    def step():
        G.sample_n(dN)
        return G.max_gelman_rubin()

    for _ in range(args.warmup):
        step()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = G.launch_count()
    barrier()
    t0 = time.perf_counter()
    G.timer_start()
    gr = None
    for _ in range(args.steps):
        gr = step()
    dev_ms = G.timer_stop()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = G.launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None

    # ---- per-kernel device times: the same steps with the kernels serialised on one stream (in the timed
    # region above, chain groups on separate streams overlap the X and S kernels, so a kernel's own
    # duration is only defined here); CUDA events on the model's stream around every launch
    G.set_stream_groups(1)
    G.set_phase_timing(True)
    G.timer_start()
    for _ in range(args.roofline_steps):
        step()
    serial_ms = G.timer_stop()
    phase_ms, phase_launches = G.phase_ms()
    G.set_phase_timing(False)
    G.set_stream_groups(0)

    # ---- end to end through the public API with host buffers
    evs = [synth.redraw_evidence(net, num_active_tfs=12, seed=500 + i)[0] for i in range(args.e2e_steps + 1)]
    G.set_evidence(net.evid_trg, evs[-1])          # warm the path once
    G.sample_n(dN)
    barrier()
    t1 = time.perf_counter()
    post = None
    for i in range(args.e2e_steps):
        G.set_evidence(net.evid_trg, evs[i])       # host -> device: evidence; chains re-initialised
        G.sample_n(dN)
        gr_e2e = G.max_gelman_rubin()              # device -> host
        post = G.posterior_means("X")              # device -> host
    barrier()
    e2e_s = time.perf_counter() - t1
    h2d = int(net.evid_trg.nbytes + evs[0].nbytes)
    d2h = int(post.nbytes + 8)

    G.close()
    secondary = None
    if not args.no_secondary:
        def reduce_max(v):
            if world == 1:
                return v
            import torch
            t = torch.tensor([v], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0])
        secondary = run_secondary(args, dist, rank, local_rank, world, reduce_max)

    if world > 1:
        import torch
        t = torch.tensor([dev_ms, wall_ms, e2e_s, float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms, e2e_s = float(t[0]), float(t[1]), float(t[2])
        tl = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
        launches = int(tl[0])

    if rank == 0:
        r_targets = net.realised()["n_trg"]
        total_updates = world * C * dN * args.steps * E
        value = total_updates / (dev_ms * 1e-3)
        e2e_value = world * C * dN * args.e2e_steps * E / e2e_s

        # roofline of the dominant kernel (algorithmic bytes: DESIGN.md §5)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # the kernels behind the three phases of a sweep at this shape (sweep_fast.cu; GBN_X1=2 / GBN_S_V=1 select the
        # chain-pair X kernel / the first form of the S kernel)
        names = ["k_fast_x" if os.environ.get("GBN_X1") == "2" else "k_fast_x1", "k_fast_t",
                 "k_fast_s" if os.environ.get("GBN_S_V") == "1" else "k_fast_s2"]
        # algorithmic bytes per edge-update for C chains sharing one network:
        #   S kernel: S state 2 x 0.25 B + two u32 counters read+write 16 B + by-target parent id / prior row 4.25 B / C
        #   X kernel: by-TF target id + sign 4.25 B / C + per-TF terms amortised over the out-degree 0.2 B
        #   whole sweep: B_edge = 16.7 + 16.5 / C
        per_unit = [0.2 + 8.25 / C, 0.0, 16.5 + 8.25 / C]
        b_edge = 16.7 + 16.5 / C
        dom = int(np.argmax(phase_ms))
        avg_ms = phase_ms[dom] / max(phase_launches, 1)
        bytes_per_launch = per_unit[dom] * C * E                    # serialised: one launch covers all C chains
        achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        traffic = None
        for tname in ("r2_traffic.json", "traffic.json"):          # per launch, from the latest `ncu --set full` capture
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath):
                traffic = json.load(open(tpath)).get(names[dom])
                break
        sweep_gbs = value / world * b_edge / 1e9          # per GPU: the peak is one GPU's
        # compulsory bytes of the round-2 layout per edge-update of the S kernel: byte counters
        # 2 B read + 2 B written, S code 0.25 B read, activity bit 0.125 B, per-target outputs (flip weights 16 B,
        # likelihood 8 B, bound 1 B) amortised over the target's edges, network words shared by the C chains
        comp_edge = 4.0 + 0.25 + 0.125 + 25.0 * r_targets / E + 4.6 / C
        comp_bytes = comp_edge * C * E
        roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms,
                    "bytes_per_unit": per_unit[dom],
                    "what": "SURVEY.md section 8(d) bytes per edge-update (u32 counter pairs read + written every sweep): "
                            "the figure round 1 was reported against",
                    "compulsory": {"bytes_per_unit": comp_edge, "bytes_per_launch": comp_bytes,
                                   "achieved": comp_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0,
                                   "frac": comp_bytes / (avg_ms * 1e-3) / 1e9 / peak if avg_ms > 0 else 0.0,
                                   "what": "bytes the round-2 layout has to move (8-bit per-batch counters, 2-bit S codes); "
                                           "the kernel is bound by instruction issue, not by HBM (profiles/r2_summary.md)"},
                    "how": "CUDA events on the model's stream around every launch, kernels serialised "
                           f"(stream groups = 1) for {args.roofline_steps} steps right after the timed region",
                    "kernel_ms_per_sweep": {n: phase_ms[i] / max(phase_launches, 1) for i, n in enumerate(names)},
                    "kernel_share_of_step": {n: phase_ms[i] / serial_ms for i, n in enumerate(names)},
                    "serialised_ms_per_step": serial_ms / args.roofline_steps,
                    "whole_sweep": {"achieved": sweep_gbs, "frac": sweep_gbs / peak, "bytes_per_edge_update": b_edge,
                                    "how": "value / n_gpus x B_edge in the timed region (kernels of the chain groups overlapped)"}}

        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(args, net, world, C),
                "wall_ms_per_step": wall_ms / args.steps, "max_gelman_rubin": gr,
                "clocks": clk,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": args.e2e_steps,
                        "what": "set_evidence(host arrays) + sample_n(dN) + max R-hat + posterior X means to host"},
                "gpu_launches": launches,
                "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_graphs = min(C, cores)
            kind, used, times = time_cpu(net, hy, n_graphs, [1, 1, 1], 1)
            line["cpu_baseline"] = {
                "value": n_graphs * 3 * E / sum(times), "unit": UNIT, "cores": used, "kind": kind,
                "sample": f"{n_graphs} of {C} chains (one per host thread), 3 sweeps after 1 warm-up sweep, "
                          "ModelBase::sample_n timed with omp_get_wtime, construction excluded"}
        if secondary is not None:
            line["secondary"] = secondary
            if "parity_check" in secondary:
                line["parity_check"] = secondary["parity_check"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
