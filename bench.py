#!/usr/bin/env python
"""Headline benchmark: system-steps/s of the fused forward+gradient RC rollout (BASELINE.json).

    python bench.py [--gpus N --steps K --warmup W]            # our arm (CUDA library)
    python bench.py --impl reference [--steps K --warmup W]    # reference arm: CPU port of the reference path

A "step" = one loss+gradient evaluation (examples/3-inverse.py:96-112) of `--systems` RC zone models
over T=8760 hourly steps on EACH rank (weak scaling: the system batch is sharded, no collective on the
data path).  The full input array of the headline config (2^20 x 8760 x 5 fp32 = 184 GB) does not fit
one GPU, so a step runs as tiles over the system axis: one u/target tile stays resident and is
re-read by every tile launch (24 GB >> 126 MB L2, so nothing is served from cache) while theta/x0
are distinct per tile.  See DESIGN.md "Measurement".

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_YEAR = 8760
DT = 3600.0
# The library computes the loss+gradient with one of two kernels:
#   tangent  forward-mode sensitivities in ONE sweep: u (20 B) + target (4 B) per system-step  -> 24 B
#   adjoint  checkpointed discrete adjoint, forward sweep + reverse sweep -> 48 B
# `alg_bytes` is the algorithmic HBM traffic of the kernel actually launched, so roofline.frac is
# comparable between the two; `value` (system-steps/s) is what the user sees.
ALGOS = {
    "tangent": {"alg_bytes": 24, "kernel": "rc_fwdsens_kernel<TN=256,KF=8,S=2,2 CTAs/SM,SPLIT=1> (packed fp32, TMA tensor maps)", "env": "1"},
    "adjoint": {"alg_bytes": 48, "kernel": "rollout_adj_kernel<RCModel,128,12,2,MSE>", "env": "0"},
}
WAVE = 148 * 2 * 256  # systems per full wave: 2 CTAs/SM x 256 (tangent) == 148 x 4 x 128; the adjoint's is 148*3*128
NOMINAL = [10384.31640625, 499089.09375, 1321535.125, 1.5348844528198242, 0.5000327825546265,
           1.000040054321289, 20.119935989379883]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--systems", type=int, default=1 << 20, help="systems per rank per step")
    ap.add_argument("--T", type=int, default=T_YEAR)
    ap.add_argument("--algo", default="tangent", choices=sorted(ALGOS),
                    help="gradient kernel: the library default (tangent) or the checkpointed adjoint")
    ap.add_argument("--tile", type=int, default=None, help="systems per kernel launch (default: 3 full waves)")
    ap.add_argument("--e2e-systems", type=int, default=32768)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-systems", type=int, default=8192, help="bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity sample against the fp64 oracle")
    ap.add_argument("--parity-systems", type=int, default=4096, help="first K + K random systems of the timed run")
    ap.add_argument("--no-config3", action="store_true", help="skip the BASELINE.json config-3 leg (strong scaling + all-reduce)")
    ap.add_argument("--config3-systems", type=int, default=65536, help="systems in TOTAL over all ranks (strong scaling)")
    ap.add_argument("--config3-steps", type=int, default=20)
    ap.add_argument("--no-other-configs", action="store_true", help="skip the legs for BASELINE.json configs[1], [3], [4]")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region through NVML
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if "-" in part:
            lo, hi = part.split("-")
            cpus.update(range(int(lo), int(hi) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs of its GPU's NUMA node, so that first-touch places the pinned host buffers of the
    e2e leg next to the GPU (8 ranks otherwise share one node's memory).  The node comes from the PCI device in sysfs
    (NVML's affinity mask was flat on the round-1 box); NVML is the fallback."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node >= 0 and len(nodes) > 1:
            cpus = _cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()) & os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                return f"gpu {bdf} on numa node {node} of {len(nodes)} (sysfs): bound to its {len(cpus)} cpus"
        sys_note = f"sysfs: gpu {bdf} numa_node={node}, {len(nodes)} node(s)"
    except Exception as e:
        sys_note = f"sysfs probe failed ({type(e).__name__})"
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{sys_note}; bound to {len(cpus)} GPU-local cpus (NVML affinity)"
    except Exception as e:  # not fatal: only a placement hint
        return f"{sys_note}; no numa binding ({type(e).__name__})"
    return f"{sys_note}; no numa binding (empty intersection)"


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic_per_step(algo):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (or None)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[algo]
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle's C port of the reference arithmetic on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_sample(n_sys, T, seed=99):
    import numpy as np
    rng = np.random.default_rng(seed)
    th = (np.array(NOMINAL)[None] * np.exp(rng.uniform(-0.2, 0.2, size=(n_sys, 7)))).astype(np.float32)
    x0 = (np.array([20.0, 30.0, 26.0])[None] + rng.uniform(-1, 1, size=(n_sys, 3))).astype(np.float32)
    u = rng.random((T, n_sys, 5), dtype=np.float32)
    u[..., 0] = 15 + 10 * u[..., 0]
    target = (20 + 2 * rng.random((T, n_sys), dtype=np.float32)).astype(np.float32)
    return th, x0, u, target


def time_cpu(n_sys, T, steps, warmup):
    """All-fp32 BPTT of the reference path (oracle/dynax_oracle.c, OpenMP over all host threads)."""
    from oracle import c_oracle
    th, x0, u, target = cpu_sample(n_sys, T)
    c_oracle.lib(native=True)
    c_oracle.use_all_cores()
    cores = c_oracle.num_threads()
    for _ in range(warmup):
        c_oracle.rc_loss_grad(th, x0, u, target, DT, native=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        c_oracle.rc_loss_grad(th, x0, u, target, DT, native=True)
    dt = time.perf_counter() - t0
    return n_sys * T * steps / dt, cores, dt / steps


def time_cpu_numba(n_sys, T):
    """Second, independently written CPU implementation (oracle/numba_oracle.py) on the same
    sample: the C port's figure is only trusted as the baseline if the two agree within a small factor."""
    try:
        from oracle import numba_oracle
    except ImportError as e:
        return {"unavailable": f"{type(e).__name__}: {e}"}
    tiny = cpu_sample(64, 16)
    cores = numba_oracle.use_all_cores()
    numba_oracle.rc_loss_grad(*tiny, DT)           # JIT / cache load outside the timed region
    th, x0, u, target = cpu_sample(n_sys, T)
    numba_oracle.rc_loss_grad(th, x0, u, target, DT)
    t0 = time.perf_counter()
    numba_oracle.rc_loss_grad(th, x0, u, target, DT)
    dt = time.perf_counter() - t0
    return {"value": n_sys * T / dt, "unit": "system-steps/s", "cores": cores,
            "impl": "numba @njit(parallel=True), all-fp32 BPTT", "sample": f"{n_sys} systems x {T} steps"}


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, cores, sec = time_cpu(a.cpu_systems, a.T, a.steps, a.warmup)
    numba_check = time_cpu_numba(a.cpu_systems, a.T)
    sample = f"{a.cpu_systems} systems x {a.T} steps per step (bounded sample of the {a.systems}-system workload)"
    line = {
        "impl": "reference", "metric": "system-steps/s, fwd+grad RC rollout", "value": val, "unit": "system-steps/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"rc_fwd_grad N={a.systems} T={a.T} dt=3600 per-system theta/u/target (per rank)", "sample": sample},
        "cpu_baseline": {"value": val, "unit": "system-steps/s", "cores": cores, "kind": "port", "sample": sample,
                         "numba_cross_check": numba_check},
        "e2e": {"value": val, "unit": "system-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is JAX/Flax (not installable here: no jax wheel, no network); this arm times the oracle's "
                "C/OpenMP port of the same arithmetic (all-fp32 BPTT storing per-step residuals like XLA's scan transpose)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# in-run parity sample (SURVEY.md section 8(d) "Parity sampling"): first K + K random systems of the TIMED run
# against the fp64 oracle, outside the timed region
# ------------------------------------------------------------------------------------------------
def parity_block(torch, theta, x0, u, target, loss, grad, tile, K, T):
    """The timed step evaluates system s with theta[s], x0[s] and column (s mod tile) of the resident u/target tile.
    Returns the `parity` object of the JSON line."""
    import numpy as np
    from oracle import c_oracle
    S = theta.shape[0]
    K = min(K, S // 2)
    rng = np.random.default_rng(4242)
    idx = np.concatenate([np.arange(K), np.sort(rng.choice(np.arange(K, S), K, replace=False))])
    col = torch.as_tensor(idx % tile, device=theta.device)
    sel = torch.as_tensor(idx, device=theta.device)
    th, x0s = theta[sel].cpu().numpy(), x0[sel].cpu().numpy()
    us, tgs = u[:, col].cpu().numpy(), target[:, col].cpu().numpy()
    c_oracle.lib(native=True)
    c_oracle.use_all_cores()
    t0 = time.perf_counter()
    l64, g64, lbud, gbud = c_oracle.rc_loss_grad_budget_f64(th, x0s, us, tgs, DT, native=True)
    _, g32 = c_oracle.rc_loss_grad(th, x0s, us, tgs, DT, native=True)      # the all-fp32 reference-order BPTT (yardstick)
    oracle_s = time.perf_counter() - t0
    lg, gg = loss[sel].double().cpu().numpy(), grad[sel].double().cpu().numpy()
    loss_frac = np.abs(lg - l64) / (1e-5 * np.abs(l64) + lbud)
    # per element: rel 1e-4 + the conditioning budget (tests/util.py), or at least as close to fp64 truth as the fp32
    # reference-order BPTT is; per system: norm-wise 1e-4 (H2 (i))
    grad_tol = np.maximum(1e-4 * np.abs(g64) + gbud, np.abs(g32.astype(np.float64) - g64))
    grad_frac = np.abs(gg - g64) / grad_tol
    plain = np.abs(gg - g64) <= 1e-4 * np.abs(g64)
    nw = np.linalg.norm(gg - g64, axis=1) / np.linalg.norm(g64, axis=1)
    nw32 = np.linalg.norm(g32 - g64, axis=1) / np.linalg.norm(g64, axis=1)
    return {
        "systems": int(len(idx)), "which": f"first {K} + {K} random systems of the timed run (seed 4242)",
        "oracle": "oracle/dynax_oracle.c::oracle_rc_loss_grad_budget_f64 (fp64, forward mode), outside the timed region",
        "oracle_seconds": round(oracle_s, 2),
        "loss_max_rel_err": float((np.abs(lg - l64) / np.abs(l64)).max()),
        "loss_max_frac_of_tolerance": float(loss_frac.max()),     # tolerance = 1e-5*|loss| + what a 1e-5 trajectory error does to the loss
        "grad_frac_elements_within_plain_rel_1e-4": float(plain.mean()),
        "grad_worst_normwise_err": float(nw.max()), "grad_median_normwise_err": float(np.median(nw)),
        "grad_max_frac_of_budgeted_tolerance": float((np.abs(gg - g64) / (1e-4 * np.abs(g64) + gbud)).max()),
        "grad_max_frac_of_tolerance": float(grad_frac.max()),     # tolerance = max(1e-4*|g| + conditioning budget, error of the fp32 reference-order BPTT)
        "fp32_reference_port": {"grad_frac_elements_within_plain_rel_1e-4": float((np.abs(g32 - g64) <= 1e-4 * np.abs(g64)).mean()),
                                "grad_worst_normwise_err": float(nw32.max()), "grad_median_normwise_err": float(np.median(nw32))},
        "ok": bool(loss_frac.max() <= 1.0 and grad_frac.max() <= 1.0 and nw.max() <= 1e-4),
    }


# ------------------------------------------------------------------------------------------------
# BASELINE.json config 3: loss+gradient over N candidates / buildings sharing ONE parameter vector, sharded over the
# ranks, gradients all-reduced over NCCL INSIDE the timed region (examples/3-inverse.py:112-113).  Strong scaling.
# ------------------------------------------------------------------------------------------------
def config3_leg(a, torch, dist, dev, rank, world):
    from dynax_b200 import ops
    from dynax_b200.distributed import shard_range
    N3, T, K3 = a.config3_systems, a.T, a.config3_steps
    theta = torch.tensor(NOMINAL, device=dev)

    def shard(r):
        lo, hi = shard_range(N3, r, world)
        n = hi - lo
        g = torch.Generator(device=dev).manual_seed(777 + 1000 * world + r)
        x0 = torch.tensor([[20.0, 30.0, 26.0]], device=dev) + torch.empty((n, 3), device=dev).uniform_(-1, 1, generator=g)
        u = torch.empty((T, n, 5), device=dev).uniform_(0, 1, generator=g)
        u[..., 0] = 15 + 10 * u[..., 0]
        tg = 20 + 2 * torch.empty((T, n), device=dev).uniform_(0, 1, generator=g)
        return x0, u, tg

    x0, u, tg = shard(rank)
    buf = torch.zeros((8,), device=dev)                       # [loss, grad[7]]: ONE buffer, all-reduced in place
    out = (buf[0:1], buf[1:8])

    def step(reduce=True):
        ops.rc_loss_grad(theta, x0, u, tg, DT, time_chunks=0, out=out)
        if reduce and world > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)

    def timed(fn, k):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / k

    for _ in range(3):
        step()
    ms = timed(step, K3)
    result = buf.double().clone()
    kernel_ms = timed(lambda: step(False), K3)
    ar_us = timed(lambda: dist.all_reduce(buf, op=dist.ReduceOp.SUM), 50) * 1e3 if world > 1 else 0.0
    # the 1-rank answer, on every rank: the same shards one after the other, summed in fp64
    del x0, u, tg
    ref = torch.zeros((8,), device=dev, dtype=torch.float64)
    for r in range(world):
        x0r, ur, tgr = shard(r)
        l, g, _ = ops.rc_loss_grad(theta, x0r, ur, tgr, DT, time_chunks=0)
        ref[0] += l.double()[0]
        ref[1:] += g.double()
        del x0r, ur, tgr
    err = float(((result - ref).norm() / ref.norm()).item())
    errs = torch.tensor([err], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    err = float(errs.item())
    assert err <= 1e-5, f"all-reduced shared-theta loss/gradient differs from the 1-rank result: {err:.3e}"
    return {
        "metric": "system-steps/s, fwd+grad RC rollout, ONE shared parameter vector (BASELINE.json configs[2])",
        "value": N3 * T / (ms * 1e-3), "unit": "system-steps/s", "ms_per_step": ms, "steps": K3, "scaling": "strong",
        "systems_total": N3, "systems_per_rank": shard_range(N3, rank, world)[1] - shard_range(N3, rank, world)[0], "T": T,
        "collective": "NCCL all-reduce (sum) of 8 fp32 [loss, grad[7]] per step, inside the timed region" if world > 1 else "none (1 rank)",
        "kernel_only_ms": kernel_ms, "allreduce_us": ar_us,
        "vs_1rank_result_rel_err": err,
        "inputs": "per-system x0, u[T,n,5], target[T,n] resident on each rank (13.8 GB in total), theta shared",
    }


def _rel(a, ref, floor):
    """max |a - ref| / max(|ref|, floor * max|ref|): tests/util.py::rel_err (floor 1e-3 for temperatures, 1 = inf-norm)."""
    import numpy as np
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float((np.abs(a - ref) / np.maximum(np.abs(ref), floor * np.abs(ref).max() + 1e-300)).max())


# In-run parity samples of the other legs: the oracle as the CHECKER of what was just timed, outside the timed regions.
def parity_forward(torch, ops, view, dev):
    import numpy as np
    from oracle import dynax_oracle as orc
    th_t, x0_t, u_t, _, _, n = view
    xs, ys, _ = ops.rc_forward(th_t, x0_t, u_t, DT, time_chunks=0)
    idx = np.sort(np.random.default_rng(4243).choice(n, min(32, n), replace=False))
    sel = torch.as_tensor(idx, device=dev)
    x64, y64 = orc.rc_rollout(th_t[sel].cpu().numpy(), x0_t[sel].cpu().numpy(), u_t[:, sel].cpu().numpy(), DT, np.float64)
    ex, ey = _rel(xs[:, sel].cpu().numpy(), x64, 1e-3), _rel(ys[:, sel].cpu().numpy(), y64, 1e-3)
    return {"systems": int(len(idx)), "oracle": "oracle/dynax_oracle.py::rc_rollout (fp64)", "x_max_rel_err": ex,
            "y_max_rel_err": ey, "tolerance": 1e-5, "ok": bool(ex <= 1e-5 and ey <= 1e-5)}


def parity_mpc(torch, ops, nom, xm, d, um, sch, dev):
    import numpy as np
    from oracle import dynax_oracle as orc
    cost, amin, cmin = ops.rc_mpc_cost(nom, xm, d, um, *sch, 900.0)
    idx = np.sort(np.random.default_rng(4244).choice(um.shape[1], 64, replace=False))
    sel = torch.as_tensor(idx, device=dev)
    ref = orc.mpc_cost(np.asarray(NOMINAL, np.float64), [20.0, 35.8, 26.0], d.cpu().numpy(), um[:, sel].cpu().numpy(), 900.0,
                       dtype=np.float64)
    ec = _rel(cost[sel].cpu().numpy(), ref, 1.0)
    arg_ok = bool(amin.item() == int(torch.argmin(cost).item()) and cmin.item() == cost.min().item())
    return {"candidates": 64, "oracle": "oracle/dynax_oracle.py::mpc_cost (fp64)", "cost_max_rel_err": ec, "tolerance": 1e-5,
            "argmin_matches_costs": arg_ok, "ok": bool(ec <= 1e-5 and arg_ok)}


def parity_node(torch, ops, W, xn, un, dev):
    import numpy as np
    from oracle import dynax_oracle as orc
    xs, ys, _ = ops.node_rk4_forward(*W, xn, un, 0.25)
    idx = np.sort(np.random.default_rng(4245).choice(xn.shape[0], 16, replace=False))
    sel = torch.as_tensor(idx, device=dev)
    x64, y64 = orc.node_rk4_rollout(*[w.double().cpu().numpy() for w in W], xn[sel].cpu().numpy(), un[:, sel].cpu().numpy(),
                                    0.25, np.float64)
    ex, ey = _rel(xs[:, sel].cpu().numpy(), x64, 1.0), _rel(ys[:, sel].cpu().numpy(), y64, 1.0)
    return {"trajectories": 16, "oracle": "oracle/dynax_oracle.py::node_rk4_rollout (fp64; PARITY UNPINNED BY THE REFERENCE, "
            "pinned to a torch-fp64 restatement)", "x_max_rel_err": ex, "y_max_rel_err": ey, "tolerance": 1e-5,
            "ok": bool(ex <= 1e-5 and ey <= 1e-5)}


def other_configs_leg(a, torch, dist, dev, rank, world, views, T):
    """BASELINE.json configs[1], [3] and [4] on the same box, same run (weak scaling: the per-rank size is the config's):
    forward rollout emitting x and y (reference semantics) over the headline run's resident inputs, the MPC candidate
    cost + argmin, the MLP-dynamics RK4 rollout.  CUDA events, max over ranks; results stay on the device.  Rank 0 also
    checks a sample of each leg's results against the fp64 oracle, outside the timed regions (`parity` of each leg)."""
    import numpy as np
    from dynax_b200 import ops
    from dynax_b200.distributed import sharded_argmin
    from tools import _synth as syn

    def timed(fn, k, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / k

    peak, peak_src = measured_peak()
    out = {}
    # ---- configs[1]: N = 2^20 systems x T = 8760, x and y emitted (36 B per system-step), the headline run's tiles
    S = sum(v[5] for v in views)

    def fwd():
        for th_t, x0_t, u_t, tg_t, o, n in views:
            ops.rc_forward(th_t, x0_t, u_t, DT, time_chunks=0)
    ms = timed(fwd, 3, warm=1)
    v = world * S * T / (ms * 1e-3)
    out["config2_forward"] = {
        "metric": "system-steps/s, RC forward rollout emitting x and y (BASELINE.json configs[1])", "value": v,
        "unit": "system-steps/s", "ms_per_step": ms, "systems_per_rank": S, "T": T, "scaling": "weak",
        "launches_per_step": len(views), "alg_bytes_per_system_step": 36,
        "roofline_frac": S * T * 36 / (ms * 1e-3) / 1e9 / peak, "peak_gbs": peak, "peak_source": peak_src}
    if rank == 0 and not a.no_parity:   # 32 systems of the last tile launch against the fp64 oracle
        out["config2_forward"]["parity"] = parity_forward(torch, ops, views[-1], dev)
    # ---- configs[3]: 4M candidate control sequences x 96 steps through one shared model, cost + global argmin
    Nm, H = 1 << 22, 96
    g = torch.Generator(device=dev).manual_seed(3456 + rank)
    um = -10 * torch.rand((H, Nm), device=dev, generator=g)
    d = torch.rand((H, 4), device=dev, generator=g)
    d[:, 0] = 15 + 10 * d[:, 0]
    sch = [torch.tensor(x, dtype=torch.float32, device=dev) for x in (syn.RC_PRICE, syn.RC_T_LOW, syn.RC_T_HIGH)]
    nom = torch.tensor(NOMINAL, device=dev)
    xm = torch.tensor([20.0, 35.8, 26.0], device=dev)

    def mpc():
        cost, amin, cmin = ops.rc_mpc_cost(nom, xm, d, um, *sch, 900.0)
        if world > 1:
            sharded_argmin(cmin, amin, rank * Nm)
    ms = timed(mpc, 20)
    out["config4_mpc"] = {
        "metric": "candidate-steps/s, MPC cost of N candidate control sequences + argmin (BASELINE.json configs[3])",
        "value": world * Nm * H / (ms * 1e-3), "unit": "candidate-steps/s", "ms_per_step": ms, "candidates_per_rank": Nm,
        "horizon": H, "scaling": "weak", "alg_bytes_per_candidate_step": 4,
        "roofline_frac": Nm * H * 4 / (ms * 1e-3) / 1e9 / peak,
        "collective": "all-gather of one (cost, index) pair per rank inside the timed region" if world > 1 else "none (1 rank)"}
    if rank == 0 and not a.no_parity:   # 64 candidates against the fp64 oracle; the argmin against the costs themselves
        out["config4_mpc"]["parity"] = parity_mpc(torch, ops, nom, xm, d, um, sch, dev)
    del um
    # ---- configs[4]: MLP dynamics 8 -> 64 -> 64 -> 3 with RK4 on tcgen05, N = 262144 trajectories x T = 2048
    Nn, Tn = 262144, 2048
    W = [torch.tensor(w).to(dev) for w in syn.synth_mlp(4567)]
    xn = torch.randn((Nn, 3), device=dev, generator=g)
    un = torch.randn((Tn, Nn, 5), device=dev, generator=g)
    ms = timed(lambda: ops.node_rk4_forward(*W, xn, un, 0.25), 2, warm=1)
    flop = 2 * (8 * 64 + 64 * 64 + 3 * 64) * 4
    out["config5_node"] = {
        "metric": "trajectory-steps/s, MLP-dynamics RK4 rollout (BASELINE.json configs[4])",
        "value": world * Nn * Tn / (ms * 1e-3), "unit": "trajectory-steps/s", "ms_per_step": ms, "trajectories_per_rank": Nn,
        "T": Tn, "scaling": "weak", "hidden": 64, "mlp_tflops_fp32_grade": Nn * Tn * flop / (ms * 1e-3) / 1e12,
        "tensor_tflops_issued_tf32": 3 * Nn * Tn * 4 * 2 * (8 * 64 + 64 * 64) / (ms * 1e-3) / 1e12,
        "kernel": "node_tc_kernel<64> (tcgen05.mma kind::tf32, 3xTF32 split, accumulators in TMEM)"}
    if rank == 0 and not a.no_parity:   # 16 trajectories of the timed launch, all 2048 steps, against the fp64 oracle
        out["config5_node"]["parity"] = parity_node(torch, ops, W, xn, un, dev)
    return out


def link_probe(torch, dist, dev, world, nbytes=1 << 30, reps=4):
    """Plain pinned cudaMemcpyAsync H2D rate of this rank while all ranks copy at once (declared, outside the e2e
    timing): the ceiling the e2e leg can reach.  Returns (GB/s of the slowest rank, sum over ranks)."""
    h = torch.empty((nbytes,), dtype=torch.uint8, pin_memory=True)
    h.zero_()
    d = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    lo = torch.tensor([gbs], device=dev, dtype=torch.float64)
    tot = lo.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    del h, d
    return float(lo.item()), float(tot.item())


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def ours(a):
    import torch
    import torch.distributed as dist

    algo = ALGOS[a.algo]
    os.environ["DYNAX_GRAD_ALGO"] = algo["env"]     # read by the library at every call
    B_ALG = algo["alg_bytes"]
    if a.tile is None:
        a.tile = 3 * WAVE if a.algo == "tangent" else 4 * 148 * 3 * 128   # both 227328 systems
    from dynax_b200 import host, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_note = bind_to_gpu_numa_node(local) if world > 1 else None   # pinned e2e buffers land next to the GPU
    if world > 1:
        # the "NCCL version ..." banner goes to stdout when the first communicator is created: point fd 1 at stderr
        # while that happens, so that stdout carries the ONE JSON line and nothing else
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    T, S = a.T, a.systems
    tile = min(a.tile, S)
    tiles = [tile] * (S // tile) + ([S % tile] if S % tile else [])
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    nominal = torch.tensor(NOMINAL, device=dev)
    theta = nominal[None] * torch.exp(torch.empty((S, 7), device=dev).uniform_(-0.2, 0.2, generator=g))
    x0 = torch.tensor([[20.0, 30.0, 26.0]], device=dev) + torch.empty((S, 3), device=dev).uniform_(-1, 1, generator=g)
    u = torch.empty((T, tile, 5), device=dev).uniform_(0, 1, generator=g)
    u[..., 0] = 15 + 10 * u[..., 0]
    target = 20 + 2 * torch.empty((T, tile), device=dev).uniform_(0, 1, generator=g)
    loss = torch.empty((S,), device=dev)
    grad = torch.empty((S, 7), device=dev)
    offs = [sum(tiles[:i]) for i in range(len(tiles))]
    views = []
    for o, n in zip(offs, tiles):
        views.append((theta[o:o + n], x0[o:o + n], u if n == tile else u[:, :n].contiguous(),
                      target if n == tile else target[:, :n].contiguous(), o, n))

    def step():
        for th_t, x0_t, u_t, tg_t, o, n in views:
            l, gth, _ = ops.rc_loss_grad(th_t, x0_t, u_t, tg_t, DT)
            loss[o:o + n].copy_(l, non_blocking=True)
            grad[o:o + n].copy_(gth, non_blocking=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    sec = ms * 1e-3
    value = world * S * T * a.steps / sec
    launches = len(tiles) * a.steps
    assert bool(torch.isfinite(loss).all()) and bool(torch.isfinite(grad).all())

    # ---- in-run parity sample: systems of THIS timed run against the fp64 oracle (rank 0, outside the timed region)
    parity = None
    if not a.no_parity and rank == 0:
        parity = parity_block(torch, theta, x0, u, target, loss, grad, tile, a.parity_systems, T)

    # ---- BASELINE.json config 3: shared parameters, sharded systems, NCCL all-reduce in the timed region
    workloads = {}
    if not a.no_config3:
        workloads["config3"] = config3_leg(a, torch, dist, dev, rank, world)
    if not a.no_other_configs:
        workloads.update(other_configs_leg(a, torch, dist, dev, rank, world, views, T))

    # ---- end to end: pinned HOST buffers -> host-API call (tiled H2D / kernel / D2H inside) -> host results
    e2e = None
    if not a.no_e2e:
        import numpy as np
        Ne = a.e2e_systems
        hu = torch.empty((T, Ne, 5), dtype=torch.float32, pin_memory=True)
        ht = torch.empty((T, Ne), dtype=torch.float32, pin_memory=True)
        hth = torch.empty((Ne, 7), dtype=torch.float32, pin_memory=True)
        hx0 = torch.empty((Ne, 3), dtype=torch.float32, pin_memory=True)
        nsrc = min(Ne, tile)
        reps = (Ne + nsrc - 1) // nsrc
        hu.copy_(u[:, :nsrc].repeat(1, reps, 1)[:, :Ne]); ht.copy_(target[:, :nsrc].repeat(1, reps)[:, :Ne])
        hth.copy_(theta[:Ne]); hx0.copy_(x0[:Ne])
        torch.cuda.synchronize()
        link_min, link_sum = link_probe(torch, dist, dev, world)      # declared: plain pinned H2D copies, all ranks at once
        args = (hth.numpy(), hx0.numpy(), hu.numpy(), ht.numpy(), DT)
        host.rc_loss_grad(*args)          # warm-up (allocates the cached device scratch)
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            l_h, g_h = host.rc_loss_grad(*args)
        e_sec = time.perf_counter() - t0
        te = torch.tensor([e_sec], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_sec = float(te.item())
        assert np.isfinite(l_h).all()
        e2e = {"value": world * Ne * T * a.e2e_steps / e_sec, "unit": "system-steps/s",
               "h2d_bytes_per_step": int(hu.numel() + ht.numel() + hth.numel() + hx0.numel()) * 4,
               "d2h_bytes_per_step": int(l_h.size + g_h.size) * 4,
               "systems_per_step": Ne, "ms_per_step": e_sec / a.e2e_steps * 1e3,
               "api": "dynax_rc_loss_grad_host_f32 (pinned host buffers; tiled H2D/kernel/D2H on 2 streams)",
               "numa": numa_note}
        # the link ceiling: a plain pinned cudaMemcpyAsync H2D stream per rank, all ranks copying at once
        e2e["link_gbs_per_rank_min"] = link_min
        e2e["link_gbs_all_ranks"] = link_sum
        e2e["h2d_gbs_per_rank"] = e2e["h2d_bytes_per_step"] / (e2e["ms_per_step"] * 1e-3) / 1e9
        e2e["link_frac"] = e2e["h2d_gbs_per_rank"] / link_min
        # declared variant: u/target uploaded ONCE (dynax_rc_dataset_*), theta/x0 in and loss/grad out per call --
        # what the 5000-epoch loop of examples/3-inverse.py:133-138 does with its one measured series
        ds = host.RcDataset(hu.numpy(), ht.numpy())
        ds.loss_grad(hth.numpy(), hx0.numpy(), DT)
        barrier()
        nrep = 5 * a.e2e_steps
        t0 = time.perf_counter()
        for _ in range(nrep):
            l_d, g_d = ds.loss_grad(hth.numpy(), hx0.numpy(), DT)
        d_sec = time.perf_counter() - t0
        td = torch.tensor([d_sec], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        # (the tiled host call runs its last, small tile on the time-parallel kernels: same results within the
        # gradient tolerance, not bit for bit)
        assert np.allclose(l_d, l_h, rtol=1e-4)
        assert (np.linalg.norm(g_d - g_h, axis=1) <= 1e-4 * np.linalg.norm(g_h, axis=1)).all()
        ds.close()
        e2e["resident_dataset_variant"] = {
            "value": world * Ne * T * nrep / float(td.item()), "unit": "system-steps/s",
            "h2d_bytes_per_step": int(hth.numel() + hx0.numel()) * 4, "d2h_bytes_per_step": int(l_d.size + g_d.size) * 4,
            "ms_per_step": float(td.item()) / nrep * 1e3,
            "api": "dynax_rc_dataset_loss_grad_f32 (u/target resident in HBM across calls; NOT the cold figure above)"}
        if world == 1:
            # what a NumPy caller really passes: pageable memory (the driver stages it), then the same arrays page-locked
            # in place by dynax_host_register (registration timed separately: a one-off for long-lived arrays)
            Np = min(Ne, 8192)
            pu, pt = np.array(hu.numpy()[:, :Np]), np.array(ht.numpy()[:, :Np])
            pth, px0 = np.array(hth.numpy()[:Np]), np.array(hx0.numpy()[:Np])
            host.rc_loss_grad(pth, px0, pu, pt, DT)
            t0 = time.perf_counter()
            host.rc_loss_grad(pth, px0, pu, pt, DT)
            p_sec = time.perf_counter() - t0
            t0 = time.perf_counter()
            with host.pinned(pu, pt):
                r_reg = time.perf_counter() - t0
                host.rc_loss_grad(pth, px0, pu, pt, DT)
                t0 = time.perf_counter()
                host.rc_loss_grad(pth, px0, pu, pt, DT)
                r_sec = time.perf_counter() - t0
            e2e["pageable_variant"] = {"systems_per_step": Np, "value": Np * T / p_sec, "unit": "system-steps/s",
                                       "h2d_gbs": (pu.nbytes + pt.nbytes) / p_sec / 1e9,
                                       "registered_value": Np * T / r_sec, "registered_h2d_gbs": (pu.nbytes + pt.nbytes) / r_sec / 1e9,
                                       "register_ms": r_reg * 1e3,
                                       "api": "dynax_rc_loss_grad_host_f32 on pageable NumPy arrays / after dynax_host_register"}
            del pu, pt
        # informational: the same loss+gradient when the disturbances are ONE shared series and only the
        # control channel is per system: 8 instead of 24 B/step over PCIe
        hc = torch.empty((T, Ne), dtype=torch.float32, pin_memory=True)
        hc.copy_(hu[:, :, 2])
        hd = hu[:, 0, :].contiguous()
        cargs = (hth.numpy(), hx0.numpy(), hd.numpy(), hc.numpy(), ht.numpy(), DT)
        host.rc_loss_grad_ctrl(*cargs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            host.rc_loss_grad_ctrl(*cargs)
        c_sec = time.perf_counter() - t0
        tc_ = torch.tensor([c_sec], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tc_, op=dist.ReduceOp.MAX)
        e2e["shared_disturbance_variant"] = {
            "value": world * Ne * T * a.e2e_steps / float(tc_.item()), "unit": "system-steps/s",
            "h2d_bytes_per_step": int(hc.numel() + ht.numel() + hd.numel() + hth.numel() + hx0.numel()) * 4,
            "api": "dynax_rc_loss_grad_ctrl_host_f32 (shared u[T,5] + per-system ctrl[T,N]; not the headline workload)"}
        host.release()
        del hu, ht, hc

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_kind = measured_peak()
    achieved = S * T * a.steps * B_ALG / sec / 1e9          # per rank: alg bytes of its launches / its time
    traffic = ncu_traffic_per_step(a.algo)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)" if peak_kind == "measured" else "fallback 6.65 TB/s",
                "kernel": algo["kernel"], "alg_bytes_per_system_step": B_ALG,
                "alg_bytes_per_launch": B_ALG * tile * T, "avg_launch_ms": ms / launches,
                "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_note": (traffic or {}).get("note")}

    cpu = None
    if not a.no_cpu_baseline and world == 1:
        v, cores, s = time_cpu(a.cpu_systems, T, 2, 1)
        cpu = {"value": v, "unit": "system-steps/s", "cores": cores, "kind": "port",
               "sample": f"{a.cpu_systems} systems x {T} steps, oracle/dynax_oracle.c (all-fp32 BPTT, OpenMP, -march=native)",
               "numba_cross_check": time_cpu_numba(a.cpu_systems, T)}

    line = {
        "metric": "system-steps/s, fwd+grad RC rollout", "value": value, "unit": "system-steps/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"rc_fwd_grad N={S} T={T} dt=3600 per-system theta/u/target (per rank)",
                   "algo": a.algo,
                   "tile_systems": tile, "tiles_per_step": len(tiles),
                   "l2": "inputs larger than L2 (u+target tile %.1f GB re-read by each tile launch)" % ((u.numel() + target.numel()) * 4 / 1e9),
                   "parallelism": f"system axis sharded over {world} rank(s); no data-path collective"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk.summary(),
        "parity": parity, "workloads": workloads,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours(a)


if __name__ == "__main__":
    main()
