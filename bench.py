#!/usr/bin/env python
"""bench.py — env-steps/s of the VMAS physics hot path behind ``Environment.step`` on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload: BASELINE.json configs[1] — scenario ``balance``, 32768 envs per GPU, 4 agents,
continuous random actions pre-generated before the timed region (weak scaling: every rank
steps its own independent shard of envs; no data-path collective).

One JSON line is printed by rank 0:
  value        whole-job env-steps/s, actions already resident in HBM
  e2e          same metric through the public API with HOST buffers: per step the actions are
               copied from pinned host memory and obs/rewards/dones are copied back to it
  roofline     the fused substep kernel: algorithmic bytes per launch / CUDA-event duration
  cpu_baseline the CPU oracle port of the same env on this box's host cores (bounded sample)
Timing: per-iteration CUDA events on the launching stream, summed; L2 is flushed (512 MiB
memset) between iterations outside the brackets; max over ranks.

``--impl reference`` times the CPU oracle port of the path (the reference is pure Python and
does not travel to the GPU box; the oracle is bit-identical to it, see tests/) on all host
threads, through the same Environment API.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SCENARIO = "balance"
SCENARIO_KWARGS = dict(n_agents=4)
ENVS_PER_GPU = 32768
METRIC = "env-steps/sec"


# --------------------------------------------------------------------------------------------
def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=1000)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    p.add_argument("--cpu-steps", type=int, default=None, help="steps of the CPU baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-flush", action="store_true", help="keep L2 warm between iterations (not a bench value)")
    p.add_argument("--no-graph", action="store_true", help="step eagerly instead of replaying a CUDA graph")
    return p.parse_args()


def dist_info():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    QUERY = (
        "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
        "clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL,
                text=True,
            )
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_first_sample(self, timeout_s: float = 8.0):
        """nvidia-smi takes a while to attach (and slows launches while it does): the timed
        region must not start before its first line has arrived."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.02)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {
            "sm_mhz": sm[len(sm) // 2] if sm else None,
            "sm_max_mhz": mx,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def pregenerate_actions(env, steps, seed, device, pin=False):
    """[steps][n_agents] tensors of shape [B, action_size], U(-u_range, u_range), from a CPU generator."""
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(steps):
        per_agent = []
        for a in env.agents:
            u = (torch.rand(env.num_envs, a.action_size, generator=gen) * 2 - 1) * a.action.u_range_tensor.cpu()
            if pin:
                u = u.pin_memory()
            else:
                u = u.to(device)
            per_agent.append(u)
        out.append(per_agent)
    return out


# --------------------------------------------------------------------------------------------
def usable_cpus() -> int:
    """Host cores this process may actually use (affinity mask and cgroup quota, not os.cpu_count)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except Exception:  # noqa: BLE001
        pass
    return n


def run_cpu_oracle_env(n_envs, max_steps, time_budget_s=20.0):
    """Times the CPU oracle port behind the same Environment API on the host cores.

    The intra-op thread count is calibrated first (one step each at a few candidates up to the
    usable core count; eager torch on many tiny ops gets *slower* with too many threads), then
    steps are timed until ``max_steps`` or ``time_budget_s``.
    Returns (env_steps_per_s, seconds, steps, threads).
    """
    import vectorizedmultiagentsimulator_b200 as b200
    from oracle.backend import use_oracle

    cores = usable_cpus()
    candidates = sorted({c for c in (cores, 64, 32, 16, 8) if c <= cores}, reverse=True) or [1]
    torch.set_num_threads(candidates[-1])
    with use_oracle():
        env = b200.make_env(SCENARIO, num_envs=n_envs, device="cpu", seed=0, **SCENARIO_KWARGS)
        actions = pregenerate_actions(env, 4, seed=1, device="cpu")
        env.step(actions[0])  # warm-up (allocator, plan compile)
        best, best_t = candidates[-1], float("inf")
        for c in reversed(candidates):  # small thread counts first; stop when it gets worse
            torch.set_num_threads(c)
            t0 = time.perf_counter()
            env.step(actions[1])
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = c, dt
            elif dt > 1.5 * best_t:
                break
        torch.set_num_threads(best)
        steps, t0 = 0, time.perf_counter()
        while steps < max_steps:
            env.step(actions[steps % 4])
            steps += 1
            if time.perf_counter() - t0 > time_budget_s and steps >= 2:
                break
        dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt, steps, best


def main_reference(args):
    rank, world, _ = dist_info()
    if rank != 0:
        return
    n_envs = args.envs_per_gpu
    # bounded sample: at most --steps env steps and ~60 s of CPU time
    value, seconds, steps, cores = run_cpu_oracle_env(n_envs, max(1, args.steps), time_budget_s=60.0)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": "env-steps/s",
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": 2,
        "ms_per_step": 1e3 * seconds / steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": f"{SCENARIO} n_agents=4, {n_envs} envs, random continuous actions, CPU",
            "note": "CPU oracle port of the reference path (bit-identical to the reference, tests/); the pure-Python reference does not travel to the GPU box",
        },
        "cpu_baseline": {
            "value": value,
            "unit": "env-steps/s",
            "cores": cores,
            "kind": "port",
            "sample": f"{steps} env steps of {n_envs} envs ({seconds:.1f} s), {cores} intra-op threads (calibrated) of {usable_cpus()} usable cores",
        },
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def main_b200(args):
    rank, world, local = dist_info()
    import torch.distributed as dist

    if world > 1:
        # NCCL_DEBUG=VERSION makes NCCL print a banner on stdout, in front of the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)

    import vectorizedmultiagentsimulator_b200 as b200
    from vectorizedmultiagentsimulator_b200.simulator import plan as P

    B, K, W = args.envs_per_gpu, args.steps, max(args.warmup, 3)
    env = b200.make_env(
        SCENARIO, num_envs=B, device=device, seed=rank, cuda_graph=not args.no_graph, **SCENARIO_KWARGS
    )
    backend = env.world._get_backend()
    backend.refresh()
    desc = backend.tables.desc
    bytes_per_env_substep = P.algorithmic_bytes_per_env_substep(desc)

    flush = None if args.no_flush else torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(step_fn, n):
        """Σ over iterations of the CUDA-event time of step_fn(i); L2 flushed outside the brackets."""
        pairs = []
        for i in range(n):
            if flush is not None:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step_fn(i)
            e1.record()
            pairs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in pairs)  # ms

    # ---- arm 1: actions resident in HBM --------------------------------------------------
    dev_actions = pregenerate_actions(env, W + K, seed=1 + rank, device=device)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # samples every timed region below (Environment.step, kernel-only, e2e)
    for t in range(W):
        env.step(dev_actions[t])
    if rank == 0:
        sampler.wait_first_sample()
    barrier()
    launches_before = backend.launches
    backend.kernel_events = []
    wall0 = time.perf_counter()
    ms_total = timed_loop(lambda i: env.step(dev_actions[W + i]), K)
    wall = time.perf_counter() - wall0
    kernel_pairs = backend.kernel_events
    backend.kernel_events = None
    launches = backend.launches - launches_before
    barrier()
    # (empty in graph mode: the step is one graph replay, no per-kernel events inside it)
    kernel_in_step_ms = sum(a.elapsed_time(b) for a, b in kernel_pairs) / len(kernel_pairs) if kernel_pairs else 0.0

    # ---- the substep kernel alone: world.step() back to back, L2 flushed before every launch.
    # The flush (~100 us on the GPU) lets the host queue the next launch ahead, so the event
    # bracket around the kernel holds no host latency (inside Environment.step it does).
    backend.kernel_events = []
    for _ in range(K):
        if flush is not None:
            flush.zero_()
        env.world.step()
    torch.cuda.synchronize()
    kernel_ms = sum(a.elapsed_time(b) for a, b in backend.kernel_events) / K
    backend.kernel_events = None

    # ---- arm 2: end to end with host buffers --------------------------------------------------
    host_actions = pregenerate_actions(env, W + K, seed=101 + rank, device=device, pin=True)
    act_dev = [torch.empty_like(a, device=device) for a in host_actions[0]]
    obs0, rew0, done0, _ = env.step(dev_actions[0])
    # the step's results are read back with one device->host copy per kind (observations of all
    # agents, rewards of all agents, done): the per-agent tensors are stacked on the device first
    host_obs = torch.empty((len(obs0),) + tuple(obs0[0].shape), dtype=obs0[0].dtype).pin_memory()
    host_rew = torch.empty((len(rew0),) + tuple(rew0[0].shape), dtype=rew0[0].dtype).pin_memory()
    host_done = torch.empty(done0.shape, dtype=done0.dtype).pin_memory()
    h2d_bytes = sum(a.numel() * a.element_size() for a in host_actions[0])
    d2h_bytes = sum(t.numel() * t.element_size() for t in (host_obs, host_rew, host_done))

    def e2e_step(i):
        # pinned host actions go straight into Environment.step (it copies them to the device)
        obs, rews, dones, _ = env.step(host_actions[W + i])
        host_obs.copy_(torch.stack(obs), non_blocking=True)
        host_rew.copy_(torch.stack(rews), non_blocking=True)
        host_done.copy_(dones, non_blocking=True)

    for t in range(3):
        e2e_step(t - W)
    barrier()
    ms_e2e = timed_loop(e2e_step, K)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    env.check_actions_now()

    # ---- reduce over ranks ---------------------------------------------------------------------
    stats = torch.tensor([ms_total, ms_e2e, kernel_ms, kernel_in_step_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, kernel_ms, kernel_in_step_ms = (float(x) for x in stats.tolist())
    value = world * B * K / (ms_total * 1e-3)
    e2e_value = world * B * K / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the fused substep kernel -----------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = bytes_per_env_substep * B  # one launch advances B envs by one substep (S=1 here)
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = traffic_src = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj["%s_%s" % (SCENARIO, "_".join(f"{k}={v}" for k, v in SCENARIO_KWARGS.items()))][str(B)]
        traffic = ent["dram_bytes_read"] + ent["dram_bytes_write"]
        traffic_src = ent["source"]
    except Exception:  # noqa: BLE001
        pass
    roofline = {
        "bound": "hbm",
        "kernel": "substep kernel, mapping=%s" % backend._dev_tables.mapping,
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": traffic,
        "traffic_source": traffic_src,
        "peak_source": peak_src,
        "bytes_per_launch": alg_bytes,
        "bytes_per_env_substep": bytes_per_env_substep,
        "kernel_us": kernel_ms * 1e3,
        "kernel_us_inside_env_step": kernel_in_step_ms * 1e3 if kernel_in_step_ms else None,
        "how": "CUDA events recorded by the library around the substep kernel; standalone world.step() loop, L2 flushed before each launch",
    }

    # ---- same kernel at a batch that is not launch/latency-bound: 1 Mi envs (state tiled) -----------------
    try:
        big_B = 1 << 20
        reps = big_B // B
        from vectorizedmultiagentsimulator_b200 import _native as nat

        class _BigSlab:
            def __init__(self, slab):
                self.t = tuple(t.repeat(reps, *([1] * (t.dim() - 1))).contiguous() for t in slab.tensors())

            def tensors(self):
                return self.t

        big = _BigSlab(env.world.slab)
        old = backend.tables.desc.batch_dim
        backend.tables.desc.batch_dim = big_B
        big_dt = nat.DeviceTables(backend.tables, None, device)
        backend.tables.desc.batch_dim = old
        times = []
        for _ in range(12):
            if flush is not None:
                flush.zero_()
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            nat.world_step(backend.lib, big_dt, big, events=ev)
            torch.cuda.synchronize()
            times.append(ev[0].elapsed_time(ev[1]))
        times = sorted(times[2:])
        big_ms = times[len(times) // 2]
        big_achieved = bytes_per_env_substep * big_B / (big_ms * 1e-3) / 1e9
        roofline["at_1Mi_envs"] = {
            "kernel_us": big_ms * 1e3,
            "achieved": big_achieved,
            "frac": big_achieved / peak,
            "note": "same kernel and state tiled to 1,048,576 envs (slab 365 MB > L2); the 32768-env launch "
            "is 11.4 MB = 1.7 us of HBM time, below launch latency",
        }
        del big, big_dt
    except Exception as err:  # noqa: BLE001
        roofline["at_1Mi_envs"] = {"error": str(err)}

    # ---- CPU baseline (bounded sample, rank 0, N=1 only) ---------------------------------------------
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        v, seconds, cpu_steps, cores = run_cpu_oracle_env(B, args.cpu_steps or 200, time_budget_s=20.0)
        cpu_baseline = {
            "value": v,
            "unit": "env-steps/s",
            "cores": cores,
            "kind": "port",
            "sample": f"{cpu_steps} env steps of {B} envs ({seconds:.1f} s), CPU oracle port behind the same "
            f"Environment API, {cores} intra-op threads (calibrated) of {usable_cpus()} usable cores",
        }

    line = {
        "metric": METRIC,
        "value": value,
        "unit": "env-steps/s",
        "n_gpus": world,
        "steps": K,
        "warmup": W,
        "ms_per_step": ms_total / K,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": f"{SCENARIO} n_agents=4, {B} envs per GPU, 1 substep, random continuous actions (BASELINE.json configs[1])",
            "timing": "sum of per-iteration CUDA-event brackets around Environment.step; max over ranks",
            "l2": "flushed between iterations (512 MiB memset outside the brackets)" if flush is not None else "NOT flushed",
            "api": "make_env(..., cuda_graph=%s); Environment.step" % (not args.no_graph),
            "wall_ms_per_step_incl_flush": 1e3 * wall / K,
        },
        "clocks": clocks,
        "e2e": {
            "value": e2e_value,
            "unit": "env-steps/s",
            "h2d_bytes_per_step": h2d_bytes,
            "d2h_bytes_per_step": d2h_bytes,
            "ms_per_step": ms_e2e / K,
        },
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_b200(args)
