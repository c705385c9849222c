#!/usr/bin/env python
"""bench.py -- QP solves/s (eps_abs = eps_rel = 1e-5) of the batched MPC hot path on B200.

A "step" is one pass of the hot path over one batch of synthetic input.  The default workload is BASELINE.json
config 2 (the configuration the metric is quoted on that fits one GPU): the reference plant and horizon
(config/MPC_API.json, n = 15, m = 30), 4096 independent controllers per GPU with random x0 / U / references
sharing P and A, each solved cold.  `--config c3 | c4 | c5` runs the other BASELINE configurations with the same
JSON contract (c3: 12-state quadrotor N = 50, 131072 controllers per GPU = the 1M batch over 8 GPUs; c4: 65536
per-instance plants N = 30; c5: warm-started closed loop of 65536 controllers, N = 100).  Weak scaling: every rank
owns its own instances, no data-path collective; NCCL carries only the barrier and the max-over-ranks.

  value        device-resident: inputs already in HBM; per step set_state (device) + controllerStep (config 2: the one-call form controller_step_from)
  e2e          the same public call with HOST (pinned) inputs in and the control / status out inside the timed region
  roofline     the ADMM kernel against the FP64 pipe (DFMA / DMMA, one shared peak on B200; measured in this run);
               it keeps all iterates on chip, so HBM is touched once per solve -- the HBM-equivalent of a
               one-launch-per-iteration design (24 (n + 2m) B per problem-iteration) is reported beside it
  cpu_baseline the CPU oracle (oracle/, an OSQP-equivalent restatement; osqp-eigen itself is not installable
               here) on the host cores, one solver per core, bounded sample

`--impl reference` times that CPU path alone with the same metric / config (rank 0 only).

With no `--config` (the driver's call) the line is config 2 and carries a `configs` object with the other BASELINE
configurations measured in the same run at reduced step counts -- c1 (single controller: latency), c3, c4, c5, each with the
same fields (value, ms_per_step, e2e, roofline, cpu_baseline, clocks) -- so that every configuration is driver-visible at
every N; `--config cK` runs one configuration alone.  The CPU arm is the perf build of the port (oracle.PERF_FLAGS, built on
this host), one controller per core on all host threads, for every configuration.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

EPS = 1e-5
METRIC = "QP solves/sec (eps 1e-5)"
CPU_NOTE = "one OSQP-equivalent solver per core (oracle/osqp_port.c, perf build; osqp-eigen is not installable offline)"


def load_plant(path):
    """The reference's MPC_API.json (keys as read at src/ModelPredictiveControlAPI.cpp:16,19,113-116,138-140) as numpy arrays.
    The product arm parses the config itself: nothing of oracle/ is imported outside the cpu_baseline / reference legs."""
    with open(path) as f:
        cfg = json.load(f)
    sc = lambda k: float(np.array(cfg[k]).reshape(-1)[0])
    return dict(Ad=np.array(cfg["Ad"], dtype=np.float64), Bd=np.array(cfg["Bd"], dtype=np.float64).reshape(-1),
                Cd=np.array(cfg["Cd"], dtype=np.float64).reshape(-1), K=np.array(cfg["K"], dtype=np.float64).reshape(-1),
                Q=sc("Q"), R=sc("R"), RD=sc("RD"))


def load_traffic(key, batch):
    """dram bytes (read + write) per launch of this configuration's ADMM kernel from the committed ncu capture
    (profiles/traffic.json, written by tools/make_traffic.py); None when there is no capture at this batch size."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        t = json.load(f).get(key)
    if not t or t.get("batch_per_gpu") != batch:
        return None, None
    return t["dram_bytes_read"] + t["dram_bytes_write"], t["source"]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while True:   # first sample at once, then every 2 ms, and one more after the stop request (short timed regions)
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                if self._stop_evt.is_set():
                    break
                time.sleep(0.002)
        except Exception as e:  # NVML missing: report that instead of inventing clocks
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "samples": len(s),
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ workloads
class Workload:
    """One BASELINE configuration: builds the controllers of this rank, steps them, and knows its flop / byte counts."""
    key = name = ""
    default_batch = 0
    l2_note = "inputs larger than L2 (no flush needed)"
    flush_l2 = False

    def __init__(self, batch, rank):
        self.B, self.rank = batch, rank

    # -- CPU side (oracle): returns (solves, seconds, cores) for `count` QPs of this workload on `threads` cores
    def cpu_solve(self, count, threads):
        raise NotImplementedError

    def executed_flops(self, n, m_exec):
        """flops one instance-iteration of the kernel executes: n x n KKT product + the two products with the m_exec stored rows"""
        return 2.0 * (n * n + 2 * m_exec * n)


E2E_DEPTH = 4   # controller batches in flight in the sustained end-to-end pass (tests/dev/dev_e2e_depth.py: 1: 110 us per step, 2: 72, 3: 68, 4: 66, 6: 65)


class C2(Workload):
    key, default_batch, flush_l2 = "c2", 4096, True
    l2_note = "flushed with a 256 MB write between timed steps"
    N = 15

    def describe(self):
        return f"config2: reference plant N=15 (n=15, m=30), {self.B} random x0/U/ref per GPU sharing P and A, cold solves"

    def _inputs(self, count, seed):
        from problems import c2_batch
        return c2_batch(count, seed=seed)

    def _conf(self):
        return os.path.join(ROOT, "config", "MPC_API.json")

    def setup(self, sm, torch, device, kernel):
        X, U, ref = self._inputs(self.B, 1000 * self.rank)
        self.host = [np.ascontiguousarray(a) for a in (X, U, ref)]
        self.mpc = sm.BatchedModelPredictiveControlAPI(self._conf(), batch=self.B, device=device, eps_abs=EPS, eps_rel=EPS, kernel=kernel)
        self.mpc.solver.set_cold_solves(True)
        self.dev = [torch.from_numpy(a).cuda() for a in self.host]
        self.pin = [torch.from_numpy(a).pin_memory() for a in self.host]
        self.out_u = torch.empty(self.B, dtype=torch.float64).pin_memory()
        self.out_st = torch.empty(self.B, dtype=torch.int32).pin_memory()
        self.n, self.m = self.mpc.n_variables, self.mpc.n_constraints
        self.solver = self.mpc.solver
        self.h2d = int(sum(a.nbytes for a in self.host))
        self.d2h = int(self.B * 8 + self.B * 4)

    def step_device(self):
        # the body of the reference's loop (solver.cpp:45-55): write X, U, ref and run controllerStep, one public call
        self.mpc.controller_step_from(self.dev[0], self.dev[1], self.dev[2])

    def step_e2e(self):
        if not getattr(self, "_bound", False):
            self.mpc.bind_results(self.out_u, self.out_st)                     # D2H: the step itself writes the control and the status
            self._bound = True
        self.mpc.controller_step_from(self.pin[0], self.pin[1], self.pin[2])   # H2D from pinned memory (read by the step's first kernel)
        self.mpc.sync()

    def pipelined_e2e(self, sm, torch, device, kernel, steps, warmup, depth=2):
        """Sustained end-to-end throughput of the public calls: `depth` controller batches (this one and twins with their own
        pinned buffers and streams) take turns, the host waits for step k - depth only after it has enqueued step k - 1, so the
        PCIe reads, the launches and the host wake-up of one batch overlap the solves of the others.  Every step still reads its
        inputs from pinned host memory and writes control + status back to pinned host memory; the host looks at every step's
        statuses."""
        ctl = [(self.mpc, self.pin, self.out_st.numpy())]
        keep = []
        for t in range(1, depth):
            twin = sm.BatchedModelPredictiveControlAPI(self._conf(), batch=self.B, device=device, eps_abs=EPS, eps_rel=EPS, kernel=kernel)
            twin.solver.set_cold_solves(True)
            tstream = torch.cuda.Stream()
            twin.set_stream(tstream.cuda_stream)
            X, U, ref = self._inputs(self.B, 1000 * self.rank + 500 * t)
            tpin = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (X, U, ref)]
            tout_u, tout_st = torch.empty(self.B, dtype=torch.float64).pin_memory(), torch.empty(self.B, dtype=torch.int32).pin_memory()
            twin.bind_results(tout_u, tout_st)
            ctl.append((twin, tpin, tout_st.numpy()))
            keep.append((tstream, tout_u, tout_st))

        def run(count):
            solved = True
            for k in range(count):
                mpc, pin, st = ctl[k % depth]
                if k >= depth:
                    mpc.sync()                                   # step k - depth of this controller: results are in its pinned buffers
                    solved = solved and bool((st == 1).all())
                mpc.controller_step_from(pin[0], pin[1], pin[2])
            for mpc, _, st in ctl:
                mpc.sync()
                solved = solved and bool((st == 1).all())
            return solved
        run(max(2 * depth, warmup))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ok = run(steps)
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / steps
        for twin, _, _ in ctl[1:]:
            twin.close()
        assert ok, "a pipelined end-to-end step did not reach SOLVED on every instance"
        return ms

    def nnz_A(self):
        return int(np.count_nonzero(self.mpc.matrix("Gbar")))

    def launches(self):
        return self.mpc.launches

    def close(self):
        self.mpc.close()

    def cpu_solve(self, count, threads):
        import oracle
        cfg = oracle.load_config(self._conf())
        m = oracle.mpc_build(**{**cfg, "N": self.N})
        X, U, ref = self._inputs(count, 0)
        f, ub = oracle.mpc_batch_vectors(m, X, U, ref)
        out = oracle.solve_batch(m["H"], m["Gbar"], m["lb"], m["W0"], f, ub, settings=oracle.default_settings(eps_abs=EPS, eps_rel=EPS), nthreads=threads, perf=True)
        return count, out["seconds"], threads


class C3(Workload):
    key, default_batch = "c3", 131072

    def executed_flops(self, n, m_exec):
        # the x-space variant of the tile kernel multiplies by V' and V only (two n x n GEMMs), the products with A are element-wise
        if "x-space" in self.solver.kernel_name:
            return 4.0 * n * n
        return super().executed_flops(n, m_exec)

    def describe(self):
        return (f"config3: 12-state / 4-input quadrotor, N=50, condensed n=200, m=400 (input box as +/- rows), {self.B} random "
                "x0 / hover references per GPU (= the 1M batch over 8 GPUs) sharing P and A, cold solves")

    def _conf(self):
        return os.path.join(ROOT, "config", "quadrotor.json")

    def setup(self, sm, torch, device, kernel):
        from problems import c3_batch
        x0, xr = c3_batch(self.B, seed=1000 * self.rank)
        self.host = [np.ascontiguousarray(x0), np.ascontiguousarray(xr)]
        self.mpc = sm.BatchedMimoMPC(self._conf(), batch=self.B, device=device, eps_abs=EPS, eps_rel=EPS, kernel=kernel)
        self.mpc.solver.set_cold_solves(True)
        self.dev = [torch.from_numpy(a).cuda() for a in self.host]
        self.pin = [torch.from_numpy(a).pin_memory() for a in self.host]
        self.out_u = torch.empty(self.B * self.mpc.nu, dtype=torch.float64).pin_memory()
        self.out_st = torch.empty(self.B, dtype=torch.int32).pin_memory()
        self.n, self.m = self.mpc.n_variables, self.mpc.n_constraints
        self.solver = self.mpc.solver
        self.h2d = int(sum(a.nbytes for a in self.host))
        self.d2h = int(self.out_u.numel() * 8 + self.B * 4)

    def step_device(self):
        self.mpc.set_state(x0=self.dev[0], xr=self.dev[1])
        self.mpc.controller_step_async()

    def step_e2e(self):
        self.mpc.set_state(x0=self.pin[0], xr=self.pin[1])
        self.mpc.controller_step_async()
        self.mpc.control_into(self.out_u)
        self.solver.status_into(self.out_st)

    def nnz_A(self):
        return 2 * self.n

    def launches(self):
        return self.mpc.launches

    def close(self):
        self.mpc.close()

    def cpu_solve(self, count, threads):
        import oracle
        from problems import c3_batch
        m = oracle.mimo_build(**oracle.load_mimo_config(self._conf()))
        x0, xr = c3_batch(count, seed=0)
        q = oracle.mimo_batch_vectors(m, x0, xr)
        out = oracle.solve_batch(m["H"], m["A"], m["lb"], m["ub"], q, np.tile(m["ub"], (count, 1)),
                                 settings=oracle.default_settings(eps_abs=EPS, eps_rel=EPS), nthreads=threads, perf=True)
        return count, out["seconds"], threads


class C4(C2):
    key, default_batch, flush_l2 = "c4", 65536, False
    l2_note = "inputs larger than L2 (no flush needed)"
    N = 30

    def describe(self):
        return (f"config4: per-instance linearised plants (distinct P, A per problem; reference dimensions), N=30 (n=30, m=60), "
                f"{self.B} controllers per GPU, cold solves; the per-plant setup (on-device assembly, scaling, factorisation and "
                f"pencil eigen-decomposition) runs once at create time, outside the timed step: config.setup_ms")

    def setup(self, sm, torch, device, kernel):
        from problems import c4_plants
        cfg = load_plant(self._conf())
        Ad, Bd = c4_plants(self.B, cfg, seed=2 + 1000 * self.rank)
        conf = dict(Ad=Ad, Bd=Bd, Cd=cfg["Cd"], K=cfg["K"], Q=cfg["Q"], R=cfg["R"], RD=cfg["RD"], horizon=self.N, per_instance=1)
        X, U, ref = self._inputs(self.B, 31 + 1000 * self.rank)
        self.host = [np.ascontiguousarray(a) for a in (X, U, ref)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        self.mpc = sm.BatchedModelPredictiveControlAPI(conf, batch=self.B, device=device, eps_abs=EPS, eps_rel=EPS)
        torch.cuda.synchronize()
        self.setup_ms = 1e3 * (time.perf_counter() - t0)   # upload of the plants + every create-time kernel (the CPU arm's constructor)
        self.mpc.solver.set_cold_solves(True)
        self.dev = [torch.from_numpy(a).cuda() for a in self.host]
        self.pin = [torch.from_numpy(a).pin_memory() for a in self.host]
        self.out_u = torch.empty(self.B, dtype=torch.float64).pin_memory()
        self.out_st = torch.empty(self.B, dtype=torch.int32).pin_memory()
        self.n, self.m = self.mpc.n_variables, self.mpc.n_constraints
        self.solver = self.mpc.solver
        self.h2d = int(sum(a.nbytes for a in self.host))
        self.d2h = int(self.B * 8 + self.B * 4)

    def nnz_A(self):
        return self.N * (self.N + 1)

    def cpu_solve(self, count, threads):
        import oracle
        from problems import c4_plants
        cfg = oracle.load_config(self._conf())
        Ad, Bd = c4_plants(count, cfg, seed=2)
        X, U, ref = self._inputs(count, 31)
        # one controller per plant and per core: constructor (builders, scaling, factor) + controllerStep, as a per-plant
        # ModelPredictiveControlAPI object pays it (oracle/batch_drivers.c)
        out = oracle.plant_batch(cfg, Ad, Bd, X, U, ref, self.N, settings=oracle.default_settings(eps_abs=EPS, eps_rel=EPS),
                                 nthreads=threads, perf=True)
        return count, out["seconds"], threads


class C5(C2):
    key, default_batch, flush_l2 = "c5", 65536, False
    l2_note = "per-step working set (q, u, iterates, solution of 65536 x (100 + 200)) larger than L2 (no flush possible inside the closed loop)"
    N = 100
    AMP, PERIOD = 0.1, 200

    def describe(self):
        return (f"config5: closed-loop warm-started MPC, reference plant N=100 (n=100, m=200), {self.B} controllers per GPU, "
                "square-wave references (amplitude 0.1, period 200 steps, random phase), synthetic plant step on device; "
                "one step = one controllerStep of every controller + plant step")

    def setup(self, sm, torch, device, kernel):
        cfg = load_plant(self._conf())
        conf = dict(Ad=cfg["Ad"], Bd=cfg["Bd"], Cd=cfg["Cd"], K=cfg["K"], Q=cfg["Q"], R=cfg["R"], RD=cfg["RD"], horizon=self.N)
        X, U, _ = self._inputs(self.B, 31 + 1000 * self.rank)
        self.host = [np.ascontiguousarray(X * 0.2), np.ascontiguousarray(U * 0.1), np.zeros(self.B)]
        self.phase = np.random.default_rng(5 + self.rank).integers(0, self.PERIOD, self.B).astype(np.int32)
        self.mpc = sm.BatchedModelPredictiveControlAPI(conf, batch=self.B, device=device, eps_abs=EPS, eps_rel=EPS, kernel=kernel)
        self.mpc.set_state(X=self.host[0], U=self.host[1], ref=self.host[2])
        self.n, self.m = self.mpc.n_variables, self.mpc.n_constraints
        self.solver = self.mpc.solver
        self.pin = [torch.from_numpy(a).pin_memory() for a in self.host]
        self.out_u = torch.empty(self.B, dtype=torch.float64).pin_memory()
        self.out_st = torch.empty(self.B, dtype=torch.int32).pin_memory()
        self.h2d = int(self.B * 4 * 8)            # e2e: the measured state X goes in every step (the plant is outside)
        self.d2h = int(self.B * 8 + self.B * 4)   # the control U and the status come back

    def closed_loop(self, steps, graph):
        """`steps` closed-loop steps on the device; returns (not_solved, iterations)."""
        return self.mpc.closed_loop(steps, self.AMP, self.PERIOD, self.phase, use_graph=graph)

    def step_e2e(self):
        # hardware-in-the-loop shape of the reference's main loop (solver.cpp:43-74): state in, controllerStep, control out
        self.mpc.set_state(X=self.pin[0])
        self.mpc.controller_step_async()
        self.mpc.results_into(self.out_u, self.out_st)

    def nnz_A(self):
        return self.N * (self.N + 1)

    def cpu_solve(self, count, threads):
        import oracle
        cfg = oracle.load_config(self._conf())
        X, U, _ = self._inputs(count, 31)
        phase = np.random.default_rng(5).integers(0, self.PERIOD, count).astype(np.int32)
        steps = 6                           # step 0 is the cold solve (run, not counted), then warm-started steps
        out = oracle.closed_loop(cfg, X * 0.2, U * 0.1, self.N, steps, self.AMP, self.PERIOD, phase, skip=1,
                                 settings=oracle.default_settings(eps_abs=EPS, eps_rel=EPS), nthreads=threads, perf=True)
        return count * (steps - 1), out["seconds"], threads


WORKLOADS = {w.key: w for w in (C2, C3, C4, C5)}


def cpu_solves_per_s(wl, min_seconds, threads, first=None):
    """The CPU oracle (perf build) on this workload's QPs, one controller per core on all host threads; bounded sample.
    Returns (solves/s, solves, seconds, cores used)."""
    count = first or {"c2": 4096, "c3": 8 * threads, "c4": 64 * threads, "c5": 8 * threads}[wl.key]
    done, secs, cores = 0, 0.0, 1
    while done == 0 or secs < min_seconds:
        c, s, cores = wl.cpu_solve(count, threads)
        done += c
        secs += s
    return done / secs, done, secs, cores


SUB_STEPS = {"c2": 50, "c3": 3, "c4": 5, "c5": 20}      # timed steps of a configuration measured inside the all-configs line
BOUND = {"c2": "fp64", "c3": "tensor", "c4": "fp64", "c5": "tensor"}   # DFMA kernels vs the DMMA (FP64 tensor pipe) tile kernel
C1_CASES = [dict(X=[.01, 0, .02, 0], U=0.0, ref=0.0), dict(X=[0, 0, .05, 0], U=0.0, ref=0.0),
            dict(X=[.02, -.1, .03, .2], U=0.5, ref=0.25), dict(X=[.1, .5, .08, -.2], U=-1.0, ref=-0.3)]   # SURVEY 8(c) cases A-D
C1_WORKLOAD = ("config1: shipped config/MPC_API.json plant and horizon (n=15, m=30), ONE controller (batch 1): latency of "
               "controllerStep in a warm-started closed loop (square-wave reference +-0.1, period 200, synthetic plant) and of a "
               "cold controllerStep from the states of SURVEY 8(c) cases A-D")


def cpu_c1_latency(steps):
    """Config 1 on the CPU port (perf build, one core): microseconds per controllerStep, warm closed loop and cold cases."""
    import oracle
    cfg = oracle.load_config(os.path.join(ROOT, "config", "MPC_API.json"))
    st = oracle.default_settings(eps_abs=EPS, eps_rel=EPS)
    out = oracle.closed_loop(cfg, np.array([[0.0, 0.0, 0.05, 0.0]]), np.zeros(1), 15, 20 + steps, 0.1, 200, None, skip=20,
                             settings=st, nthreads=1, perf=True)
    warm = 1e6 * out["step_seconds"][20:]
    cold = []
    for _ in range(max(1, steps // 8)):
        for c in C1_CASES:
            o = oracle.closed_loop(cfg, np.array([c["X"]]), np.array([c["U"]]), 15, 1, c["ref"], 0, None, skip=0, settings=st,
                                   nthreads=1, perf=True)
            cold.append(1e6 * o["step_seconds"][0])     # timed around the step only (the constructor is outside)
    return float(np.mean(warm)), float(np.median(warm)), float(np.mean(cold)), out["iterations"] / steps


def reference_sample(key, threads, batch=None):
    """Workload `key` and the size of one bounded CPU sample of it (a `step` of the reference arm)."""
    wl = WORKLOADS[key](batch or WORKLOADS[key].default_batch, 0)
    return wl, {"c2": wl.B, "c3": 4 * threads, "c4": 32 * threads, "c5": 4 * threads}[key]


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    import oracle
    flags = oracle.PERF_FLAGS

    def arm(key, steps, warmup):
        wl, count = reference_sample(key, threads, args.batch if args.config == key else None)
        per_step, solves, cores = [], 0, threads
        for i in range(warmup + steps):
            c, s, cores = wl.cpu_solve(count, threads)     # one bounded sample of the workload per step
            if i >= warmup:
                per_step.append(s)
                solves = c
        ms = 1e3 * float(np.mean(per_step))
        value = solves / (ms / 1e3)
        sample = f"{steps} samples of {solves} solves of this workload, {CPU_NOTE}, gcc {flags}"
        return {"value": value, "unit": "solves/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
                "config": {"workload": wl.describe(), "eps_abs": EPS, "eps_rel": EPS, "solves_per_step": solves},
                "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample, "flags": flags},
                "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}

    head_key = "c2" if args.config == "all" else args.config
    r = arm(head_key, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": r["config"],
            "cpu_baseline": r["cpu_baseline"], "e2e": r["e2e"], "gpu_launches": 0}
    if args.config == "all":
        cfgs = {}
        for key in ("c3", "c4", "c5"):
            cfgs[key] = arm(key, 3, 1)
        wm, wmed, cold, its = cpu_c1_latency(400)
        cfgs["c1"] = {"config": {"workload": C1_WORKLOAD}, "latency_us": {"cpu": wm, "cpu_median": wmed, "cpu_cold": cold},
                      "value": 1e6 / wm, "unit": "solves/s", "iters_mean": its,
                      "cpu_baseline": {"value": 1e6 / wm, "unit": "solves/s", "cores": 1, "kind": "port",
                                       "sample": f"400 warm closed-loop steps of one controller, {CPU_NOTE}, gcc {flags}", "flags": flags}}
        line["configs"] = cfgs
    print(json.dumps(line), flush=True)


class Bench:
    """The GPU arm: owns the process group, the stream and the timing helpers; measure() runs one configuration."""

    def __init__(self, args, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        import solvempc_b200 as sm
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: solvempc_b200 has no CPU fallback")
        self.torch, self.dist, self.sm = torch, dist, sm
        self.args, self.rank, self.local_rank, self.world = args, rank, local_rank, world
        torch.cuda.set_device(local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        self.stream = torch.cuda.current_stream()
        self.flush = None
        self.fp64_peak = None
        self._pending_cpu = []

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def dgemm_peak(self):
        """FP64 denominator: MEASURED_PEAKS.json has none, so a cuBLAS DGEMM is measured here."""
        if self.fp64_peak is None:
            torch = self.torch
            a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
            b = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
            best = 1e9
            for _ in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            self.fp64_peak = 2 * 4096 ** 3 / (best * 1e-3) / 1e12
        return self.fp64_peak

    # ------------------------------------------------------------------------------------------ one configuration
    def measure(self, key, steps, warmup, batch=0, cpu_seconds=10.0):
        torch, args, world, rank = self.torch, self.args, self.world, self.rank
        wl = WORKLOADS[key](batch or WORKLOADS[key].default_batch, rank)
        B = wl.B
        wl.setup(self.sm, torch, self.local_rank, args.kernel)
        stream = self.stream
        wl.mpc.set_stream(stream.cuda_stream)
        if wl.flush_l2 and self.flush is None:
            self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")   # 256 MB > 126 MB L2
        flush = self.flush if wl.flush_l2 else None
        closed_loop = key == "c5"
        barrier = self.barrier
        sampler = ClockSampler(self.local_rank)
        fresh = None
        # ---- device-resident value
        if closed_loop:
            bad, _ = wl.closed_loop(max(warmup, 3), graph=True)
            assert bad == 0, "warm-up closed-loop steps did not reach SOLVED on every instance"
            wl.solver.enable_timing(True)
            wl.solver.kernel_ms(reset=True)
            launches0 = wl.launches()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            sampler.start()
            wall0 = time.perf_counter()
            e0.record(stream)
            bad, total_iters = wl.closed_loop(steps, graph=False)   # (kernel timing events cannot be captured into a graph)
            e1.record(stream)
            barrier()
            wall_ms = 1e3 * (time.perf_counter() - wall0)
            assert bad == 0
            step_ms = np.array([e0.elapsed_time(e1) / steps])
            prob_iters = total_iters / steps                          # per step
            iters_mean, iters_max = total_iters / (steps * B), None
        else:
            for _ in range(warmup):
                wl.step_device()
            torch.cuda.synchronize()
            assert wl.solver.count_solved() == B, "warm-up solve did not reach SOLVED on every instance"
            iters = wl.solver.info()["iter"].astype(np.int64)

            def timed_pass(step):
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
                barrier()
                w0 = time.perf_counter()
                for k, (e0, e1) in enumerate(evs):
                    if flush is not None:
                        flush.zero_()                 # evict the inputs from L2 between timed steps
                    e0.record(stream)
                    step(k)
                    e1.record(stream)
                barrier()
                return np.array([a.elapsed_time(b) for a, b in evs]), 1e3 * (time.perf_counter() - w0)

            # pass 1 (value): exactly K steps, nothing but the public calls inside the timed region.  Pass 2 (roofline): the same K
            # steps with the library's CUDA events around the ADMM kernel -- they sit between the kernels of a step and switch off
            # the programmatic-dependent-launch overlap, so they are kept out of the pass that produces `value`.
            launches0 = wl.launches()
            sampler.start()
            step_ms, wall_ms = timed_pass(lambda k: wl.step_device())
            launches = wl.launches() - launches0
            if key == "c2":
                # pass 1b: the same K steps, every step on inputs it has never seen (fresh seeds): the scheduler's difficulty classes
                # and the queue order are recomputed per step anyway; this shows ms_per_step does not lean on a repeated batch
                sets = [[torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in wl._inputs(B, 777 + 1000 * rank + 17 * k)] for k in range(steps)]
                wl.mpc.controller_step_from(*sets[0]); torch.cuda.synchronize()
                fr_ms, _ = timed_pass(lambda k: wl.mpc.controller_step_from(*sets[k]))
                torch.cuda.synchronize()
                fresh = {"ms_per_step": float(fr_ms.mean()), "value": world * B / (float(fr_ms.mean()) / 1e3),
                         "solved_last_step": int(wl.solver.count_solved()),
                         "note": "every timed step on a batch drawn from its own seed (never seen before)"}
                wl.step_device(); torch.cuda.synchronize()
            wl.solver.enable_timing(True)
            wl.solver.kernel_ms(reset=True)
            step_ms_k, _ = timed_pass(lambda k: wl.step_device())
            prob_iters, iters_mean, iters_max = int(iters.sum()), float(iters.mean()), int(iters.max())
        if closed_loop:
            launches = wl.launches() - launches0
            step_ms_k = step_ms
        kern_ms, kern_n = wl.solver.kernel_ms(reset=True)
        wl.solver.enable_timing(False)
        ms_per_step = self.max_over_ranks(float(step_ms.mean()))
        value = world * B / (ms_per_step / 1e3)

        # ---- end to end through the public call with host buffers
        for _ in range(max(3, warmup)):
            wl.step_e2e()
        barrier()
        t_e2e = []
        for _ in range(steps):
            if flush is not None:
                flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            wl.step_e2e()
            torch.cuda.synchronize()
            t_e2e.append(time.perf_counter() - t0)
        barrier()
        assert (wl.out_st.numpy() == 1).all()
        e2e_ms = self.max_over_ranks(1e3 * float(np.mean(t_e2e)))
        e2e_value = world * B / (e2e_ms / 1e3)
        e2e_extra = {}
        if key == "c2":
            # sustained throughput of the same calls with two batches in flight (what the CPU arm measures too: all cores busy, no
            # per-step wait); the one-step-at-a-time figure above stays in the line as e2e.sync_*
            pipe_ms = self.max_over_ranks(wl.pipelined_e2e(self.sm, torch, self.local_rank, args.kernel, max(steps, 40), warmup, depth=E2E_DEPTH))
            barrier()
            e2e_extra = {"sync_value": e2e_value, "sync_ms_per_step": e2e_ms,
                         "batches_in_flight": E2E_DEPTH,
                         "mode": f"{E2E_DEPTH} controller batches of this size take turns on {E2E_DEPTH} streams (each with its own pinned inputs / results; the "
                                 f"host waits for step k-{E2E_DEPTH} after enqueuing step k-1); every step reads its inputs from pinned host memory "
                                 "and writes control + status to pinned host memory; sync_* = one step at a time, synchronised per step, "
                                 "L2 flushed between steps"}
            e2e_ms, e2e_value = pipe_ms, world * B / (pipe_ms / 1e3)
        clocks = sampler.stop()

        res = None
        if rank == 0:
            peaks, peak_src = load_peaks()
            fp64_peak = self.dgemm_peak()
            n, m = wl.n, wl.m
            nnzA = wl.nnz_A()
            flops_per_launch = prob_iters * (2.0 * n * n + 4.0 * nnzA)       # 2 n^2 (KKT contraction) + 4 nnz(A) (A x, A'y)
            m_exec = wl.solver.row_pairs or m                                # [G; -G] row pairs: the kernels multiply the top half only
            executed_per_launch = prob_iters * wl.executed_flops(n, m_exec)  # what the kernel's iteration executes
            bytes_per_launch = prob_iters * 24.0 * (n + 2 * m)               # what one launch per iteration would stream
            kms = kern_ms / max(kern_n, 1)
            achieved = flops_per_launch / (kms * 1e-3) / 1e12
            traffic, traffic_src = load_traffic(wl.key, B)
            res = {
                "value": value, "unit": "solves/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
                "config": {"workload": wl.describe(), "batch_per_gpu": B, "eps_abs": EPS, "eps_rel": EPS, "adaptive_rho_interval": 25,
                           "l2": wl.l2_note, "kernel": wl.solver.kernel_name, "row_pairs_exploited": wl.solver.row_pairs,
                           "iters_mean": iters_mean, "iters_max": iters_max, "wall_ms_timed_region": wall_ms},
                "e2e": {"value": e2e_value, "unit": "solves/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": wl.h2d, "d2h_bytes_per_step": wl.d2h, **e2e_extra},
                "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": {"bound": BOUND[key], "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                             "traffic": traffic, "traffic_unit": "bytes of DRAM read + written per launch", "traffic_source": traffic_src,
                             "kernel": wl.solver.kernel_name, "kernel_ms": kms, "kernel_share_of_step": kms / float(step_ms_k.mean()),
                             "step_ms_in_kernel_timing_pass": float(step_ms_k.mean()),
                             "kernel_timing": "CUDA events around the ADMM kernel inside the library, second pass of the same K steps"
                                              if not closed_loop else "CUDA events around the ADMM kernel inside the library, same K steps",
                             "flops_per_launch": flops_per_launch, "executed_flops_per_launch": executed_per_launch,
                             "executed_frac": executed_per_launch / (kms * 1e-3) / 1e12 / fp64_peak,
                             "algorithmic_flops_per_instance_iteration": 2.0 * n * n + 4.0 * nnzA,
                             "bound_note": "fp64 = FP64 DFMA pipe, tensor = FP64 DMMA (mma.sync.m8n8k4.f64) pipe; the two share one measured peak on B200",
                             "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (MEASURED_PEAKS.json has no fp64 entry)",
                             "hbm_equivalent": {"achieved": bytes_per_launch / (kms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                                "frac": bytes_per_launch / (kms * 1e-3) / 1e9 / peaks["hbm_gbs"], "peak_source": peak_src,
                                                "note": "24(n+2m) B per problem-iteration that a one-launch-per-iteration kernel would stream; "
                                                        "this kernel keeps the iterates on chip and reads/writes HBM once per solve"}},
            }
            if wl.solver.kernel_name.startswith("admm_shared_tile"):
                # what bounds this kernel was measured, not assumed (A/B builds, no-load diagnostics, DMMA skeleton, L2 stream): DESIGN 4
                res["roofline"]["bound_evidence"] = "profiles/r02_tile_kernel_bound_study.json"
            if fresh is not None:
                res["fresh_inputs"] = fresh
            if getattr(wl, "setup_ms", None) is not None:
                # the CPU arm of this configuration pays the constructor per plant inside its timed region; the GPU arm pays it once
                # at create time: both ways of counting, so that nothing hides in the setup
                res["config"]["setup_ms"] = wl.setup_ms
                res["config"]["value_with_setup_every_step"] = world * B / ((ms_per_step + wl.setup_ms) / 1e3)
            self._pending_cpu.append((res, wl.__class__, B, cpu_seconds))
        wl.close()
        del wl
        torch.cuda.empty_cache()
        return res

    def cpu_legs(self):
        """The CPU arm of every measured configuration, after all GPU work (rank 0; the other ranks have nothing to wait for)."""
        import oracle
        threads = os.cpu_count() or 1
        for res, cls, B, secs in self._pending_cpu:
            wl = cls(B, 0)
            cpu_v, cpu_done, cpu_secs, cpu_cores = cpu_solves_per_s(wl, secs, threads)
            res["cpu_baseline"] = {"value": cpu_v, "unit": "solves/s", "cores": cpu_cores, "kind": "port", "flags": oracle.PERF_FLAGS,
                                   "sample": f"{cpu_done} solves of this workload's QPs in {cpu_secs:.1f} s, {CPU_NOTE}, gcc {oracle.PERF_FLAGS}"}
        self._pending_cpu = []

    # ------------------------------------------------------------------------------------------ one global batch over the ranks
    def sharded_batch_check(self, per_rank=4096):
        """ONE global config-2 batch split contiguously over the ranks (solvempc_b200.sharding.shard_bounds), every
        rank steps its shard on its own GPU, the controls and statuses are gathered over NCCL (the optional final gather), and
        rank 0 compares them bitwise with the whole batch stepped on its GPU alone.  Outside every timed region."""
        from solvempc_b200.sharding import gather_results, shard_arrays
        torch, sm = self.torch, self.sm
        wl = C2(per_rank, 0)
        total = per_rank * self.world + 3                       # ragged on purpose
        X, U, ref = wl._inputs(total, 4242)

        def step(Xs, Us, rs):
            mpc = sm.BatchedModelPredictiveControlAPI(wl._conf(), batch=Xs.shape[0], device=self.local_rank, eps_abs=EPS, eps_rel=EPS, kernel=self.args.kernel)
            mpc.set_stream(self.stream.cuda_stream)
            mpc.set_state(X=np.ascontiguousarray(Xs), U=np.ascontiguousarray(Us), ref=np.ascontiguousarray(rs))
            mpc.controllerStep()
            _, Uo = mpc.state()
            st = mpc.solver.info()["status"]
            mpc.close()
            return Uo, st

        Us, st = step(*shard_arrays((X, U, ref), self.world, self.rank))
        Ug, stg = gather_results(Us, total), gather_results(st, total)
        if self.rank != 0:
            return None
        U1, st1 = step(X, U, ref)
        return {"global_batch": total, "ranks": self.world, "gather": "nccl all_gather (solvempc_b200.sharding.gather_results)",
                "bitwise_equal_to_single_gpu": bool(np.array_equal(Ug, U1) and np.array_equal(stg, st1)),
                "solved": int((stg == 1).sum())}

    # ------------------------------------------------------------------------------------------ config 1: latency
    def measure_c1(self, steps):
        """BASELINE config 1 (parity and latency only): one controller, microseconds per controllerStep through
        (i) the C ABI with pinned host state in / control + status out, (ii) the reference's own unmodified class on the shim
        (oracle/_ref/ref_mpc_gpu, built in the build container), (iii) the device-resident CUDA-graph closed loop."""
        torch, sm = self.torch, self.sm
        if self.rank != 0:
            return None
        conf = os.path.join(ROOT, "config", "MPC_API.json")
        cfg = load_plant(conf)
        mpc = sm.BatchedModelPredictiveControlAPI(conf, batch=1, device=self.local_rank, eps_abs=EPS, eps_rel=EPS, kernel=self.args.kernel)
        mpc.set_stream(self.stream.cuda_stream)
        X = torch.tensor([[0.0, 0.0, 0.05, 0.0]], dtype=torch.float64).pin_memory()
        U = torch.zeros(1, dtype=torch.float64).pin_memory()
        ref = torch.zeros(1, dtype=torch.float64).pin_memory()
        out_u = torch.zeros(1, dtype=torch.float64).pin_memory()
        out_st = torch.zeros(1, dtype=torch.int32).pin_memory()
        mpc.bind_results(out_u, out_st)
        Xn, Un = X.numpy(), U.numpy()
        lat, its, bad = [], 0, 0
        launches0 = mpc.launches
        for s in range(20 + steps):
            ref[0] = 0.1 if 2 * (s % 200) < 200 else -0.1
            t0 = time.perf_counter()
            mpc.controller_step_from(X, U, ref)
            mpc.sync()
            t1 = time.perf_counter()
            if s >= 20:
                lat.append(1e6 * (t1 - t0))
                bad += int(out_st[0] != 1)
            Un[0] = out_u[0]
            Xn[0] = cfg["Ad"] @ Xn[0] + cfg["Bd"] * Un[0]
        launches = mpc.launches - launches0
        its = float(mpc.solver.info()["iter"][0])
        # cold controllerStep from the known-answer states (every solve from x = z = y = 0, rho = rho0)
        mpc.solver.set_cold_solves(True)
        cold = []
        for k in range(max(4, steps // 2)):
            c = C1_CASES[k % 4]
            Xn[0] = c["X"]; Un[0] = c["U"]; ref[0] = c["ref"]
            t0 = time.perf_counter()
            mpc.controller_step_from(X, U, ref)
            mpc.sync()
            cold.append(1e6 * (time.perf_counter() - t0))
        mpc.solver.set_cold_solves(False)
        mpc.bind_results(None, None)
        # device-resident closed loop, one CUDA-graph replay per step (no host round trip)
        mpc.set_state(X=np.array([[0.0, 0.0, 0.05, 0.0]]), U=np.zeros(1), ref=np.zeros(1))
        mpc.closed_loop(20, 0.1, 200, None, use_graph=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(self.stream)
        gbad, git = mpc.closed_loop(steps, 0.1, 200, None, use_graph=True)
        e1.record(self.stream)
        torch.cuda.synchronize()
        graph_us = 1e3 * e0.elapsed_time(e1) / steps
        kname = mpc.solver.kernel_name
        mpc.close()
        shim, shim_note = None, "oracle/_ref/ref_mpc_gpu not present (built only where /root/reference exists)"
        exe = os.path.join(ROOT, "oracle", "_ref", "ref_mpc_gpu")
        if os.path.exists(exe):
            import subprocess
            try:
                out = subprocess.run([exe, "latency", str(steps)], cwd=ROOT, env=dict(os.environ, SOLVEMPC_EPS=str(EPS), SOLVEMPC_DEVICE=str(self.local_rank)),
                                     capture_output=True, text=True, timeout=300)
                shim = json.loads(out.stdout)
                shim_note = "the reference's unmodified ModelPredictiveControlAPI.cpp + include/OsqpEigen shim + libsolvempc_b200.so, timed around controllerStep()"
            except Exception as e:   # report instead of failing the whole line
                shim_note = f"ref_mpc_gpu failed: {type(e).__name__}"
        lat = np.array(lat)
        res = {"config": {"workload": C1_WORKLOAD, "kernel": kname, "eps_abs": EPS, "eps_rel": EPS},
               "value": 1e6 / float(lat.mean()), "unit": "solves/s", "steps": steps, "iters_last_step": its, "not_solved": bad + int(gbad),
               "latency_us": {"gpu": float(lat.mean()), "gpu_median": float(np.median(lat)), "gpu_cold": float(np.mean(cold)),
                              "gpu_graph_closed_loop": graph_us,
                              "gpu_shim": None if shim is None else shim["latency_us_mean"],
                              "gpu_shim_median": None if shim is None else shim["latency_us_median"]},
               "latency_notes": {"gpu": "smpc_mpc_controller_step_from (pinned X, U, ref read in place; U + status written to bound pinned buffers) + smpc_mpc_sync, host clock, ctypes call overhead included",
                                 "gpu_graph_closed_loop": "smpc_mpc_closed_loop(use_graph=1): device-resident, CUDA events / steps (graph instantiation included)",
                                 "gpu_shim": shim_note},
               "e2e": {"value": 1e6 / float(lat.mean()), "unit": "solves/s", "h2d_bytes_per_step": 48, "d2h_bytes_per_step": 12},
               "gpu_launches": int(launches)}
        self._c1 = res
        return res

    def c1_cpu(self, res, steps):
        wm, wmed, cold, its = cpu_c1_latency(steps)
        import oracle
        res["latency_us"].update({"cpu": wm, "cpu_median": wmed, "cpu_cold": cold})
        res["cpu_baseline"] = {"value": 1e6 / wm, "unit": "solves/s", "cores": 1, "kind": "port", "flags": oracle.PERF_FLAGS,
                               "sample": f"{steps} warm closed-loop steps of one controller (mean {its:.1f} iterations), {CPU_NOTE}, gcc {oracle.PERF_FLAGS}"}


def run_ours(args, rank, local_rank, world):
    bench = Bench(args, rank, local_rank, world)
    head_key = "c2" if args.config == "all" else args.config
    if head_key == "c1":
        res = bench.measure_c1(args.steps)
        if rank == 0:
            bench.c1_cpu(res, args.steps)
            print(json.dumps({"metric": "controllerStep latency (eps 1e-5)", "higher_is_better": True, "n_gpus": 1, "dtype": "f64", "data": "synthetic", **res}), flush=True)
        return
    head = bench.measure(head_key, args.steps, args.warmup, args.batch, args.cpu_seconds)
    subs = {}
    if args.config == "all":
        for key in ("c3", "c4", "c5"):
            subs[key] = bench.measure(key, SUB_STEPS[key], 3, 0, args.cpu_seconds)
        subs["c1"] = bench.measure_c1(300)
    shard = None
    if world > 1:
        bench.barrier()
        shard = bench.sharded_batch_check()
        bench.barrier()
        bench.dist.destroy_process_group()
    if rank == 0:
        bench.cpu_legs()
        if "c1" in subs:
            bench.c1_cpu(subs["c1"], 300)
        line = {"metric": METRIC, "value": head["value"], "unit": "solves/s", "n_gpus": world, "steps": head["steps"],
                "warmup": head["warmup"], "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
        line.update({k: v for k, v in head.items() if k not in line})
        if subs:
            for v in subs.values():
                v.update({"metric": METRIC, "scaling": "weak", "dtype": "f64", "data": "synthetic", "higher_is_better": True})
            line["configs"] = subs
        if shard is not None:
            line["sharded_batch_check"] = shard
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="all", choices=["all", "c1"] + sorted(WORKLOADS),
                    help="all (default): config 2 as the headline + a `configs` object with c1, c3, c4, c5; cK: that configuration alone")
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (0 = the configuration's own size)")
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    if args.steps is None:
        args.steps = {"all": 50, "c1": 300, "c2": 50, "c3": 10, "c4": 10, "c5": 50}[args.config]
    args.warmup = max(args.warmup, 3)
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
