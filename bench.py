#!/usr/bin/env python
"""bench.py -- shooting-interval solves + sensitivities per second (FP64) on B200.

    python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   the restated reference CPU path on the host cores

One unit = one (interval, candidate) system: integrate [x; S] across the interval's control pieces and emit the end
state, the sensitivities and the shooting defect.  One step = one pass of the hot path over the whole batch.

N = 1   headline = BASELINE.json configs[3] `dense20_256x65k`, the largest single-GPU configuration (BASELINE.json
        quotes the metric on no particular config): synthetic 20-state problem x' = W x - x^3 + b u, 256 shooting
        intervals x 65 536 candidates, Tsit5 with K = 2 fixed steps on each of the M = 4 control pieces, the FULL
        sensitivity Jacobian (ns = 24 directions, nz = 500) -- 16.8 M units, 64 GB of sensitivities resident in HBM.
        `all_configs` carries configs 1 (latency), 2, 3 and 4, each with kernel time, roofline and e2e.
N > 1   BASELINE.json configs[4] `scale_1M`: 256 intervals x 4096 candidates of the same model = 2^20 units, STRONG
        scaling: the candidate axis is sharded over the ranks; inside the timed region every step ends with
        allreduce(objective_sum, grad_sum) and allgather(defects) over NVLink -- fused into the evaluation's own
        kernels through the library's communicator (cb200_comm_*, P2P windows), not NCCL calls from the host.
        `secondary` keeps the config-2 weak-scaling number of round 1.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "shooting-interval solves+sensitivities/sec (FP64)"

# ---- algorithmic work per unit (DESIGN.md "roofline"):
#   E = M (6K + 1) RHS evaluations (Tsit5, FSAL re-seeded at every piece start)
#   W = E (C_f + C_JS + C_F) + M K 42 nz flop;  Q = 8 (nz_in + np + nc_i + nz_out + n_defect + n_obj) bytes
WORK = {
    # config 4/5: M = 4, K = 2: E = 52; C_f = 2*400 + 80; C_JS = 2*400*24 + 3*20*24; nz = 20 * 25
    "dense20": dict(flops=52 * (880 + 20640) + 4 * 2 * 42 * 500, bytes=8 * (20 + 4 + 500 + 20)),
    # config 2: M = 4, K = 2: E = 52; C_f = 10, C_JS = 36; nz = 2 * 9
    "lqr": dict(flops=52 * (10 + 36) + 4 * 2 * 42 * 18, bytes=8 * (2 + 2 + 4 + 18 + 4)),
    # config 3: M = 2, K = 2: E = 26; C_f + C_JS + C_F = 51; nz = 9 (SURVEY.md 8d worked example)
    "lotka_oed": dict(flops=26 * 51 + 2 * 2 * 42 * 9, bytes=8 * (9 + 3 + 6 + 9 + 9)),
    # config 1: M = 5, K = 4: E = 125; Lotka3: C_f = 20, C_JS = 2*6*10 + 6; nz = 3 * 11
    "lotka_ms24": dict(flops=125 * (20 + 126) + 5 * 4 * 42 * 33, bytes=8 * (3 + 2 + 5 + 33 + 3 + 1)),
}

try:
    ORIG_AFFINITY = os.sched_getaffinity(0)      # before bind_to_gpu_numa_node narrows it
except AttributeError:
    ORIG_AFFINITY = None


def host_threads():
    """Threads for the CPU arms: every core this process may run on (torchrun exports OMP_NUM_THREADS=1, which
    would otherwise make the reference arm single-threaded)."""
    if ORIG_AFFINITY is not None:
        return max(1, len(ORIG_AFFINITY))
    return max(1, os.cpu_count() or 1)


def bind_to_gpu_numa_node(index):
    """Run this rank on the CPU cores next to its GPU (NVML's ideal affinity) BEFORE any pinned host buffer is
    allocated: first touch places the buffers on the local node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or allowed
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), "MEASURED_PEAKS.json (measured)"
    return 6650.0, "B200_PROFILING.md fallback (MEASURED_PEAKS.json absent)"


def ncu_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the config's dominant kernel, from the committed
    ncu --set full capture (profiles/r2_traffic.json: value, workload of the capture, file, date) -- or None."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get(kernel_key)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.start = [], None, index, 0

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None
        return self

    def wait_first(self, timeout=5.0):
        t_end = time.time() + timeout
        while self.proc and not self.rows and time.time() < t_end:
            time.sleep(0.02)

    def mark(self):
        """samples taken from now on are the ones reported (the GPU is under load from here)"""
        self.start = len(self.rows)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[self.start:]:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ================================================================================================== workloads
def dense20_inputs_soa(torch, dev, ps0, B, seed):
    """[nvars, B] on the device: every node and control value of the template perturbed by 0.1 N(0,1)
    (x0 ~ U[-1,1], u ~ U[-1,1] in the template)"""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    base = torch.from_numpy(np.asarray(ps0)).to(dev)
    x = torch.randn((len(ps0), B), dtype=torch.float64, device=dev, generator=g)
    x.mul_(0.1).add_(base[:, None])
    return x


def lqr_inputs(nvars, ps0, B, rng, n_intervals=100):
    """candidate matrix (B, nvars): nodes x ~ U[0,4], cost node 0, controls N(0,1)"""
    ps = np.tile(ps0, (B, 1))
    off = 0
    for i in range(n_intervals):
        nu0 = 0 if i == 0 else 2
        if nu0:
            ps[:, off] = rng.uniform(0, 4, B)
            ps[:, off + 1] = 0.0
        ps[:, off + nu0 + 2: off + nu0 + 6] = rng.standard_normal((B, 4))
        off += nu0 + 6
    assert off == nvars
    return ps


def oed_inputs(N, ps0, B, rng, n_intervals=64):
    ps = np.tile(ps0, (B, 1))
    blocks = N.block_structure()
    pp = rng.uniform(0.8, 1.2, (B, 2))
    for i in range(n_intervals):
        nu0 = 0 if i == 0 else 9
        ps[:, blocks[i] + nu0: blocks[i] + nu0 + 2] = pp
    return ps


def fp64_peak(nat, dev_index):
    """FP64 roof measured in this run: DFMA and DMMA share one datapath on B200 (tools/dmma_micro.cu), the higher probe is
    the peak.  MEASURED_PEAKS.json (driver-written) carries HBM and bf16 figures only."""
    a = nat.probe_fp64_tflops(dev_index, 40000)
    b = nat.probe_fp64_mma_tflops(dev_index, 20000)
    return max(a, b), (f"FP64 probes measured in this run, max of DFMA chains ({a:.2f}) and mma.sync.m8n8k4.f64 ({b:.2f}) "
                       "TFLOP/s (nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2); MEASURED_PEAKS.json has no FP64 figure")


def roofline_entry(work, units, kern_ms, peak, peak_src, hbm_peak, hbm_src, kernel, traffic_key):
    ach = work["flops"] * units / (kern_ms * 1e-3) / 1e12
    gbs = work["bytes"] * units / (kern_ms * 1e-3) / 1e9
    tr = ncu_traffic(traffic_key)
    return {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak if peak > 0 else None,
            "traffic": tr["bytes_per_unit"] * units if tr else None,
            "traffic_source": (f"{tr['file']} ({tr['date']}; ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one "
                               f"launch on {tr['workload']}: {tr['bytes_per_launch']:.4g} B for {tr['units_in_capture']:.0f} units = "
                               f"{tr['bytes_per_unit']:.1f} B per unit, times the units of this launch)") if tr else None,
            "algorithmic_bytes_per_launch": work["bytes"] * units, "kernel": kernel, "kernel_ms": kern_ms,
            "flops_per_unit": work["flops"], "bytes_per_unit": work["bytes"], "peak_source": peak_src,
            "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "peak_source": hbm_src}}


def timed_steps(torch, stream, step, steps, warmup, barrier=None, N=None):
    """W untimed steps, then K steps between CUDA events on the launching stream, a synchronize on both sides.
    Returns (milliseconds of the K steps, kernels the handle N launched inside the timed region)"""
    for s in range(warmup):
        step(s)
    (barrier or torch.cuda.synchronize)()
    l0 = N.launch_count() if N is not None else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s in range(steps):
        step(s)
    e1.record(stream)
    (barrier or torch.cuda.synchronize)()
    return e0.elapsed_time(e1), (N.launch_count() - l0 if N is not None else 0)


# -------------------------------------------------------------------------------------------------- config 4 (headline)
def run_config4(args, torch, dev, local_rank, clk, n_candidates=65536, e2e=True):
    from corleone_b200 import Tsit5, configs, native as nat, to_vec
    I = 256
    spec = configs.config4_dense20(I)
    lay, ps_t, st, N = configs.native_from_spec(spec, Tsit5(2), device=local_rank)
    B, nvars = n_candidates, N.nvars
    stream = torch.cuda.current_stream()
    N.set_stream(stream.cuda_stream)
    sz = N.out_sizes()
    d_ps = dense20_inputs_soa(torch, dev, to_vec(ps_t), B, configs.SEED0 + 4)          # [nvars, B], 3.2 GB: larger than L2
    fields = ("defects_kind", "xend")
    d_out = {f: torch.empty((sz[f], B), dtype=torch.float64, device=dev) for f in fields}
    # the sensitivity Jacobian (n_sens x B doubles = 64 GB) stays resident in handle-owned HBM (cb200_out.resident)
    want = nat.WANT_SENS | nat.WANT_DEFECTS | nat.WANT_XEND
    flags = nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA
    ptrs = {f: d_out[f].data_ptr() for f in fields}

    def step(s):
        N.eval_raw(d_ps.data_ptr(), B, B, want, flags, ptrs, 1, resident=nat.WANT_SENS)

    step(0)
    torch.cuda.synchronize()
    clk.mark()   # the GPU is under this workload's load from here on (a step takes most of a second)
    ms_total, launches = timed_steps(torch, stream, step, args.steps, max(3, args.warmup), N=N)
    kern_ms = float(np.mean(N.kernel_ms_history(min(256, args.steps))))
    units = I * B
    res = {"workload": "dense20_256x65k" if B == 65536 else f"dense20_256x{B}", "units_per_step": units, "e2e": None,
           "value": units * args.steps / (ms_total * 1e-3), "unit": "units/s", "ms_per_step": ms_total / args.steps,
           "kernel_ms": kern_ms, "steps": args.steps,
           "outputs": f"sens ({8 * sz['sens'] * B / 1e9:.1f} GB, resident in HBM), defects_kind, xend",
           "chunking": "none on the device-resident arm: one persistent launch over all units of the step (items of 8 candidates x 1 interval)"}
    res["_launches_timed"] = launches
    res["_N"] = N
    e2e_res = None
    if e2e:
        # ---- end to end through cb200_eval with HOST buffers: candidate-major (Julia nvars x B) pinned input, the
        # defects and end states come back to pinned host memory, the 64 GB Jacobian stays on the device (resident)
        e2e_steps = max(1, min(args.steps, args.e2e_steps or 3))
        pin_ps = torch.empty((B, nvars), dtype=torch.float64, pin_memory=True)
        pin_ps.copy_(d_ps.t())          # the same candidates, candidate-major
        del d_ps
        for f in fields:
            del d_out[f]
        torch.cuda.empty_cache()
        pin_out = {f: torch.empty((B, sz[f]), dtype=torch.float64, pin_memory=True) for f in fields}
        hp = {f: pin_out[f].data_ptr() for f in fields}

        def host_eval():
            N.eval_raw(pin_ps.data_ptr(), B, nvars, want, 0, hp, 1, resident=nat.WANT_SENS)

        host_eval()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_eval()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e_res = {"value": units * e2e_steps / dt, "unit": "units/s", "h2d_bytes_per_step": 8 * B * nvars,
                   "d2h_bytes_per_step": 8 * B * sum(sz[f] for f in fields), "steps": e2e_steps,
                   "ms_per_step": 1e3 * dt / e2e_steps,
                   "api": "cb200_eval with pinned host buffers, candidate-major (Julia layout), 8 candidate chunks on 3 streams; "
                          "defects_kind and xend return to the host, the sensitivity Jacobian stays device-resident "
                          "(cb200_out.resident = CB200_WANT_SENS, cb200_resident)"}
        del pin_ps, pin_out
    res["e2e"] = e2e_res
    return res


# -------------------------------------------------------------------------------------------------- config 2
def run_config2(args, torch, dev, local_rank, steps, rank=0):
    from corleone_b200 import Tsit5, configs, native as nat, to_vec
    I, B, K = 100, 4096, 2
    spec = configs.config2_lqr()
    lay, ps_t, st, N = configs.native_from_spec(spec, Tsit5(K), device=local_rank)
    nvars = N.nvars
    rng = np.random.default_rng(configs.SEED0 + 2 + 1000 * rank)
    ps_host = lqr_inputs(nvars, to_vec(ps_t), B, rng)
    stream = torch.cuda.current_stream()
    N.set_stream(stream.cuda_stream)
    sz = N.out_sizes()
    fields = ("jac_vals", "defects_kind", "objective", "grad_compact")   # cons_j, c, f, grad f (structural zeros omitted)
    sz["grad_compact"] = N.grad_structure(2)[1]
    set_bytes = 8 * B * (nvars + sum(sz[f] for f in fields))
    nsets = max(2, int(np.ceil(3 * 126e6 / set_bytes)))
    d_ps, d_out, d_sums = [], [], []
    for s in range(nsets):
        d_ps.append(torch.from_numpy(np.ascontiguousarray(np.roll(ps_host, s, axis=0).T)).to(dev))
        d_out.append({f: torch.empty((sz[f], B), dtype=torch.float64, device=dev) for f in fields})
        d_sums.append(torch.zeros(nvars + 1, dtype=torch.float64, device=dev))
    want = nat.WANT_JAC | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD | nat.WANT_GRAD_COMPACT
    flags = nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA

    def step(s):
        k = s % nsets
        ptrs = {f: d_out[k][f].data_ptr() for f in fields}
        ptrs["objective_sum"] = d_sums[k].data_ptr()
        ptrs["grad_sum"] = d_sums[k].data_ptr() + 8
        N.eval_raw(d_ps[k].data_ptr(), B, B, want, flags, ptrs, 2)

    ms_total, launches = timed_steps(torch, stream, step, steps, 3, N=N)
    kern_ms = float(np.mean(N.kernel_ms_history(min(256, steps))))
    units = I * B
    # e2e: pinned host buffers, the NLP consumer's outputs (f, c, cons_j, compact grad f)
    e2e_steps = min(steps, 10)
    pin_ps = torch.from_numpy(ps_host).pin_memory()
    pin_out = {f: torch.empty((B, sz[f]), dtype=torch.float64, pin_memory=True) for f in fields}
    hp = {f: pin_out[f].data_ptr() for f in fields}

    def host_eval(res_bits=0, flds=fields):
        N.eval_raw(pin_ps.data_ptr(), B, nvars, want, 0, {f: hp[f] for f in flds}, 2, resident=res_bits)

    def time_host(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / e2e_steps

    dt_full = time_host(host_eval)
    # the consumer that factorises cons_j on the device: only f, c and the compact gradient cross PCIe
    slim = ("defects_kind", "objective", "grad_compact")
    dt_slim = time_host(lambda: host_eval(nat.WANT_JAC, slim))
    # the consumer of the cross-candidate sums only: every per-candidate result stays in HBM
    h_sums = np.zeros(1 + nvars)

    def sums_eval():
        N.eval_raw(pin_ps.data_ptr(), B, nvars, want | nat.WANT_GRAD, 0,
                   {"objective_sum": h_sums.ctypes.data, "grad_sum": h_sums.ctypes.data + 8}, 2,
                   resident=nat.WANT_JAC | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD_COMPACT | nat.WANT_GRAD)

    dt_sums = time_host(sums_eval)
    out = {"workload": "lqr_ms100_b4096", "units_per_step": units, "value": units * steps / (ms_total * 1e-3), "unit": "units/s",
           "ms_per_step": ms_total / steps, "kernel_ms": kern_ms, "steps": steps,
           "l2": f"inputs larger than L2: {nsets} rotating buffer sets of {set_bytes / 1e6:.0f} MB",
           "e2e": {"value": units / dt_full, "unit": "units/s", "h2d_bytes_per_step": 8 * B * nvars,
                   "d2h_bytes_per_step": 8 * B * sum(sz[f] for f in fields), "steps": e2e_steps, "ms_per_step": 1e3 * dt_full,
                   "api": "cb200_eval, pinned host buffers, candidate-major; f, c, cons_j nonzeros, compact gradient to the host"},
           "e2e_jac_resident": {"value": units / dt_slim, "unit": "units/s", "h2d_bytes_per_step": 8 * B * nvars,
                                "d2h_bytes_per_step": 8 * B * sum(sz[f] for f in slim), "steps": e2e_steps,
                                "ms_per_step": 1e3 * dt_slim,
                                "api": "the same with cons_j kept on the device (cb200_out.resident = CB200_WANT_JAC)"},
           "e2e_sums_only": {"value": units / dt_sums, "unit": "units/s", "h2d_bytes_per_step": 8 * B * nvars,
                             "d2h_bytes_per_step": 8 * (1 + nvars), "steps": e2e_steps, "ms_per_step": 1e3 * dt_sums,
                             "api": "the same with every per-candidate result device-resident; objective_sum and grad_sum "
                                    "(6.4 KB) return to the host"},
           "gpu_launches": launches, "_N": N, "_spec": spec, "_ps_host": ps_host}
    return out


# -------------------------------------------------------------------------------------------------- config 3
def run_config3(args, torch, dev, local_rank, steps):
    from corleone_b200 import Tsit5, configs, native as nat, to_vec
    I, B = 64, 16384
    spec = configs.config3_lotka_oed(I)
    lay, ps_t, st, N = configs.native_from_spec(spec, Tsit5(2), device=local_rank)
    nvars = N.nvars
    rng = np.random.default_rng(configs.SEED0 + 3)
    ps = oed_inputs(N, to_vec(ps_t), B, rng, I)
    sz = N.out_sizes()
    fields = ("fim", "criteria", "defects_kind")
    nsets = 3   # 3 x (ps 0.1 GB + results 0.15 GB) > L2
    d_ps = [torch.from_numpy(np.ascontiguousarray(np.roll(ps, s, axis=0).T)).to(dev) for s in range(nsets)]
    outs = [{f: torch.empty((sz[f], B), dtype=torch.float64, device=dev) for f in fields} for _ in range(nsets)]
    fsum = torch.zeros(4, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    N.set_stream(stream.cuda_stream)
    want = nat.WANT_FIM | nat.WANT_CRIT | nat.WANT_DEFECTS
    flags = nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA

    def step(s):
        k = s % nsets
        ptrs = {f: outs[k][f].data_ptr() for f in fields}
        ptrs["fim_sum"] = fsum.data_ptr()
        N.eval_raw(d_ps[k].data_ptr(), B, B, want, flags, ptrs, 1)

    ms_total, launches = timed_steps(torch, stream, step, steps, 3, N=N)
    kern_ms = float(np.mean(N.kernel_ms_history(min(256, steps))))
    units = I * B
    e2e_steps = min(steps, 10)
    pin_ps = torch.from_numpy(ps).pin_memory()
    pin_out = {f: torch.empty((B, sz[f]), dtype=torch.float64, pin_memory=True) for f in fields}
    hsum = np.zeros(4)
    hp = {f: pin_out[f].data_ptr() for f in fields}
    hp["fim_sum"] = hsum.ctypes.data

    def host_eval():
        N.eval_raw(pin_ps.data_ptr(), B, nvars, want, 0, hp, 1)

    for _ in range(2):
        host_eval()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_eval()
    dt = (time.perf_counter() - t0) / e2e_steps
    return {"workload": "lotka_oed_64x16k", "units_per_step": units, "value": units * steps / (ms_total * 1e-3), "unit": "units/s",
            "ms_per_step": ms_total / steps, "kernel_ms": kern_ms, "steps": steps, "gpu_launches": launches,
            "outputs": "Fisher information per sample, its sum over the samples, six criteria, shooting defects",
            "e2e": {"value": units / dt, "unit": "units/s", "h2d_bytes_per_step": 8 * B * nvars,
                    "d2h_bytes_per_step": 8 * B * sum(sz[f] for f in fields) + 32, "steps": e2e_steps, "ms_per_step": 1e3 * dt,
                    "api": "cb200_eval, pinned host buffers, candidate-major"}, "_N": N}


# -------------------------------------------------------------------------------------------------- config 1
def run_config1(args, torch, dev, local_rank):
    """single-candidate latency of cb200_eval through host pointers (f, grad f, c, cons_j), the reference's own use"""
    from corleone_b200 import Tsit5, configs, native as nat, to_vec
    spec = configs.config1_lotka_ms24()
    lay, ps_t, st, N = configs.native_from_spec(spec, Tsit5(4), device=local_rank)
    ps = to_vec(ps_t)[None, :].copy()
    sz = N.out_sizes()
    fields = ("jac_vals", "defects_kind", "objective", "grad")
    out = {f: np.empty((1, sz[f])) for f in fields}
    ptrs = {f: out[f].ctypes.data for f in fields}
    want = nat.WANT_JAC | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD
    call = N.prepare_call(ps.ctypes.data, 1, N.nvars, want, 0, ptrs, 3)   # the argument block is built once, as a Julia host would
    for _ in range(50):
        call()
    n = 2000
    t0 = time.perf_counter()
    for _ in range(n):
        call()
    us = (time.perf_counter() - t0) / n * 1e6
    # kernel time alone: the same request with device pointers (no graph), CUDA events around the integration kernel
    d_ps = torch.from_numpy(ps.T.copy()).to(dev)
    d_out = {f: torch.empty((sz[f], 1), dtype=torch.float64, device=dev) for f in fields}
    N.set_stream(torch.cuda.current_stream().cuda_stream)
    for _ in range(20):
        N.eval_raw(d_ps.data_ptr(), 1, 1, want, nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA, {f: d_out[f].data_ptr() for f in fields}, 3)
    torch.cuda.synchronize()
    kern_us = 1e3 * float(np.mean(N.kernel_ms_history(16)))
    units = 24
    return {"workload": "lotka_ms24 (B = 1)", "units_per_step": units, "value": units / (us * 1e-6), "unit": "units/s",
            "latency_us_per_call": us, "ms_per_step": us * 1e-3, "kernel_ms": kern_us * 1e-3,
            "api": "cb200_eval with host pointers, one candidate: f, grad f, c, cons_j; the whole call (H2D, kernels, D2H) is one "
                   "captured CUDA graph",
            "e2e": {"value": units / (us * 1e-6), "unit": "units/s", "h2d_bytes_per_step": 8 * N.nvars,
                    "d2h_bytes_per_step": 8 * sum(sz[f] for f in fields), "ms_per_step": us * 1e-3}, "_N": N, "_spec": spec, "_ps": ps}


# ================================================================================================== CPU arms
def cpu_baseline(spec, ps, K, seconds_target=12.0, nthreads=0, n_intervals=None):
    """The oracle port timed on the host cores on a bounded sample of the same workload (~seconds_target of CPU work)."""
    from oracle import pyoracle as po
    if ORIG_AFFINITY is not None:
        os.sched_setaffinity(0, ORIG_AFFINITY)   # the CPU arm gets every core again, not only the GPU's NUMA node
    O = po.OracleLayer(spec)
    nthreads = nthreads or host_threads()
    O.bench_units(ps[: max(1, min(len(ps), nthreads))], K=K, with_sens=True, nthreads=nthreads)   # warm-up (threads, caches)
    tot_t, tot_u, passes = 0.0, 0, 0
    while tot_t < seconds_target and passes < 1000:
        t, u, _ = O.bench_units(ps, K=K, with_sens=True, nthreads=nthreads)
        tot_t += t
        tot_u += u
        passes += 1
    return (tot_u / tot_t, nthreads,
            f"{passes} passes over {ps.shape[0]} candidates x {O.I} intervals ({tot_u} units), {tot_t:.1f} s")


def dense20_sample(n, seed):
    """host sample of the config-4/5 inputs for the CPU arms: (n, nvars) candidate-major"""
    from corleone_b200 import configs
    from oracle import pyoracle as po
    spec = configs.config4_dense20(256)
    O = po.OracleLayer(spec)
    rng = np.random.default_rng(seed)
    return spec, O, O.default_ps()[None, :] + 0.1 * rng.standard_normal((n, O.nvars))


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The real reference is
    Julia (absent here), so this is the oracle port with all host threads (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from corleone_b200 import configs
    nthreads = host_threads()
    n_gpu_candidates = 65536 if args.gpus == 1 else 4096
    workload = "dense20_256x65k" if args.gpus == 1 else "scale_1M"
    nb = max(nthreads, 16)
    spec, O, ps = dense20_sample(nb, configs.SEED0 + 4)
    t, u, _ = O.bench_units(ps, K=2, nthreads=nthreads)
    budget = min(10.0, 120.0 / max(1, args.steps + args.warmup))
    passes = max(1, int(budget / max(t, 1e-6)))

    def one_step():
        tt, uu = 0.0, 0
        for _ in range(passes):
            t1, u1, _ = O.bench_units(ps, K=2, nthreads=nthreads)
            tt += t1
            uu += u1
        return tt, uu

    for _ in range(args.warmup):
        one_step()
    tot_t, tot_u = 0.0, 0
    for _ in range(args.steps):
        t, u = one_step()
        tot_t += t
        tot_u += u
    value = tot_u / tot_t
    sample = (f"{passes} passes over {nb} candidates x 256 intervals per step (of {n_gpu_candidates} candidates in the "
              f"{args.gpus}-GPU job)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "units/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dense20_config(workload, n_gpu_candidates, args.gpus),   # identical to our arm's `config`
        "cpu_baseline": {"value": value, "unit": "units/s", "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "units/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def dense20_config(workload, candidates, world, **extra):
    c = {"workload": workload, "model": "dense20: x' = W x - x^3 + b u (synthetic, BASELINE.md section 5)", "intervals": 256,
         "candidates": candidates, "method": "tsit5", "steps_per_piece": 2, "pieces_per_interval": 4,
         "sens_directions": 24, "units_per_step": 256 * candidates,
         "l2": "inputs larger than L2 (candidate matrix 3.2 GB at 65 536 candidates, 200 MB at 4096); results stream to HBM",
         "parallelism": ("single GPU" if world == 1 else
                         f"candidate axis sharded x{world} (strong scaling), per step allreduce(objective_sum, grad_sum) + "
                         "allgather(defects) inside the timed region")}
    c.update(extra)
    return c


# ================================================================================================== N > 1: config 5
def run_scale(args, torch, dist, dev, world, rank, local_rank):
    from corleone_b200 import Tsit5, configs, native as nat, to_vec
    I, B_total = 256, 4096
    assert B_total % world == 0
    B = B_total // world
    b0 = rank * B
    spec = configs.config4_dense20(I)
    lay, ps_t, st, N = configs.native_from_spec(spec, Tsit5(2), device=local_rank)
    nvars = N.nvars
    stream = torch.cuda.current_stream()
    N.set_stream(stream.cuda_stream)
    # communicator: rank 0 makes the NCCL id, the host (torch.distributed here, Distributed.jl / MPI for a Julia host)
    # hands it round; from then on the library runs its own data plane
    ident = [nat.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    N.comm_init(world, rank, ident[0], max_B_total=B_total)
    info = N.comm_info()
    sz = N.out_sizes()
    # all ranks draw the whole batch from one seed and keep their slice: identical to the 1-GPU job
    full = dense20_inputs_soa(torch, dev, to_vec(ps_t), B_total, configs.SEED0 + 5)
    d_ps = full[:, b0:b0 + B].contiguous()
    del full
    d_obj = torch.empty((1, B), dtype=torch.float64, device=dev)
    d_xend = torch.empty((sz["xend"], B), dtype=torch.float64, device=dev)
    d_sums = torch.zeros(1 + nvars, dtype=torch.float64, device=dev)
    want = nat.WANT_SENS | nat.WANT_DEFECTS | nat.WANT_XEND | nat.WANT_OBJ | nat.WANT_GRAD
    flags = nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA | nat.COMM_REDUCE | nat.COMM_GATHER
    ptrs = {"objective": d_obj.data_ptr(), "xend": d_xend.data_ptr(), "objective_sum": d_sums.data_ptr(),
            "grad_sum": d_sums.data_ptr() + 8}

    def step(s):
        N.eval_raw(d_ps.data_ptr(), B, B, want, flags, ptrs, 1, resident=nat.WANT_SENS, comm_B_total=B_total, comm_b0=b0)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    with ClockSampler(local_rank) as clk:
        clk.wait_first()
        step(0)
        barrier()
        clk.mark()
        ms_total, launches = timed_steps(torch, stream, step, args.steps, max(3, args.warmup), barrier, N=N)
        kern_ms = float(np.mean(N.kernel_ms_history(min(256, args.steps))))
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        # the timed region is tens of milliseconds at N = 8: keep the same workload running (untimed, the same number of
        # steps on every rank) until nvidia-smi has sampled the clocks under it
        for _ in range(max(0, int(np.ceil(400.0 / max(ms_total / args.steps, 1e-3))) - args.steps)):
            step(0)
        barrier()
        clocks = clk.summary()
    units = I * B_total
    value = units * args.steps / (ms_total * 1e-3)
    # check of the collectives on the timed workload: the gathered defects and the reduced sums are the same on every rank
    gp, gld = N.comm_gathered()
    chk = torch.tensor([float(d_sums[0].item())], dtype=torch.float64, device=dev)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    sums_identical = bool(lo.item() == hi.item())

    # ---- e2e: every rank feeds its shard from pinned host memory; its own defects, end states and the all-reduced sums
    # come back to the host; the all-gathered defect vector and the Jacobian stay on the device
    e2e_steps = max(1, min(args.steps, args.e2e_steps or 5))
    pin_ps = torch.empty((B, nvars), dtype=torch.float64, pin_memory=True)
    pin_ps.copy_(d_ps.t())
    pin_dk = torch.empty((B, sz["defects_kind"]), dtype=torch.float64, pin_memory=True)
    pin_x = torch.empty((B, sz["xend"]), dtype=torch.float64, pin_memory=True)
    pin_obj = torch.empty((B, 1), dtype=torch.float64, pin_memory=True)
    h_sums = np.zeros(1 + nvars)
    hp = {"defects_kind": pin_dk.data_ptr(), "xend": pin_x.data_ptr(), "objective": pin_obj.data_ptr(),
          "objective_sum": h_sums.ctypes.data, "grad_sum": h_sums.ctypes.data + 8}
    hflags = nat.COMM_REDUCE | nat.COMM_GATHER

    def host_eval():
        N.eval_raw(pin_ps.data_ptr(), B, nvars, want, hflags, hp, 1, resident=nat.WANT_SENS, comm_B_total=B_total, comm_b0=b0)

    host_eval()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_eval()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    e2e = {"value": units * e2e_steps / dt, "unit": "units/s", "h2d_bytes_per_step": 8 * B * nvars * world,
           "d2h_bytes_per_step": 8 * (B * (sz["defects_kind"] + sz["xend"] + 1) + 1 + nvars) * world, "steps": e2e_steps,
           "ms_per_step": 1e3 * dt / e2e_steps,
           "api": "cb200_eval with pinned host buffers per rank + CB200_COMM_REDUCE | CB200_COMM_GATHER; every rank's own "
                  "defects / end states / objectives and the all-reduced sums return to the host, the all-gathered defect "
                  "vector and the Jacobian stay on the device (bytes are whole-job totals)"}
    barrier()
    N.comm_destroy()
    secondary = None
    if not args.no_secondary:
        # round 1's workload for comparison: config 2, weak scaling, 4096 candidates per rank, sums all-reduced
        secondary = run_config2_weak(args, torch, dist, dev, world, rank, local_rank)
    if rank == 0:
        peak, peak_src = fp64_peak(nat, local_rank)
        hbm_peak, hbm_src = load_peaks()
        tmap = {0: "none", 1: "nccl (ncclAllReduce + ncclAllGather after the kernels)",
                2: "p2p (defects stored into every rank's window by the shooting kernel, one-shot all-reduce in the tail of the "
                   "objective kernel; no NCCL call per step)"}
        line = {
            "metric": METRIC, "value": value, "unit": "units/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dense20_config("scale_1M", B_total, world),
            "config_details": {"candidates_per_gpu": B, "transport": tmap[info["transport"]],
                               "collectives_checked": {"reduced_sums_identical_on_all_ranks": sums_identical,
                                                       "gather_leading_dimension": gld}},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roofline_entry(WORK["dense20"], I * B, kern_ms, peak, peak_src, hbm_peak, hbm_src,
                                       "shoot_dense_mma_queue_kernel<Tsit5,uniform,16 warps> (FP64 tensor path) incl. fused defect all-gather "
                                       "stores, per rank", "dense20"),
            "secondary": secondary,
        }
        print(json.dumps(line))


def run_config2_weak(args, torch, dist, dev, world, rank, local_rank):
    """round 1's workload: config 2, 4096 candidates per GPU, sums all-reduced by the fused tail.  A 0.12 ms evaluation is
    short against the exchange's synchronisation (7-20 us: flags over NVLink + the slowest of N GPUs), so the batch consumer
    keeps TWO evaluations in flight -- two handles of the layer on two streams, each with its own communicator: the tail of
    evaluation k waits for its peers while the shooting kernel of evaluation k+1 runs"""
    from corleone_b200 import Tsit5, configs, native as nat, to_vec
    I, B, K = 100, 4096, 2
    spec = configs.config2_lqr()
    rng = np.random.default_rng(configs.SEED0 + 2 + 1000 * rank)
    streams = [torch.cuda.current_stream(), torch.cuda.Stream(device=dev)]
    Ns = []
    for q in range(2):
        lay, ps_t, st, Nq = configs.native_from_spec(spec, Tsit5(K), device=local_rank)
        Nq.set_stream(streams[q].cuda_stream)
        ident = [nat.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        Nq.comm_init(world, rank, ident[0], max_B_total=0)
        Ns.append(Nq)
    N = Ns[0]
    nvars = N.nvars
    ps_host = lqr_inputs(nvars, to_vec(ps_t), B, rng)
    stream = streams[0]
    sz = N.out_sizes()
    fields = ("jac_vals", "defects_kind", "objective", "grad_compact")
    sz["grad_compact"] = N.grad_structure(2)[1]
    nsets = 6
    d_ps = [torch.from_numpy(np.ascontiguousarray(np.roll(ps_host, s, axis=0).T)).to(dev) for s in range(nsets)]
    d_out = [{f: torch.empty((sz[f], B), dtype=torch.float64, device=dev) for f in fields} for _ in range(nsets)]
    d_sums = [torch.zeros(nvars + 1, dtype=torch.float64, device=dev) for _ in range(nsets)]
    want = nat.WANT_JAC | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD | nat.WANT_GRAD_COMPACT
    flags = nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA | nat.COMM_REDUCE

    def step(s):
        k = s % nsets
        ptrs = {f: d_out[k][f].data_ptr() for f in fields}
        ptrs["objective_sum"] = d_sums[k].data_ptr()
        ptrs["grad_sum"] = d_sums[k].data_ptr() + 8
        Ns[s % 2].eval_raw(d_ps[k].data_ptr(), B, B, want, flags, ptrs, 2)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    steps = max(20, args.steps)
    steps += steps % 2
    for s_ in range(6):
        step(s_)
    barrier()
    e0, e1, ej = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(streams[0])
    streams[1].wait_event(e0)
    for s_ in range(steps):
        step(s_)
    ej.record(streams[1])
    streams[0].wait_event(ej)
    e1.record(streams[0])
    barrier()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    # the sums of the last evaluations are the same on every rank
    chk = d_sums[(steps - 1) % nsets][:4].clone()
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    sums_identical = bool(torch.equal(lo, hi))
    barrier()
    # ---- e2e of the consumer that needs only the sums across all ranks (an expectation-type objective and its gradient):
    # every rank feeds its 4096 candidates from pinned host memory, f / c / cons_j / the per-candidate gradients stay in HBM
    # (cb200_out.resident), the all-reduced sums (6.4 KB) come back to the host
    pin_ps = torch.from_numpy(ps_host).pin_memory()
    h_sums = np.zeros(1 + nvars)
    hp = {"objective_sum": h_sums.ctypes.data, "grad_sum": h_sums.ctypes.data + 8}
    res_bits = nat.WANT_JAC | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD_COMPACT | nat.WANT_GRAD
    e2e_steps = 10

    def host_eval():
        N.eval_raw(pin_ps.data_ptr(), B, nvars, want, nat.COMM_REDUCE, hp, 2, resident=res_bits)

    for _ in range(3):
        host_eval()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_eval()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    barrier()
    for Nq in Ns:
        Nq.comm_destroy()
    return {"workload": "lqr_ms100_b4096 per GPU (weak scaling, round 1's bench)", "value": I * B * world * steps / (ms_total * 1e-3),
            "unit": "units/s", "ms_per_step": ms_total / steps, "steps": steps,
            "collective": "all-reduce of (objective_sum, grad_sum) fused into the tail of the objective kernel; two evaluations in "
                          "flight (two handles on two streams): the exchange of one overlaps the shooting kernel of the next",
            "reduced_sums_identical_on_all_ranks": sums_identical,
            "e2e_sums_only": {"value": I * B * world * e2e_steps / dt, "unit": "units/s", "ms_per_step": 1e3 * dt / e2e_steps,
                              "h2d_bytes_per_step": 8 * B * nvars * world, "d2h_bytes_per_step": 8 * (1 + nvars) * world,
                              "api": "cb200_eval with pinned host candidates per rank + CB200_COMM_REDUCE; every per-candidate "
                                     "result stays device-resident, the all-reduced sums return to the host"}}


# ================================================================================================== main
def strip(d):
    return {k: v for k, v in d.items() if not k.startswith("_")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--only", default="", help="N = 1: comma list of configs to run (1,2,3,4); default all")
    ap.add_argument("--candidates", type=int, default=65536, help="N = 1: candidates of the config-4 run (tuning; 65536 = the config)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from corleone_b200 import native as nat

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        run_scale(args, torch, dist, dev, world, rank, local_rank)
        dist.barrier()
        dist.destroy_process_group()
        return

    only = set(args.only.split(",")) if args.only else {"1", "2", "3", "4"}
    peak, peak_src = fp64_peak(nat, local_rank)
    hbm_peak, hbm_src = load_peaks()
    all_configs = []
    line = {}
    with ClockSampler(local_rank) as clk:
        clk.wait_first()
        if "4" in only:
            c4 = run_config4(args, torch, dev, local_rank, clk, args.candidates)
            clocks = clk.summary()
            units = c4["units_per_step"]
            c4["roofline"] = roofline_entry(WORK["dense20"], units, c4["kernel_ms"], peak, peak_src, hbm_peak, hbm_src,
                                            "shoot_dense_mma_queue_kernel<Tsit5,uniform,16 warps> (FP64 tensor path, mma.sync.m8n8k4.f64; persistent blocks, tile queue) "
                                            "incl. fused defect epilogue", "dense20")
            line = {"metric": METRIC, "value": c4["value"], "unit": "units/s", "n_gpus": 1, "steps": args.steps,
                    "warmup": max(3, args.warmup), "ms_per_step": c4["ms_per_step"], "higher_is_better": True,
                    "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": dense20_config(c4["workload"], args.candidates, 1),
                    "config_details": {"chunking": c4["chunking"], "outputs": c4["outputs"]},
                    "e2e": c4["e2e"], "gpu_launches": int(c4["_launches_timed"]),
                    "clocks": clocks, "roofline": c4["roofline"]}
            if not args.no_cpu_baseline:
                from corleone_b200 import configs
                spec, O, ps = dense20_sample(max(host_threads(), 16), configs.SEED0 + 4)
                v, cores, sample = cpu_baseline(spec, ps, 2)
                line["cpu_baseline"] = {"value": v, "unit": "units/s", "cores": cores, "kind": "port", "sample": sample}
            c4["_N"].close()
            all_configs.append(strip(c4))
            torch.cuda.empty_cache()
            # the N = 1 base of the strong-scaling sweep (`--gpus N`, N > 1): config 5's 2^20 units on one GPU, no collectives
            c5 = run_config4(args, torch, dev, local_rank, clk, 4096, e2e=False)
            c5["workload"] = "scale_1M on one GPU (256 intervals x 4096 candidates: the base of the N > 1 strong-scaling runs)"
            c5["roofline"] = roofline_entry(WORK["dense20"], c5["units_per_step"], c5["kernel_ms"], peak, peak_src, hbm_peak, hbm_src,
                                            c4["roofline"]["kernel"], "dense20")
            c5["_N"].close()
            all_configs.append(strip(c5))
            torch.cuda.empty_cache()
        sub_steps = max(20, args.steps)
        if "2" in only:
            c2 = run_config2(args, torch, dev, local_rank, sub_steps)
            c2["roofline"] = roofline_entry(WORK["lqr"], c2["units_per_step"], c2["kernel_ms"], peak, peak_src, hbm_peak, hbm_src,
                                            "shoot_kernel<Lqr,G=4,Tsit5,uniform,4 blocks/SM> incl. fused defect/gradient/cons_j epilogue",
                                            "lqr")
            if not args.no_cpu_baseline:
                v, cores, sample = cpu_baseline(c2["_spec"], c2["_ps_host"][:1024], 2, seconds_target=4.0)
                c2["cpu_baseline"] = {"value": v, "unit": "units/s", "cores": cores, "kind": "port", "sample": sample}
            c2["_N"].close()
            all_configs.append(strip(c2))
        if "3" in only:
            c3 = run_config3(args, torch, dev, local_rank, sub_steps)
            c3["roofline"] = roofline_entry(WORK["lotka_oed"], c3["units_per_step"], c3["kernel_ms"], peak, peak_src, hbm_peak,
                                            hbm_src, "shoot_kernel<OedAug<Lotka2>,G=0,Tsit5,uniform> incl. fused Fisher read-out, "
                                            "criteria and defect epilogue", "lotka_oed")
            c3["_N"].close()
            all_configs.append(strip(c3))
        if "1" in only:
            c1 = run_config1(args, torch, dev, local_rank)
            c1["roofline"] = roofline_entry(WORK["lotka_ms24"], c1["units_per_step"], c1["kernel_ms"], peak, peak_src, hbm_peak,
                                            hbm_src, "shoot_kernel<Lotka3,G=2,Tsit5,uniform>, flat launch (latency-bound: 120 threads)",
                                            "lotka_ms24")
            if not args.no_cpu_baseline:
                from oracle import pyoracle as po
                O = po.OracleLayer(c1["_spec"])
                t0 = time.perf_counter()
                for _ in range(200):
                    O.bench_units(c1["_ps"], K=4, with_sens=True, nthreads=1)
                c1["cpu_baseline"] = {"value": 24 / ((time.perf_counter() - t0) / 200), "unit": "units/s", "cores": 1, "kind": "port",
                                      "sample": "200 single-candidate evaluations, 1 thread (the reference default is serial)"}
            c1["_N"].close()
            all_configs.append(strip(c1))
    if not line:   # tuning runs without config 4: the first config that ran is the line
        first = all_configs[0]
        line = {"metric": METRIC, "value": first["value"], "unit": "units/s", "n_gpus": 1, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": first["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": first["workload"]},
                "e2e": first.get("e2e"), "roofline": first.get("roofline"), "gpu_launches": None}
    line["all_configs"] = all_configs
    print(json.dumps(line))


if __name__ == "__main__":
    main()
