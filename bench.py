#!/usr/bin/env python
"""bench.py -- shooting-interval solves + sensitivities per second (FP64) on B200.

    python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   the restated reference CPU path on the host cores

Workload (N=1): BASELINE.json configs[1] = `lqr_ms100_b4096`: linear-quadratic regulator, 100 shooting
intervals x 4096 candidates, Tsit5 with 2 fixed steps per control piece (M = 4 pieces per interval), full
forward sensitivities (ns = 8 directions per interval).  One unit = one (interval, candidate) system.
A step = one pass of the hot path over the batch: integration + sensitivities of every unit, delivered as what
the reference's NLP callbacks consume (src/dynprob.jl:68-73,105-131,143-164): objective f, gradient grad f,
shooting constraints c (kind-major, the order dynprob uses) and the nonzeros of the constraint Jacobian cons_j
(grad f without its structural zeros: the objective state of the last interval depends on that interval's block only).  Weak scaling: every rank owns B candidates;
only the cross-candidate sums (objective, gradient) are all-reduced (DESIGN.md "multi-GPU").
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

WORKLOAD = "lqr_ms100_b4096"
I_INTERVALS, B_PER_GPU, K_STEPS_PER_PIECE, M_PIECES, NS_DIRS, NX = 100, 4096, 2, 4, 8, 2
OBJECTIVE_ROW = 2   # x2 = accumulated cost (examples/linear_quadratic/main.jl:49, loss = :x2)

# Algorithmic work per unit (DESIGN.md "roofline"):
#   E = M (6K + 1) RHS evaluations (Tsit5, FSAL re-seeded at every piece start)
#   W = E (C_f + C_JS) + M K 42 nz,  nz = nx (1 + ns);   LQR: C_f = 10, C_JS = 2 nnz(J) ns + nnz(f_p,f_u) = 36
E_EVALS = M_PIECES * (6 * K_STEPS_PER_PIECE + 1)
NZ = NX * (1 + NS_DIRS)
FLOPS_PER_UNIT = E_EVALS * (10 + 36) + M_PIECES * K_STEPS_PER_PIECE * 42 * NZ      # 8440
#   Q = 8 (nz_in + np + nc_i + nz_out + n_defect + n_obj) bytes of compulsory HBM traffic (BASELINE.md formula)
BYTES_PER_UNIT = 8 * (NX + 2 + M_PIECES + NZ + 4 + 0)                                # 240
NCU_DRAM_BYTES_PER_LAUNCH = 26944000 + 17881344   # profiles/r1_shoot_lqr_G4_fused.txt


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
    allocated: the e2e arm moves 117 MB per step over PCIe, and first-touch places the buffers on the local node."""
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
        return d.get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


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
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
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


def make_inputs(nvars, ps0, B, rng):
    """candidate matrix (B, nvars): nodes x ~ U[0,4], cost node 0, controls N(0,1)"""
    ps = np.tile(ps0, (B, 1))
    # layout per interval (u0 (0 or 2), p (2), controls (4)); interval 1 has no u0
    off = 0
    for i in range(I_INTERVALS):
        nu0 = 0 if i == 0 else 2
        if nu0:
            ps[:, off] = rng.uniform(0, 4, B)
            ps[:, off + 1] = 0.0
        ps[:, off + nu0 + 2: off + nu0 + 6] = rng.standard_normal((B, 4))
        off += nu0 + 6
    assert off == nvars
    return ps


def cpu_baseline(spec, ps, seconds_target=12.0, nthreads=0):
    """The oracle port timed on the host cores on a bounded sample of the same workload (~seconds_target of
    CPU work: whole passes over the batch, repeated)."""
    from oracle import pyoracle as po
    if ORIG_AFFINITY is not None:
        os.sched_setaffinity(0, ORIG_AFFINITY)   # the CPU arm gets every core again, not only the GPU's NUMA node
    O = po.OracleLayer(spec)
    nthreads = nthreads or host_threads()
    O.bench_units(ps[:64], K=K_STEPS_PER_PIECE, with_sens=True, nthreads=nthreads)          # warm-up (threads, caches)
    tot_t, tot_u, passes = 0.0, 0, 0
    while tot_t < seconds_target and passes < 1000:
        t, u, _ = O.bench_units(ps, K=K_STEPS_PER_PIECE, with_sens=True, nthreads=nthreads)
        tot_t += t
        tot_u += u
        passes += 1
    return (tot_u / tot_t, nthreads,
            f"{passes} passes over {ps.shape[0]} candidates x {I_INTERVALS} intervals ({tot_u} units), {tot_t:.1f} s")


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The real
    reference is Julia (absent here), so this is the oracle port with all host threads (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from corleone_b200 import configs
    from oracle import pyoracle as po
    spec = configs.config2_lqr()
    O = po.OracleLayer(spec)
    rng = np.random.default_rng(configs.SEED0 + 2)
    B = B_PER_GPU * args.gpus
    nthreads = host_threads()
    # each step = a bounded sample of the workload (whole passes over one rank's batch), sized so that the run
    # ends within a few minutes
    ps = make_inputs(O.nvars, O.default_ps(), B_PER_GPU, rng)
    t, u, _ = O.bench_units(ps, K=K_STEPS_PER_PIECE, nthreads=nthreads)
    budget = min(10.0, 120.0 / max(1, args.steps + args.warmup))
    passes = max(1, int(budget / max(t, 1e-6)))
    nb = B_PER_GPU

    def one_step():
        tt, uu = 0.0, 0
        for _ in range(passes):
            t1, u1, _ = O.bench_units(ps, K=K_STEPS_PER_PIECE, nthreads=nthreads)
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
    sample = f"{passes} passes over {nb} candidates x {I_INTERVALS} intervals per step (of {B} in the {args.gpus}-GPU job)"
    print(json.dumps({
        "impl": "reference", "metric": "shooting-interval solves+sensitivities/sec (FP64)", "value": value,
        "unit": "units/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "intervals": I_INTERVALS, "candidates_per_gpu": B_PER_GPU,
                   "method": "tsit5", "steps_per_piece": K_STEPS_PER_PIECE, "pieces_per_interval": M_PIECES,
                   "sens_directions": NS_DIRS, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "units/s", "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "units/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="default: min(steps, 20)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from corleone_b200 import Tsit5, configs, native as nat

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    spec = configs.config2_lqr()
    lay, ps_t, st, N = configs.native_from_spec(spec, Tsit5(K_STEPS_PER_PIECE), device=local_rank)
    from corleone_b200 import to_vec
    B = B_PER_GPU
    nvars = N.nvars
    rng = np.random.default_rng(configs.SEED0 + 2 + 1000 * rank)
    ps_host = make_inputs(nvars, to_vec(ps_t), B, rng)                      # (B, nvars) candidate-major
    stream = torch.cuda.current_stream()
    N.set_stream(stream.cuda_stream)

    # ---- device-resident arm: rotate over buffer sets whose total footprint exceeds L2 (126 MB)
    sz = N.out_sizes()
    fields = ("jac_vals", "defects_kind", "objective", "grad_compact")   # cons_j, c, f, grad f (structural zeros omitted)
    sz["grad_compact"] = N.grad_structure(OBJECTIVE_ROW)[1]
    set_bytes = 8 * B * (nvars + sum(sz[f] for f in fields))
    nsets = max(2, int(np.ceil(3 * 126e6 / set_bytes)))
    d_ps, d_out, d_sums = [], [], []
    for s in range(nsets):
        d_ps.append(torch.from_numpy(np.ascontiguousarray(np.roll(ps_host, s, axis=0).T)).to(dev))   # [nvars, B] SoA
        d_out.append({f: torch.empty((sz[f], B), dtype=torch.float64, device=dev) for f in fields})
        d_sums.append(torch.zeros(nvars + 1, dtype=torch.float64, device=dev))
    # WANT_GRAD stays set for the cross-candidate gradient sum (grad_sum); the per-candidate gradient is delivered compact
    want = nat.WANT_JAC | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD | nat.WANT_GRAD_COMPACT
    flags = nat.MEM_DEVICE | nat.PS_SOA | nat.OUT_SOA

    def step(s, comm=True):
        k = s % nsets
        ptrs = {f: d_out[k][f].data_ptr() for f in fields}
        ptrs["objective_sum"] = d_sums[k].data_ptr()
        ptrs["grad_sum"] = d_sums[k].data_ptr() + 8
        if world > 1 and sum_done[k] is not None:
            stream.wait_event(sum_done[k])        # buffer set k: its previous all-reduce must have finished
        N.eval_raw(d_ps[k].data_ptr(), B, B, want, flags, ptrs, OBJECTIVE_ROW)
        if world > 1 and comm:
            # the path's only exchange: cross-candidate partial sums (1 + nvars doubles), all-reduced on a side
            # stream so that the next evaluation (another buffer set) overlaps the NCCL latency
            ev = torch.cuda.Event()
            ev.record(stream)
            comm_stream.wait_event(ev)
            with torch.cuda.stream(comm_stream):
                dist.all_reduce(d_sums[k])
                sum_done[k] = torch.cuda.Event()
                sum_done[k].record(comm_stream)

    comm_stream = torch.cuda.Stream() if world > 1 else None
    sum_done = [None] * nsets

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(max(3, args.warmup)):
        step(s)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        clk.wait_first()
        # put the GPU under the same load for ~0.7 s before the timed region so that the sampler (100 ms
        # period) sees the clocks this workload actually runs at; these steps are extra warm-up
        t_end, s2 = time.time() + 0.7, 0
        clk.mark()
        while time.time() < t_end:   # wall-clock bounded, so no collective here: ranks run different counts
            step(s2, comm=False)
            s2 += 1
            if s2 % 64 == 0:
                torch.cuda.synchronize()
        barrier()
        launches0 = N.launch_count()
        ev0.record(stream)
        for s in range(args.steps):
            step(s)
        if world > 1:
            stream.wait_stream(comm_stream)   # the timed region ends when the last all-reduce has finished
        ev1.record(stream)
        barrier()
        ms_total = ev0.elapsed_time(ev1)
        launches_timed = N.launch_count() - launches0
        # kernel durations of the timed steps (event pairs the library records around the integration kernel)
        hist = N.kernel_ms_history(min(256, args.steps))
        kern_ms = float(np.mean(hist)) if len(hist) else float("nan")
        t_end = time.time() + 0.3
        while time.time() < t_end:
            step(0, comm=False)
        torch.cuda.synchronize()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    units_per_step = I_INTERVALS * B * world
    value = units_per_step * args.steps / (ms_total * 1e-3)

    # ---- end-to-end arm: the public host API (Julia-style candidate-major matrices in pinned host memory)
    e2e_steps = args.e2e_steps or min(args.steps, 20)
    pin_ps = torch.from_numpy(ps_host).pin_memory()
    pin_out = {f: torch.empty((B, sz[f]), dtype=torch.float64).pin_memory() for f in fields}
    pinned = {f: v.numpy() for f, v in pin_out.items()}
    ps_np = pin_ps.numpy()
    host_ptrs = {f: pinned[f].ctypes.data for f in fields}

    def host_eval():   # cb200_eval on host pointers: returns when the results are in the caller's buffers
        N.eval_raw(ps_np.ctypes.data, B, nvars, want, 0, host_ptrs, OBJECTIVE_ROW)

    for _ in range(3):
        host_eval()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_eval()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = units_per_step * e2e_steps / float(t.item())
    # context for e2e: the same bytes as plain pinned copies, both directions concurrently on two streams
    d_in = torch.empty(B * nvars, dtype=torch.float64, device=dev)
    d_res = torch.empty(B * sum(sz[f] for f in fields), dtype=torch.float64, device=dev)
    h_res = torch.empty(d_res.shape, dtype=torch.float64).pin_memory()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s_in):
            d_in.copy_(pin_ps.view(-1), non_blocking=True)
        with torch.cuda.stream(s_out):
            h_res.copy_(d_res, non_blocking=True)
    torch.cuda.synchronize()
    copy_ms = (time.perf_counter() - t0) / 5 * 1e3
    del d_in, d_res, h_res
    h2d = 8 * B * nvars
    d2h = 8 * B * sum(sz[f] for f in fields)

    if rank == 0:
        hbm_peak, hbm_src = load_peaks()
        # FP64 roof: DFMA and DMMA share one datapath on B200 (tools/dmma_micro.cu); the higher probe is the peak
        peak_dfma = nat.probe_fp64_tflops(local_rank, 40000)
        peak_dmma = nat.probe_fp64_mma_tflops(local_rank, 20000)
        fp64_peak = max(peak_dfma, peak_dmma)
        ach_tflops = FLOPS_PER_UNIT * I_INTERVALS * B / (kern_ms * 1e-3) / 1e12
        ach_gbs = BYTES_PER_UNIT * I_INTERVALS * B / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": "shooting-interval solves+sensitivities/sec (FP64)", "value": value, "unit": "units/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "intervals": I_INTERVALS, "candidates_per_gpu": B,
                       "method": "tsit5", "steps_per_piece": K_STEPS_PER_PIECE, "pieces_per_interval": M_PIECES,
                       "sens_directions": NS_DIRS, "units_per_step": units_per_step,
                       "l2": f"inputs larger than L2: {nsets} rotating buffer sets of {set_bytes / 1e6:.0f} MB",
                       "parallelism": (f"batch-sharded x{world}, NCCL allreduce(objective_sum, grad_sum) per step on a side stream"
                                       if world > 1 else "single GPU")},
            "e2e": {"value": e2e_value, "unit": "units/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": 1e3 * units_per_step / e2e_value,
                    "pcie_copy_only_ms": copy_ms,
                    "api": "cb200_eval with pinned host buffers, candidate-major (Julia layout), 8 chunks on 3 streams"},
            "gpu_launches": int(launches_timed),
            "clocks": clk.summary(),
            "roofline": {"bound": "fp64", "achieved": ach_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": ach_tflops / fp64_peak if fp64_peak > 0 else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture of this
                         # command (profiles/r1_shoot_lqr_G4_fused.txt): 26.8 MB read + 15.3 MB written; the rest of
                         # the 65 MB of results is still in the 126 MB write-back L2 when the kernel ends
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "traffic_unit": "bytes per launch",
                         "algorithmic_bytes_per_launch": BYTES_PER_UNIT * I_INTERVALS * B,
                         "kernel": "shoot_kernel<Lqr,G=4,Tsit5,uniform,4 blocks/SM> incl. fused defect/gradient/cons_j epilogue",
                         "kernel_ms": kern_ms,
                         "flops_per_unit": FLOPS_PER_UNIT, "bytes_per_unit": BYTES_PER_UNIT,
                         "peak_source": "FP64 probes measured in this run, max of DFMA chains "
                                        f"({peak_dfma:.2f}) and mma.sync.m8n8k4.f64 ({peak_dmma:.2f}) TFLOP/s; "
                                        "MEASURED_PEAKS.json has no FP64 figure",
                         "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                                 "peak_source": f"MEASURED_PEAKS.json ({hbm_src})"}},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_baseline(spec, ps_host)
            line["cpu_baseline"] = {"value": v, "unit": "units/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
