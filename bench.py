#!/usr/bin/env python
"""bench.py -- the hot path of BASELINE.json on N GPUs of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference            # the reference's own CPU path on the host cores

Workload (BASELINE.json configs[1]): batched forward + inverse negacyclic NTT,
N = 2^16, 55-bit prime (GeneratePrimes(1, 55, true, N)), 8192 polynomials per GPU,
coefficients uniform in [0, q).  A "step" is one forward pass (x -> y) and one
inverse pass (y -> z) over the batch = 2 * 8192 transforms per GPU.  Batches of
independent polynomials shard across GPUs with no data-path collective
(scaling = weak: 8192 polynomials per GPU); the only torch.distributed calls
are the barrier and the max-over-ranks of the measured time.

`value`  : NTTs/sec, whole job, inputs resident in HBM, CUDA-event timed.
`e2e`    : the same metric through the C ABI's HOST-pointer path (pinned host
           buffers in, pinned host buffers out, copies inside the timed region).
`roofline`: the forward transform (all kernels of one hexl_b200_ntt_forward call)
           against the measured HBM copy bandwidth; 16*N algorithmic bytes/NTT.
`cpu_baseline`: the compiled reference (oracle/_ref) timed on the host cores
           (rank 0, N = 1 only), on a bounded sample of the same workload.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N = 16
MOD_BITS = 55
BATCH = 8192
METRIC = "ntt_per_sec_fwd_inv_N65536_q55_batch8192"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="polynomials per GPU")
    ap.add_argument("--logn", type=int, default=LOG_N)
    ap.add_argument("--bits", type=int, default=MOD_BITS)
    ap.add_argument("--e2e-batch", type=int, default=None, help="polynomials per GPU for the host-pointer leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-eltwise", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------ reference arm
def cpu_threads() -> int:
    """host threads this process can actually run concurrently: the affinity mask
    capped by the cgroup CPU quota (a container may see 128 CPUs but own 16)."""
    try:
        n = max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        n = os.cpu_count() or 1
    quota = None
    try:  # cgroup v2
        q, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            quota = int(q) / int(period)
    except (OSError, ValueError):
        try:  # cgroup v1
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / period
        except (OSError, ValueError):
            pass
    if quota:
        n = max(1, min(n, int(quota + 0.999)))
    return n


def cpu_leg(n, q, threads, polys, reps):
    """forward + inverse over `polys` polynomials with the reference's CPU path;
    returns (NTTs/sec best-of-reps, description).  Uses oracle/_ref when it is
    here, else the C restatement."""
    import numpy as np
    import oracle
    chk = oracle.best_checker()
    rng = np.random.default_rng(42)
    x = rng.integers(0, q, size=n * polys, dtype=np.uint64)
    if chk.kind == "reference":
        y, z = np.empty_like(x), np.empty_like(x)
        fwd = lambda: chk.ntt_forward(x, n, q, 1, 1, threads=threads, out=y)
        inv = lambda: chk.ntt_inverse(y, n, q, 1, 1, threads=threads, out=z)
        tier = chk.tier(q)
    else:
        chk.tables(n, q)
        state = {}
        fwd = lambda: state.__setitem__("y", chk.ntt_forward(x, n, q, 1, 1, threads=threads))
        inv = lambda: state.__setitem__("z", chk.ntt_inverse(state["y"], n, q, 1, 1, threads=threads))
        tier = "scalar C restatement"
    fwd(); inv()  # warm: tables, page faults
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fwd(); inv()
        best = min(best, time.perf_counter() - t0)
    if chk.kind == "reference":
        assert (z == x).all(), "reference round trip failed"
    return 2 * polys / best, chk.kind, tier


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    n = 1 << args.logn
    q = oracle.best_checker().generate_primes(1, args.bits, True, n)[0]
    threads = cpu_threads()
    polys = max(threads * 16, 64)
    # each step = one bounded sample (forward + inverse over `polys` polynomials)
    vals = []
    kind = tier = None
    for _ in range(args.warmup + args.steps):
        v, kind, tier = cpu_leg(n, q, threads, polys, 1)
        vals.append(v)
    vals = vals[args.warmup:]
    value = 2 * polys * len(vals) / sum(2 * polys / v for v in vals)
    ms = 1e3 * 2 * polys / value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "NTT/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"batched Fwd+Inv NTT, N=2^{args.logn}, {args.bits}-bit prime q={q}, "
                               f"sample of {polys} polynomials per step on the host cores",
                   "tier": tier},
        "cpu_baseline": {"value": value, "unit": "NTT/s", "cores": threads, "kind": kind,
                         "sample": f"{polys} polynomials x (forward + inverse) per step, {threads} threads, tier {tier}"},
        "e2e": {"value": value, "unit": "NTT/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------ multi-rank helpers
def max_over_ranks(value: float, world: int, device="cuda") -> float:
    """device-timed duration -> max over ranks (NCCL on GPUs, gloo in the CPU test)"""
    if world <= 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_value(units_per_rank: int, world: int, seconds: float, weak: bool = True, total_units: int = 0) -> float:
    """units/second of the whole job: weak scaling counts world * units_per_rank"""
    units = world * units_per_rank if weak else total_units
    return units / seconds


# ----------------------------------------------------------------- b200 arm
def run_b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import hexl_b200 as hb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the b200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    # stdout is reserved for the one JSON line: libraries that print there (NCCL's version banner
    # does) are sent to stderr at the file-descriptor level, the line goes to the saved descriptor
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << args.logn
    q = hb.GeneratePrimes(1, args.bits, True, n)[0]
    ntt = hb.NTT(n, q)
    batch = args.batch
    g = torch.Generator(device="cuda").manual_seed(42 + rank)
    x = torch.randint(0, q, (batch, n), dtype=torch.int64, device="cuda", generator=g)
    y = torch.empty_like(x)
    z = torch.empty_like(x)

    def step():
        ntt.ComputeForward(y, x, 1, 1)
        ntt.ComputeInverse(z, y, 1, 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    assert torch.equal(z, x), "round trip Inv(Fwd(x)) != x"

    # ---- device-resident throughput
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = hb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = hb.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(ms_total, world)
    ms_step = ms_total / args.steps
    value = whole_job_value(2 * batch, world, ms_step * 1e-3)

    # ---- roofline of the forward transform (its kernels, events on the launch stream)
    fe = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ie = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        fe[k][0].record(); ntt.ComputeForward(y, x, 1, 1); fe[k][1].record()
        ie[k][0].record(); ntt.ComputeInverse(z, y, 1, 1); ie[k][1].record()
    barrier()
    fwd_ms = statistics.mean(a.elapsed_time(b) for a, b in fe)
    inv_ms = statistics.mean(a.elapsed_time(b) for a, b in ie)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    alg_bytes = 16.0 * n * batch
    achieved = alg_bytes / (fwd_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("ntt_forward_bytes_per_launch")
    except (OSError, ValueError):
        pass
    roofline = {"bound": "hbm", "kernel": "hexl_b200_ntt_forward (all kernels of one call)", "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": alg_bytes, "fwd_ms": fwd_ms, "inv_ms": inv_ms,
                "inv_achieved": alg_bytes / (inv_ms * 1e-3) / 1e9,
                "butterflies_per_ntt": (n // 2) * args.logn}
    # The bound that actually binds 64-bit moduli: the FMA-heavy integer pipe.
    # A Shoup butterfly is >= 5 IMAD.WIDE + 4 IMAD = 31 pipe cycles per warp (measured issue
    # intervals 4.6 / 2 cycles, tools/inst_bench.cu), one such pipe per SM sub-partition.
    if q >= (1 << 30):
        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
        sm_hz = 1965.0e6  # max SM clock of this pool's B200s (the clocks line reports the one seen under load)
        peak_bf = sms * 4 * 32 * sm_hz / 31.0
        ach_bf = batch * (n // 2) * args.logn / (fwd_ms * 1e-3)
        roofline["secondary"] = {"bound": "int-multiply pipe (IMAD/IMAD.WIDE)", "achieved": ach_bf / 1e9,
                                 "peak": peak_bf / 1e9, "unit": "G butterflies/s", "frac": ach_bf / peak_bf,
                                 "model": "148 SMs x 4 sub-partitions x 32 lanes x 1.965 GHz / 31 pipe cycles per warp-butterfly"}

    # ---- eltwise kernels (BASELINE configs[2]): algorithmic GB/s at 4096 x 2^16 elements, 60-bit q
    elt = None
    if not args.no_eltwise:
        en = 4096 << 16
        eq = hb.GeneratePrimes(1, 60, True, 1 << 16)[0]
        a = torch.randint(0, eq, (en,), dtype=torch.int64, device="cuda", generator=g)
        b = torch.randint(0, eq, (en,), dtype=torch.int64, device="cuda", generator=g)
        r = torch.empty_like(a)
        ops = {
            "mult_mod": (24, lambda: hb.EltwiseMultMod(r, a, b, en, eq, 1)),
            "fma_mod": (24, lambda: hb.EltwiseFMAMod(r, a, 123456789, b, en, eq, 1)),
            "reduce_mod": (16, lambda: hb.EltwiseReduceMod(r, a, en, eq, eq, 1)),
            "add_mod": (24, lambda: hb.EltwiseAddMod(r, a, b, en, eq)),
        }
        elt = {"n": en, "q_bits": 60}
        for name, (bpe, fn) in ops.items():
            for _ in range(3):
                fn()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            s0.record()
            for _ in range(5):
                fn()
            s1.record()
            torch.cuda.synchronize()
            gbs = bpe * en * 5 / (s0.elapsed_time(s1) * 1e-3) / 1e9
            elt[name] = {"GBps": gbs, "frac_of_hbm_peak": gbs / peak}
        del a, b, r

    # ---- end to end through the host-pointer path of the C ABI
    e2e = None
    if not args.no_e2e:
        # per-GPU batch of the host-pointer leg: the full batch on one GPU; 2048 polynomials
        # (1 GiB per pinned buffer, ~50 ms per step) per rank when several ranks pin host memory at once
        eb = args.e2e_batch or (batch if world == 1 else min(batch, 2048))
        hx = hy = hz = None
        while eb >= 64:
            try:
                hx, hy, hz = (hb.pinned_empty(n * eb) for _ in range(3))
                break
            except hb.HexlB200Error:
                hx = hy = hz = None
                eb //= 2
        # every rank must run the same number of units (and the same barriers): agree on the minimum
        agreed = int(-max_over_ranks(-float(eb if hx is not None else 0), world))
        if agreed < 64:
            hx = None
        elif agreed < eb:
            eb = agreed
            hx, hy, hz = hx[:n * eb], hy[:n * eb], hz[:n * eb]
        if hx is not None:
            rng = np.random.default_rng(7 + rank)
            hx[:] = rng.integers(0, q, size=n * eb, dtype=np.uint64)

            def estep():
                ntt.ComputeForward(hy, hx, 1, 1)   # H2D, kernels, D2H inside the call
                ntt.ComputeInverse(hz, hy, 1, 1)

            estep()
            assert (hz == hx).all(), "e2e round trip failed"
            esteps = max(2, min(args.steps, 5))
            barrier()
            t0 = time.perf_counter()
            for _ in range(esteps):
                estep()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            dt = max_over_ranks(dt, world)
            e2e = {"value": whole_job_value(2 * eb * esteps, world, dt), "unit": "NTT/s",
                   "h2d_bytes_per_step": 2 * 8 * n * eb, "d2h_bytes_per_step": 2 * 8 * n * eb,
                   "batch_per_gpu": eb, "steps": esteps, "ms_per_step": 1e3 * dt / esteps,
                   "path": "hexl_b200_ntt_forward/inverse with pinned HOST pointers (library stages H2D/kernel/D2H in 32 MiB chunks on 3 streams)"}
            # (pinned buffers are released at process exit)

    # ---- CPU baseline (rank 0, single-GPU runs only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = cpu_threads()
        polys = max(threads * 32, 128)
        v, kind, tier = cpu_leg(n, q, threads, polys, 3)
        cpu = {"value": v, "unit": "NTT/s", "cores": threads, "kind": kind,
               "sample": f"{polys} polynomials x (forward + inverse), best of 3, {threads} threads, tier {tier}"}
        # the checker also looks at what the timed kernels produced: the first and last 4 polynomials of
        # the device-resident forward output y and round trip z, bit for bit (outside every timed region)
        import oracle
        chk = oracle.best_checker()
        idx = list(range(4)) + list(range(batch - 4, batch))
        hx = np.stack([x[i].cpu().numpy().view(np.uint64) for i in idx])
        hy = np.stack([y[i].cpu().numpy().view(np.uint64) for i in idx])
        hz = np.stack([z[i].cpu().numpy().view(np.uint64) for i in idx])
        ok = bool((hy.reshape(-1) == chk.ntt_forward(hx.reshape(-1), n, q, 1, 1)).all() and (hz == hx).all())
        cpu["parity"] = {"polynomials_checked": len(idx), "bit_exact": ok, "checker": chk.kind}
        assert ok, "device results differ from the checker"

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "NTT/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"batched Fwd+Inv NTT (configs[1]), N=2^{args.logn}, {args.bits}-bit prime q={q}, "
                                   f"batch={batch} polynomials per GPU, out of place",
                       "parallelism": f"{world} x independent shards, no data-path collective",
                       "l2": f"inputs ({8 * n * batch >> 20} MiB per buffer per GPU) exceed the 126 MB L2; no flush needed"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "eltwise": elt,
        }
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
