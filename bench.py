#!/usr/bin/env python
"""bench.py -- candidate-circuit evals/sec (energy + adjoint gradient) of the QAS hot path.

One "step" = one pass of the hot path over one batch of synthetic candidate circuits that
share a parameter super-set: per circuit the energy E and dE/dtheta of its trainable angles,
plus the batch-mean dense gradient (the reduction search() applies, qas/mcts/mcts.py:340-347).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload lih_10q] [--batch B]
    python bench.py --impl reference ...      # the CPU port of the reference path, all host cores

Default workload: LiH sto-6g near_cx search (n=10, c=38, p=20; search-scripts-legacy/
search_LiH_near_cx.py:64-116) -- the configuration BASELINE.json's metric and north_star target
are quoted on ("energy + adjoint gradient at 1/2/4/8 GPUs", ">=1000x on the H2O/LiH searches").
Multi-GPU: one process per GPU (torchrun), the batch is sharded by circuit (weak scaling: B per
GPU fixed), rewards are all-gathered and the dense gradient all-reduced over NCCL.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before CUDA starts: see qas_b200/__init__.py
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


# ------------------------------------------------------------------------------- workloads
def _line(n):
    e = []
    for i in range(n - 1):
        e += [[i, i + 1], [i + 1, i]]
    return e


def _ring(n):
    e = []
    for i in range(n):
        e += [[i, (i + 1) % n], [(i + 1) % n, i]]
    return e


# name -> (model class, n, p, coupling, CNOT limit, param scale, script batch sizes, source)
WORKLOADS = {
    "lih_10q": ("LiH", 10, 20, _line(10), 10, 1.0, (25, 500), "search_LiH_near_cx.py:64-116"),
    "h2o_8q": ("H2O", 8, 50, _line(8), 25, 1.0, (50, 500), "search_H2O_near_cx.py:66-118"),
    "h2_4q": ("FourQubitH2", 4, 40, _line(4), 40, 1.0, (100, 5000), "four_qubit_h2_only_near_cnot.py:43-97"),
    "qaoa_7q": ("QAOAVQCDemo", 7, 15, _ring(7), 7, (2.0 / 2 ** 7) ** 0.5, (100, 100), "qaoa_search_demo.py:41-81"),
    "h2_2q": ("TwoQubitH2", 2, 3, [[0, 1], [1, 0]], 2, 1.0, (50, 10), "two_qubit_h2.py:40-94"),
    # new-API models (qas/qml_models/qchem_model.py, kagome_heisenberg_model.py): larger registers; the
    # 14- and 16-qubit ones run the global-memory instantiation of the evaluation kernel (use --batch 2048)
    "lih_12q": ("LiHSTO3G", 12, 20, _line(12), 10, 1.0, (25, 500), "../h2_sto-3g.py:43-53 (LiHSTO3G, qchem_model.py:36-47)"),
    "h2o_14q": ("H2OSTO3G", 14, 20, _line(14), 10, 1.0, (25, 500), "../h2_sto-3g.py:43-53 (H2OSTO3G, qchem_model.py:49-60)"),
    "kagome_16q": ("KagomeHeisenberg", 16, 500, [[0, 1], [1, 4], [1, 2], [1, 0], [2, 1], [2, 3], [3, 2], [3, 5], [4, 1], [4, 7],
                                                  [5, 8], [5, 3], [6, 7], [7, 4], [7, 10], [7, 6], [8, 5], [8, 9], [8, 11], [9, 8],
                                                  [10, 12], [10, 7], [11, 8], [11, 14], [12, 15], [12, 10], [12, 13], [13, 14],
                                                  [13, 12], [14, 13], [14, 11], [15, 12]], 500, 1.0, (50, 100), "../kagome.py:43-53"),
}


def make_workload(name, B, seed, param_seed=None):
    """structure lists from `seed` (each rank its own shard); the parameter super-set from `param_seed` (the same on
    every rank: data parallelism replicates it, qas/mcts/mcts.py:340-352 averages the gradients of ONE params array)"""
    from qas_b200.qml_models import kagome_heisenberg_model, qchem_model
    from qas_b200.qml_models import qml_models_legacy as L
    from qas_b200.qml_models.qml_gate_ops import QMLPool
    from qas_b200.sampling import sample_structure_lists
    mname, n, p, coupling, cx_lim, scale, script_b, src = WORKLOADS[name]
    model = next(getattr(m, mname) for m in (L, qchem_model, kagome_heisenberg_model) if hasattr(m, mname))
    pool = QMLPool(n, ["Rot", "PlaceHolder"], ["CNOT"], False, coupling)
    c = len(pool)
    params = np.random.default_rng(seed if param_seed is None else param_seed).standard_normal((p, c, 3)) * scale
    ks = sample_structure_lists(pool, p, {"CNOT": cx_lim}, B, seed=seed)
    return model, pool, params, ks


def circuit_flops(ks, pool, n, S, T):
    """Algorithmic flop of a batch.  N1 = non-PlaceHolder one-qubit gates per circuit.
    executed-minimum model: 2^n (58 N1 + 4 S + 4)
    SURVEY.md 8(d) formula:            2^n (122 N1 + 2 (8 G + 2 T))"""
    is_1q = np.array([1 if ("Rot" in pool[k]) else 0 for k in range(len(pool))])
    n1 = is_1q[ks].sum(axis=1).astype(np.float64)
    dim = float(2 ** n)
    return float(np.sum(dim * (58.0 * n1 + 4.0 * S + 4.0))), float(np.sum(dim * (122.0 * n1 + 2.0 * (8.0 * S + 2.0 * T))))


def support_flops(ks, pool, n, xmasks, init, nsample=4096):
    """Flop of the support-restricted adjoint algorithm, mean per circuit over a sample.
    From a basis state the vector after gate g lives on 2^d_g amplitudes, d_g = rank of the xor masks of
    the CNOT-folded tape so far; forward 14, adjoint 14 + 14 + 16 flop per amplitude and gate, all inside
    that support; lambda = H psi is only needed on the final support: 4 flop per amplitude and X mask
    inside the support span, + 4 for the energy.  Uniform start state: d = n throughout."""
    from qas_b200.evaluator import compile_tape_host
    tot = 0.0
    m = min(nsample, len(ks))
    for b in range(m):
        ops, _ = compile_tape_host(n, pool, ks[b], init[1] if init[0] == "basis" else 0)
        piv = {}
        d = n if init[0] == "plus" else 0
        f = 0.0
        for o in ops[::-1]:                      # time order
            x = int(o[0])
            while x and init[0] != "plus":
                h = x.bit_length() - 1
                if h in piv:
                    x ^= piv[h]
                else:
                    piv[h] = x
                    d += 1
                    break
            f += 58.0 * (1 << d)
        s_in = 0
        for xm in xmasks:
            x = int(xm)
            while x and init[0] != "plus":
                h = x.bit_length() - 1
                if h not in piv:
                    break
                x ^= piv[h]
            s_in += 1 if (x == 0 or init[0] == "plus") else 0
        f += (4.0 * s_in + 4.0) * (1 << d)
        tot += f
    return tot / m


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        """nvidia-smi needs ~0.1-0.3 s before its first row: start it BEFORE the warm-up steps, then call mark()
        when the timed region begins."""
        if os.environ.get("QAS_BENCH_NO_SAMPLER") == "1":      # diagnostic: is the sampler itself visible in the numbers?
            self.proc = None
            self.first = 0
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        self.first = 0

    def mark(self):
        """the timed region starts here: rows sampled before it (warm-up, idle) are not used"""
        self.first = len(self.rows)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def samples(self):
        return len(self.rows) - self.first

    def stop(self, keep_busy=None, want=3, budget_s=2.0):
        """keep_busy: callable that runs one more (untimed) step of the same workload; a timed region shorter than
        the sampling period is extended with such steps until `want` rows were taken under load."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        extended = 0
        t0 = time.perf_counter()
        while keep_busy is not None and self.samples() < want and time.perf_counter() - t0 < budget_s:
            keep_busy()
            extended += 1
        self.proc.terminate()
        try:                                   # gone before the next phase is timed
            self.proc.wait(timeout=2.0)
        except Exception:
            self.proc.kill()
        rows = self.rows[self.first:]
        sm = [float(r[0]) for r in rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if extended:
            out["untimed_steps_added_for_sampling"] = extended
        return out


# ------------------------------------------------------------------------------- CPU arms
def _cpu_worker(args):
    name, lo, hi, seed, B = args
    from oracle.refpattern import REF_METHOD, RefPatternEvaluator
    model, pool, params, ks = make_workload(name, B, seed)
    method, sparse = REF_METHOD[model.__name__]
    ev = RefPatternEvaluator(WORKLOADS[name][1], pool.pool, model.terms(), model.init_state, method, sparse)
    t0 = time.perf_counter()
    acc = 0.0
    for b in range(lo, hi):
        e, g = ev.evaluate_one(ks[b], params)
        acc += e + float(g.sum())
    return time.perf_counter() - t0, acc


def cpu_port_throughput(name, nsample, cores, seed):
    """evals/s of the reference-pattern oracle on `cores` host processes over nsample circuits."""
    import multiprocessing as mp
    per = max(1, nsample // cores)
    jobs = [(name, i * per, (i + 1) * per, seed, per * cores) for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        _cpu_worker(jobs[0])
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_worker, jobs)
    dt = time.perf_counter() - t0
    return per * cores / dt, per * cores, dt


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port, reference call pattern) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    name = args.workload
    mname, n, p = WORKLOADS[name][:3]
    # bounded sample per step, sized from a 1-circuit probe so the whole run stays within minutes
    t_probe, _ = _cpu_worker((name, 0, 1, args.seed, 8))
    per_step = max(cores, int(min(4096, max(1, 6.0 / max(t_probe, 1e-4)) * cores)) // cores * cores)
    for _ in range(args.warmup):
        cpu_port_throughput(name, cores, cores, args.seed)
    tot_n, tot_t = 0, 0.0
    for _ in range(args.steps):
        v, nn, dt = cpu_port_throughput(name, per_step, cores, args.seed)
        tot_n += nn
        tot_t += dt
    value = tot_n / tot_t
    from oracle.refpattern import REF_METHOD
    line = {
        "impl": "reference", "metric": "candidate-circuit evals/sec (energy+adjoint grad)", "value": value,
        "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "model": mname, "n_qubits": n, "depth_p": p,
                   "circuits_per_step": per_step, "gradient_method": REF_METHOD[mname][0]},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": "%d circuits/step x %d steps of the %s batch, oracle in the reference call pattern "
                                   "(new model per eval, per-gate tensordot, %s gradient)" % (per_step, args.steps, name, REF_METHOD[mname][0])},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        ge.build()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        dist.barrier()
    if rank != 0:
        ge.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    from qas_b200.evaluator import fp64_peak_tflops

    name = args.workload
    mname, n, p = WORKLOADS[name][:3]
    B = args.batch
    model, pool, params, ks = make_workload(name, B, args.seed + 1000 * rank, param_seed=args.seed)   # each rank its own shard, one parameter super-set
    c = len(pool)
    ev = model.evaluator(pool, 3, local)
    T = len(model.terms())

    import ctypes
    from qas_b200 import _lib
    lib = _lib.load()
    npar = p * c * 3
    d_k = torch.from_numpy(ks).to(dev)
    d_params = torch.from_numpy(params).to(dev)
    d_gc = torch.empty(B, p, 3, dtype=torch.float64, device=dev) if args.compact else None
    # Each rank writes [energies | dense gradient] into ONE buffer, so that the exchange step is one collective:
    # all_gather of the packed rows, then a fixed-order sum of the gathered gradient rows scaled by 1 / (B * world)
    # (qas_reduce_rows: bit-identical on every rank).  Two send / receive buffers alternate so that the exchange of
    # step i runs on a side stream while the kernels of step i + 1 do.
    nrow = (B + npar + 1) & ~1
    packs = [torch.zeros(nrow, dtype=torch.float64, device=dev) for _ in range(2 if world > 1 else 1)]
    gathered = [torch.empty(world * nrow, dtype=torch.float64, device=dev) for _ in range(2)] if world > 1 else None
    # Exchange over NVLink peer memory (csrc/exchange.cu: one push kernel + flag wait + fixed-order reduction) when
    # torch symmetric memory can map the receive buffers into the peers; otherwise ONE NCCL all-gather + reduction.
    px, px_why = None, None
    if world > 1:
        if os.environ.get("QAS_BENCH_EXCHANGE", "") != "nccl":
            try:
                from qas_b200.distributed import PeerExchange
                px = PeerExchange(nrow, dev)
            except Exception as exc:                       # noqa: BLE001 -- any failure means "not available here"
                px_why = "%s: %s" % (type(exc).__name__, str(exc).splitlines()[0][:160])
        else:
            px_why = "QAS_BENCH_EXCHANGE=nccl"
        agree = torch.tensor([1 if px is not None else 0], device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        if not int(agree.item()):
            px = None
    d_mean = torch.empty(p, c, 3, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MB > 126 MB L2
    stream = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=dev, priority=-1) if world > 1 else None   # the exchange kernels take the first free CTA slots

    def compute(i, kk=None):
        pk = packs[i % len(packs)]
        kk = d_k if kk is None else kk
        nb = kk.shape[0]
        # per-circuit energies + the batch-mean dense gradient, which is all search() consumes (mcts.py:340-347);
        # --compact also materialises the per-circuit [B][p][l] gradients.  Multi-GPU: gradient SUM per rank.
        ev.evaluate_device(kk, d_params, pk[:nb], d_gc if (args.compact and nb == B) else None, pk[B:B + npar].view(p, c, 3),
                           avg=(world == 1), stream=stream)

    last_rows = [None]

    def exchange(i, nb_total):
        """the exchange step on the CURRENT stream: the packed rows to every rank + the fixed-order mean of the
        gradient parts"""
        cur = torch.cuda.current_stream()
        if px is not None:
            px.push(packs[i % 2], cur)
            px.wait_reduce(B, npar, 1.0 / nb_total, d_mean, cur)
            last_rows[0] = px.rows()
            return
        g = gathered[i % 2]
        dist.all_gather_into_tensor(g, packs[i % 2])
        rc = lib.qas_reduce_rows(g.data_ptr() + 8 * B, world, nrow, npar, 1.0 / nb_total, d_mean.data_ptr(),
                                 ctypes.c_void_p(cur.cuda_stream))
        assert rc == 0
        last_rows[0] = g.view(world, nrow)

    def step_sync(i, kk=None, nb_total=None):
        compute(i, kk)
        if world > 1:
            exchange(i, nb_total if nb_total is not None else B * world)

    for i in range(args.warmup):
        flush.zero_()
        step_sync(i)
    torch.cuda.synchronize()

    # ---- multi-GPU parity self-check (untimed): the sharded batch and the sharded statevector against one GPU
    parity = None
    if world > 1:
        parity = multi_gpu_parity_check(args, name, model, pool, ev, B, p, c, rank, world, dev, packs, last_rows, d_mean,
                                        d_params, step_sync, lib)
        flag = torch.tensor([0 if (rank != 0 or parity["ok"]) else 1], device=dev)
        dist.all_reduce(flag)
        if int(flag.item()):
            if rank == 0:
                print(json.dumps({"parity_check": parity, "error": "multi-GPU parity check failed"}), flush=True)
            dist.destroy_process_group()
            sys.exit(3)
        for i in range(2):
            step_sync(i)
        torch.cuda.synchronize()

    sampler = ClockSampler(local)

    def busy_step():
        compute(0)
        torch.cuda.synchronize()

    if rank == 0:
        sampler.start()
        t_w = time.perf_counter()
        while not sampler.rows and sampler.proc is not None and time.perf_counter() - t_w < 2.0:
            busy_step()              # nvidia-smi start-up: the GPU stays under load until its first row
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark()
    launches0 = ev.stats()["launches"]
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    x0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    x1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    fl = [torch.cuda.Event() for _ in range(args.steps)]
    eval_ms = []
    # Timed region.  Every step is bracketed by its own event pair; the 256 MB L2 flush sits between the pairs.
    # Multi-GPU: the exchange of step i runs on the side stream and may only START after the flush that follows
    # step i (so the untimed flush never hides it) and must be FINISHED before the end event of step i + 1 (so
    # it is either overlapped with timed kernels or shows up in that step's time); the exchange of the last step
    # has nothing to hide behind and is added in full.
    serial_exchange = os.environ.get("QAS_BENCH_EXCHANGE", "overlap") == "serial"
    flush.zero_()
    for i in range(args.steps):
        e0[i].record(stream)
        compute(i)
        if world > 1 and i >= 1:
            stream.wait_event(x1[i - 1])
        e1[i].record(stream)
        flush.zero_()
        if world > 1 and serial_exchange:      # diagnostic: the exchange on the compute stream, nothing overlapped
            x0[i].record(stream)
            exchange(i, B * world)
            x1[i].record(stream)
        elif world > 1:
            fl[i].record(stream)
            with torch.cuda.stream(side):
                side.wait_event(fl[i])
                x0[i].record(side)
                exchange(i, B * world)
                x1[i].record(side)
        e1[i].synchronize()
        eval_ms.append(ev.stats()["last_eval_kernel_ms"])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    step_ms = [a.elapsed_time(b) for a, b in zip(e0, e1)]
    comm_ms = [a.elapsed_time(b) for a, b in zip(x0, x1)] if world > 1 else []
    # the exchange of the last step has nothing left to hide behind: its whole duration is exposed
    tail_ms = comm_ms[-1] if comm_ms else 0.0
    total_ms = torch.tensor([sum(step_ms) + tail_ms], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        mine_ms = torch.tensor([float(total_ms.item()) / args.steps, statistics.mean(eval_ms)], dtype=torch.float64, device=dev)
        all_ms = torch.empty(2 * world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(all_ms, mine_ms)
        per_rank = all_ms.view(world, 2).cpu().numpy()
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    launches = ev.stats()["launches"] - launches0        # this library's kernels only
    clocks = sampler.stop(keep_busy=busy_step) if rank == 0 else None
    launches += (3 if px is not None else 1) * args.steps if world > 1 else 0   # + push / wait / reduce (or qas_reduce_rows) per exchange
    if px is not None and px.timed_out():
        print(json.dumps({"error": "peer exchange: a rank's row did not arrive within 5 s"}), flush=True)
        sys.exit(4)
    total_s = float(total_ms.item()) * 1e-3
    value = world * B * args.steps / total_s
    d_energy = packs[0][:B]

    # ---- reward-only evaluations (energy without gradient): rewards outnumber gradients 10-170x in a
    # real search, so their rate is reported next to the headline metric
    for _ in range(2):
        ev.evaluate_device(d_k, d_params, d_energy, stream=stream)
    torch.cuda.synchronize()
    r0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    r1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()
        r0[i].record(stream)
        ev.evaluate_device(d_k, d_params, d_energy, stream=stream)
        r1[i].record(stream)
    torch.cuda.synchronize()
    reward_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in zip(r0, r1))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(reward_ms, op=dist.ReduceOp.MAX)
    reward_value = world * B * args.steps / (float(reward_ms.item()) * 1e-3)

    # ---- end to end through the public host API: pinned host buffers in, host results out
    # structure lists as one byte per key when the pool allows it (QAS_F_K_U8: a quarter of the upload)
    k8 = c <= 256
    h_k = torch.from_numpy(ks.astype(np.uint8) if k8 else ks).pin_memory()
    h_params = torch.from_numpy(params).pin_memory()
    hk_np, hp_np = h_k.numpy(), h_params.numpy()
    # results land in page-locked buffers too ("host buffers in and out"): a true DMA read-back
    h_out = {"energy": torch.empty(B, dtype=torch.float64).pin_memory().numpy(),
             "grad_dense": torch.empty((p, c, 3), dtype=torch.float64).pin_memory().numpy()}
    for _ in range(2):
        model.evaluate_batch(hk_np, hp_np, pool, want_grad=True, out=h_out)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = model.evaluate_batch(hk_np, hp_np, pool, want_grad=True, out=h_out)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(e2e_s.item())
    h2d = hk_np.nbytes + params.nbytes
    d2h = out["energy"].nbytes + out["grad_dense"].nbytes
    # ... and with plain (pageable) NumPy arrays, which is what the drop-in model classes receive from search()
    pk_np, pp_np = ks.copy(), params.copy()
    for _ in range(2):
        model.evaluate_batch(pk_np, pp_np, pool, want_grad=True)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.evaluate_batch(pk_np, pp_np, pool, want_grad=True)
    e2e_p = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_p, op=dist.ReduceOp.MAX)
    e2e_pageable = world * B * args.steps / float(e2e_p.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the evaluation kernel), FP64 pipe
    S = None
    try:
        import ctypes  # noqa: F401
        S = len({(w.replace("Y", "X").replace("Z", "I")) for _, w in model.terms()})
    except Exception:
        pass
    f_exec, f_survey = circuit_flops(ks, pool, n, S, T)
    from qas_b200 import hamiltonians as _H
    xmasks = sorted({_H.word_to_masks(w)[0] for _, w in model.terms()})
    f_supp = support_flops(ks, pool, n, xmasks, model.init_state) * len(ks)
    k_ms = statistics.mean(eval_ms)
    peak = fp64_peak_tflops(local, 5)
    achieved = f_supp / (k_ms * 1e-3) / 1e12
    traffic = None
    try:      # DRAM bytes per launch of the same kernel/workload from the committed ncu capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if name in tj and B == 65536:
            traffic = tj[name]["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic,
                "kernel": ("evalc_kernel<T> per support bin (T = 16/32/64 lanes per circuit) + eval_kernel<%d> for the full bin" % n)
                          if ev.stats().get("path") == 3 else "eval_kernel<%d>" % n,
                "kernel_ms": k_ms,
                "kernel_share_of_step": k_ms * args.steps / sum(step_ms),
                "achieved_dense_adjoint_model": f_exec / (k_ms * 1e-3) / 1e12,
                "achieved_survey_8d_formula": f_survey / (k_ms * 1e-3) / 1e12,
                "peak_source": "measured live: qas_fp64_peak_tflops DFMA micro-benchmark (MEASURED_PEAKS.json has no FP64 entry)",
                "flop_model": "support-restricted adjoint algorithm: sum over gates of 58 * 2^d_g (d_g = dimension of the "
                              "state's support after gate g) + (4 S_in + 4) * 2^d_final for H on the final support "
                              "(mean over a 4096-circuit sample of the batch). The dense-algorithm figures "
                              "2^n(58 N1 + 4 S + 4) and SURVEY 8(d)'s 2^n(122 N1 + 2(8G+2T)) are reported alongside: "
                              "they count work on amplitudes that are exactly zero, which the kernel never touches"}

    # ---- CPU baseline: reference-pattern oracle on a bounded sample, 1 core (how the reference runs)
    if args.tune:      # kernel-tuning runs skip the CPU leg; such a line is not a bench result
        cpu_baseline = None
    else:
        t_probe, _ = _cpu_worker((name, 0, 1, args.seed, 8))
        ns = int(max(2, min(4096, 15.0 / max(t_probe, 1e-3))))
        cpu_v, cpu_n, cpu_t = cpu_port_throughput(name, ns, 1, args.seed)
        from oracle.refpattern import REF_METHOD
        cpu_baseline = {"value": cpu_v, "unit": "evals/s", "cores": 1, "kind": "port",
                        "sample": "%d circuits of the %s batch (%.1f s), oracle in the reference call pattern, %s gradient"
                                  % (cpu_n, name, cpu_t, REF_METHOD[mname][0])}

    line = {
        "metric": "candidate-circuit evals/sec (energy+adjoint grad)", "value": value, "unit": "evals/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "model": mname, "n_qubits": n, "depth_p": p, "pool_c": c,
                   "circuits_per_gpu_per_step": B, "pauli_terms": T, "xmask_groups": S,
                   "source": "search-scripts-legacy/" + WORKLOADS[name][7], "l2_flush_between_steps": True,
                   "parallelism": "circuits sharded over %d GPU(s); packed [rewards | gradient sum] rows to every rank + fixed-order mean" % world},
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "key_dtype": "u8" if k8 else "i32", "host_buffers": "pinned"},
        "e2e_pageable": {"value": e2e_pageable, "unit": "evals/s",
                         "note": "same call with pageable int32 NumPy arrays (what search() hands the drop-in classes)"},
        "reward_only": {"value": reward_value, "unit": "evals/s", "ms_per_step": float(reward_ms.item()) / args.steps,
                        "note": "energy without gradient (getReward), same batch, device-resident inputs"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    if world > 1:
        line["parity_check"] = parity
        line["per_rank"] = {"ms_per_step": [round(float(x), 4) for x in per_rank[:, 0]], "kernel_ms": [round(float(x), 4) for x in per_rank[:, 1]]}
        line["exchange"] = ("one push kernel per rank over NVLink peer memory + flag wait + fixed-order reduction (csrc/exchange.cu), "
                            "no collective library on the data path") if px is not None else \
                           ("NCCL all_gather_into_tensor + qas_reduce_rows (peer path not used: %s)" % px_why)
        line["comm_us_per_step"] = 1e3 * statistics.mean(comm_ms)
        line["comm_exposed_us_per_step"] = 1e3 * tail_ms / args.steps
        line["config"]["exchange"] = ("one all_gather of [rewards | gradient sum] per rank (%d B) + fixed-order reduction kernel, on a "
                                      "side stream overlapped with the next step's kernels" % (8 * (B + npar)))
    elif not args.tune and not args.no_others:
        line["other_workloads"] = {w: secondary_workload(w, args, dev, local, flush, peak) for w in ("h2o_8q", "qaoa_7q", "h2_4q")
                                   if w != name}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def secondary_workload(name, args, dev, local, flush, peak):
    """The other BASELINE configurations in the same run (same box, same clocks record): device-timed value,
    evaluation-kernel time and roofline fraction, measured like the headline workload (fewer steps)."""
    import torch
    mname, n, p = WORKLOADS[name][:3]
    B = args.batch
    model, pool, params, ks = make_workload(name, B, args.seed)
    c = len(pool)
    ev = model.evaluator(pool, 3, local)
    d_k = torch.from_numpy(ks).to(dev)
    d_params = torch.from_numpy(params).to(dev)
    d_e = torch.empty(B, dtype=torch.float64, device=dev)
    d_gd = torch.empty(p, c, 3, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    steps = max(5, args.steps // 2)
    for _ in range(3):
        flush.zero_()
        ev.evaluate_device(d_k, d_params, d_e, None, d_gd, stream=stream)
    torch.cuda.synchronize()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    k_ms = []
    for i in range(steps):
        flush.zero_()
        e0[i].record(stream)
        ev.evaluate_device(d_k, d_params, d_e, None, d_gd, stream=stream)
        e1[i].record(stream)
        e1[i].synchronize()
        k_ms.append(ev.stats()["last_eval_kernel_ms"])
    ms = sum(a.elapsed_time(b) for a, b in zip(e0, e1)) / steps
    from qas_b200 import hamiltonians as _H
    xmasks = sorted({_H.word_to_masks(w)[0] for _, w in model.terms()})
    f_supp = support_flops(ks, pool, n, xmasks, model.init_state, nsample=1024) * len(ks)
    km = statistics.mean(k_ms)
    return {"value": B / (ms * 1e-3), "unit": "evals/s", "ms_per_step": ms, "kernel_ms": km, "steps": steps,
            "frac": f_supp / (km * 1e-3) / 1e12 / peak, "n_qubits": n, "depth_p": p, "circuits_per_step": B,
            "source": "search-scripts-legacy/" + WORKLOADS[name][7]}


# ------------------------------------------------------------------------------- large-state mode
LARGE_WORKLOADS = {"hea_%dq" % q: (q, 4) for q in range(24, 34)}     # 2^n x 16 B must fit the GPUs it is sharded over


def run_largestate(args):
    """--workload hea_NNq: BASELINE.json config 5 -- one statevector of NN qubits (hardware-efficient ansatz, L = 4
    layers of Rot on every wire + nearest-neighbour CNOT ladder, observable sum Z_i Z_{i+1}), on one
    GPU or sharded over --gpus N with NCCL qubit swaps.  One step = reset + all gate passes + expectation value.
    HBM roofline of the gate passes: passes * 32 * 2^n_local bytes per GPU / device time of the pass kernels."""
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        ge.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        dist.barrier()
    if rank != 0:
        ge.build()
    from qas_b200.largestate import LargeStateSimulator, hea_gate_list, zz_chain_terms
    n, layers = LARGE_WORKLOADS[args.workload]
    gl = hea_gate_list(n, layers, seed=args.seed)
    terms = zz_chain_terms(n)
    sim = LargeStateSimulator(n, device=local)
    solo = dist.new_group([0]) if world > 1 else None          # collective: every rank takes part
    steps = min(args.steps, 20)
    sampler = ClockSampler(local)

    def one():
        sim.reset()
        sim.run(gl)
        return sim.expval(terms)

    for _ in range(max(2, min(args.warmup, 3))):
        energy = one()
    torch.cuda.synchronize()
    be = sim.backend
    be.passes, be.tma_passes, be.kernel_ms, be.bytes, sim.swaps = 0, 0, 0.0, 0.0, 0
    if rank == 0:
        sampler.start()
        time.sleep(0.3)              # nvidia-smi start-up (a step is tens of milliseconds: rows follow by themselves)
        sampler.mark()
    if world > 1:
        be.swap_ms()
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        energy = one()
    e1.record()
    torch.cuda.synchronize()
    total_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    swap_ms = be.swap_ms() if world > 1 else 0.0
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        dist.destroy_process_group()
        return
    total_s = float(total_ms.item()) * 1e-3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = be.bytes / (be.kernel_ms * 1e-3) / 1e9
    # The gate passes are co-bound by the FP64 pipe: a 2x2 complex unitary on an amplitude pair is 16 DFMA-class
    # instructions (32 flop with FMA = 2, the convention of the live DFMA peak), CNOTs are folded into the
    # indexing and cost none.  Both rooflines are reported; the larger of the two peak times bounds a pass.
    from qas_b200.evaluator import fp64_peak_tflops
    n_u = sum(1 for g in gl if g[0] not in ("CNOT", "SWAP", "PlaceHolder"))
    flop_step = 32.0 * n_u * 2.0 ** (sim.n_local - 1)
    fp64_peak = fp64_peak_tflops(local)
    fp64_ach = flop_step * steps / (be.kernel_ms * 1e-3) / 1e12
    hbm_floor_ms = be.bytes / steps / (peak * 1e9) * 1e3
    fp64_floor_ms = flop_step / (fp64_peak * 1e12) * 1e3
    # CPU baseline: the same algorithm through the NumPy slab backend (tests/sv_numpy_backend.py) on a 22-qubit,
    # one-layer sample, scaled by 2^(n - 22) * L (the work is linear in both)
    cpu_baseline = None
    if not args.tune:
        from sv_numpy_backend import NumpySvBackend
        ns, ls = 22, 1
        cs = LargeStateSimulator(ns, backend=NumpySvBackend(), group=solo)
        t0 = time.perf_counter()
        cs.run(hea_gate_list(ns, ls, seed=args.seed))
        cs.expval(zz_chain_terms(ns))
        dt = time.perf_counter() - t0
        scaled = dt * (2.0 ** (n - ns)) * layers / ls
        cpu_baseline = {"value": 1.0 / scaled, "unit": "evals/s", "cores": 1, "kind": "port",
                        "sample": "%d-qubit, %d-layer HEA through the NumPy slab backend (%.1f s), scaled by 2^%d x %d" % (ns, ls, dt, n - ns, layers)}
    line = {
        "metric": "large-state HEA energy evaluations/sec", "value": steps / total_s, "unit": "evals/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / steps, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "n_qubits": n, "layers": layers, "gates": len(gl), "n_local": sim.n_local,
                   "observable": "sum_i Z_i Z_{i+1}", "state_bytes_per_gpu": 16 * 2 ** sim.n_local,
                   "inputs_larger_than_l2": True,
                   "parallelism": "one statevector sharded over %d GPU(s), NCCL all-to-all qubit swaps" % world},
        "energy": energy,
        "e2e": {"value": steps / total_s, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8,
                "note": "the public call takes a gate list and returns one double: end to end = the timed region"},
        "gpu_launches": int(be.passes + 2 * steps),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "kernel": "sv_block_kernel", "kernel_ms": be.kernel_ms / steps, "passes_per_step": be.passes / steps,
                     "tma_passes_per_step": be.tma_passes / steps,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                     "note": "algorithmic bytes = passes * 32 * 2^n_local per GPU (one read + one write of the slab per pass)",
                     "fp64": {"achieved": fp64_ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_ach / fp64_peak,
                              "unitary_gates": n_u, "flop_per_step_per_gpu": flop_step,
                              "peak_source": "measured live: qas_fp64_peak_tflops DFMA micro-benchmark"},
                     "floor_ms_per_step": {"hbm": hbm_floor_ms, "fp64": fp64_floor_ms,
                                           "note": "time of one step's gate passes at the HBM peak / at the FP64 peak: the passes "
                                                   "cannot be faster than the larger of the two, so the HBM fraction is capped at "
                                                   "hbm / max(hbm, fp64)"}},
        "cpu_baseline": cpu_baseline,
    }
    if world > 1:
        per_swap = 16.0 * 2 ** sim.n_local * (world - 1) / world
        line["swaps_per_step"] = sim.swaps / steps
        line["swap"] = {"ms_per_swap": swap_ms / max(1, sim.swaps), "gb_per_s_per_direction": per_swap * sim.swaps / (swap_ms * 1e-3) / 1e9 if swap_ms else None,
                        "nvlink_gb_per_s_per_direction": 900.0, "bytes_sent_per_gpu_per_swap": per_swap}
        dist.destroy_process_group()
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- search epoch
# per workload: (search batch, sampling rounds, exploit rounds, warm-up batch, lr, placeholder-penalty scale)
# from the reference scripts: four_qubit_h2_only_near_cnot.py:88-97, search_LiH_near_cx.py:107-116,
# search_H2O_near_cx.py:109-118, qaoa_search_demo.py:72-81
SEARCH_CONFIGS = {"h2_4q": (100, 10, 100, 5000, 1.0, 1.0), "lih_10q": (25, 10, 20, 500, 0.1, 1.0),
                  "h2o_8q": (50, 10, 20, 500, 0.1, 10.0), "qaoa_7q": (100, 100, 500, 100, 0.1, 1.0)}


def run_search_epoch(args):
    """--search-epoch: wall time of whole search() epochs in a reference script's
    configuration (default workload lih_10q = search_LiH_near_cx.py) with the B200 evaluator behind
    the drop-in model class -- native tree policy and the reference's Python tree policy -- next to
    the same driver on the CPU oracle in the reference's one-model-per-evaluation pattern."""
    import contextlib
    import io
    import random
    import re
    import __graft_entry__ as ge
    ge.build()
    from oracle.refpattern import REF_METHOD, RefPatternEvaluator
    from qas_b200.mcts import QMLStateBasicGates, driver, search
    from qas_b200.mcts.native import PlaceholderPenalty
    from qas_b200.models import ModelFromK
    from qas_b200.optim import AdamOptimizer
    name = args.workload
    mname, n, p, coupling, cx_lim, scale = WORKLOADS[name][:6]
    bs, rs, re_, warm_bs, lr, pen_scale = SEARCH_CONFIGS[name]
    model, pool, _, _ = make_workload(name, 1, 0)
    l, c = 3, len(pool)
    method, sparse = REF_METHOD[mname]
    ref_ev = RefPatternEvaluator(n, pool.pool, model.terms(), model.init_state, method, sparse)
    sign = float(getattr(model, "reward_sign", -1.0))

    class RefPatternModel(ModelFromK):          # one circuit per call, like the reference's class
        name = mname + "_cpu_reference_pattern"

        def __init__(self, p_, c_, l_, structure_list, op_pool):
            self.k = [int(x) for x in structure_list]

        def getLoss(self, params):
            return ref_ev._energy(self.k, params)

        def getReward(self, params):
            return sign * ref_ev._energy(self.k, params)

        def getGradient(self, params):
            return ref_ev.gradient(self.k, params)

        def toList(self, params):
            return []

    def epochs(mdl, n_warm, n_search, wbs, native, leaves=1):
        driver.USE_NATIVE_TREE = native
        np.random.seed(args.seed)
        random.seed(args.seed)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            search(model=mdl, op_pool=pool, target_circuit_depth=p, init_qubit_with_controls=set(),
                   init_params=np.random.randn(p, c, l) * scale, num_iterations=n_warm + n_search,
                   num_warmup_iterations=n_warm, super_circ_train_optimizer=AdamOptimizer,
                   super_circ_train_gradient_noise_factor=0.0, early_stop_threshold=1e9, early_stop_lookback_count=5,
                   super_circ_train_lr=lr, penalty_function=PlaceholderPenalty(pool, p, pen_scale),
                   gate_limit_dict={"CNOT": cx_lim}, warmup_arc_batchsize=wbs, search_arc_batchsize=bs, alpha_max=2,
                   alpha_decay_rate=0.99, prune_constant_max=0.99, prune_constant_min=0.80,
                   max_visits_prune_threshold=20, min_num_children=c // 2 + 1, sampling_execute_rounds=rs,
                   exploit_execute_rounds=re_, verbose=1, state_class=QMLStateBasicGates, search_reset=True,
                   leaf_parallel=leaves)
        secs = [float(x) for x in re.findall(r"\| ([0-9.]+) s", buf.getvalue())]
        return secs[:n_warm], secs[n_warm:]

    epochs(model, 1, 1, warm_bs, True)                       # first-call costs (context, tables)
    gw, gs = epochs(model, 2, 3, warm_bs, True)              # native tree policy (C++), rewards resident on the GPU
    lw, ls = epochs(model, 2, 3, warm_bs, True, args.leaves)  # ... with leaf-parallel rounds under a virtual loss (opt-in)
    pw, ps = epochs(model, 2, 3, warm_bs, False)             # the reference's tree policy in Python, GPU rewards
    cpu_wbs = max(10, warm_bs // 10)
    cw, cs = epochs(RefPatternModel, 1, 1, cpu_wbs, False)   # CPU: a tenth of the warm-up batch, scaled below
    driver.USE_NATIVE_TREE = "auto"
    line = {"metric": "search() epoch wall time", "unit": "s",
            "config": {"workload": name, "model": mname, "depth_p": p, "search_arc_batchsize": bs, "sampling_rounds": rs,
                       "exploit_rounds": re_, "warmup_arc_batchsize": warm_bs, "source": "search-scripts-legacy/" + WORKLOADS[name][7]},
            "b200": {"warmup_epoch_s": min(gw), "search_epoch_s": statistics.median(gs), "tree": "native (libqas_b200 qas_mcts_*)"},
            "b200_leaf_parallel": {"warmup_epoch_s": min(lw), "search_epoch_s": statistics.median(ls), "leaves_per_wave": args.leaves,
                                   "note": "opt-in, not semantics-preserving (virtual loss); final energies over 20 seeds: tools/leaf_parallel_study.py"},
            "b200_python_tree": {"warmup_epoch_s": min(pw), "search_epoch_s": statistics.median(ps)},
            "cpu_reference_pattern": {"warmup_epoch_s_scaled": cw[0] * warm_bs / cpu_wbs, "search_epoch_s": cs[0], "cores": 1,
                                      "note": "oracle in the reference call pattern (%s gradient); warm-up epoch run with %d "
                                              "arcs and scaled to %d" % (method, cpu_wbs, warm_bs)},
            "note": "epoch times as printed by the driver (0.1 ms resolution); with the Python tree the epochs are bound by the "
                    "tree policy, not by the evaluator; the native tree removes that"}
    print(json.dumps(line), flush=True)


def multi_gpu_parity_check(args, name, model, pool, ev, B, p, c, rank, world, dev, packs, last_rows, d_mean, d_params,
                           step_sync, lib):
    """Untimed self-check of both NCCL paths on the hardware the bench runs on (the GPU tests that cover them
    need >= 2 GPUs and are skipped on a one-GPU test box).
      batch: rank 0 re-evaluates, on its own GPU, the first 1024 / world circuits of EVERY rank's shard and compares
             them with the all-gathered rewards of a step;
      grad:  a 4096-circuit batch, sharded round-robin, packed all-gather + fixed-order reduction, against the same
             batch evaluated on rank 0 alone -- the reduction is np.mean(np.nan_to_num(stack), axis=0) of
             qas/mcts/mcts.py:345-347;
      sv:    a (26 + log2 world)-qubit HEA (L = 2) statevector sharded over all ranks with NCCL qubit swaps against
             the unsharded run on rank 0.
    Returns the report on rank 0 ({} elsewhere)."""
    import torch
    import torch.distributed as dist
    from qas_b200.largestate import LargeStateSimulator, hea_gate_list, zz_chain_terms
    npar = p * c * 3
    rep = {}
    # ---- (1) rewards of the sharded bench batch
    step_sync(0)
    torch.cuda.synchronize()
    all_e = last_rows[0][:, :B].cpu().numpy()
    m = max(1, 1024 // world)
    if rank == 0:
        worst = 0.0
        for r in range(world):
            ks_r = make_workload(name, B, args.seed + 1000 * r)[3][:m]
            e_r = ev.evaluate(ks_r, d_params.cpu().numpy(), want_grad=False)["energy"]
            worst = max(worst, float(np.max(np.abs(e_r - all_e[r, :m]))))
        rep["batch_max_abs"] = worst
        rep["batch_circuits_checked"] = m * world
    dist.barrier()
    # ---- (2) mean gradient of a small full batch
    Bs = 4096
    ks_s = make_workload(name, Bs, args.seed + 77)[3]
    mine = torch.from_numpy(np.ascontiguousarray(ks_s[rank::world])).to(dev)
    step_sync(0, mine, Bs)
    torch.cuda.synchronize()
    if rank == 0:
        single = ev.evaluate(ks_s, d_params.cpu().numpy(), want_grad=True, dense=True, compact=False)
        rep["grad_max_abs"] = float(np.max(np.abs(single["grad_dense"] - d_mean.cpu().numpy())))
        e_sh = last_rows[0][:, :Bs // world].cpu().numpy()
        e_full = np.empty(Bs)
        for r in range(world):
            e_full[r::world] = e_sh[r]
        rep["grad_batch_energy_max_abs"] = float(np.max(np.abs(e_full - single["energy"])))
    dist.barrier()
    # ---- (3) sharded statevector against the unsharded one
    n_sv = 26 + int(round(np.log2(world)))
    gl = hea_gate_list(n_sv, 2, seed=0)
    terms = zz_chain_terms(n_sv)
    solo = dist.new_group([0])
    sim = LargeStateSimulator(n_sv, device=dev.index)
    sim.run(gl)
    e_sharded = sim.expval(terms)
    torch.cuda.synchronize()
    swaps = sim.swaps
    del sim
    torch.cuda.empty_cache()
    if rank == 0:
        one = LargeStateSimulator(n_sv, device=dev.index, group=solo)
        one.run(gl)
        e_one = one.expval(terms)
        del one
        torch.cuda.empty_cache()
        rep["sv_abs"] = abs(e_sharded - e_one)
        rep["sv_qubits"] = n_sv
        rep["sv_swaps"] = swaps
        rep["sv_energy"] = e_one
        rep["ok"] = bool(rep["batch_max_abs"] <= 1e-12 and rep["grad_max_abs"] <= 1e-12 and
                         rep["grad_batch_energy_max_abs"] <= 1e-12 and rep["sv_abs"] <= 1e-9)
        rep["tolerances"] = {"batch_max_abs": 1e-12, "grad_max_abs": 1e-12, "sv_abs": 1e-9}
    dist.barrier()
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 600 (b200 arm: >= 0.5 s of timed region), 10 (reference arm)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lih_10q", choices=sorted(WORKLOADS) + sorted(LARGE_WORKLOADS))
    ap.add_argument("--batch", type=int, default=65536, help="circuits per GPU per step")
    ap.add_argument("--seed", type=int, default=int(os.environ.get("QAS_BENCH_SEED", "0")))
    ap.add_argument("--tune", action="store_true", help="kernel tuning: skip the CPU-baseline leg")
    ap.add_argument("--no-others", action="store_true", help="skip the other_workloads block of the default line")
    ap.add_argument("--compact", action="store_true", help="also write the per-circuit compact gradients [B][p][l] in the timed step")
    ap.add_argument("--search-epoch", action="store_true", help="time whole search() epochs instead (SURVEY 8d)")
    ap.add_argument("--leaves", type=int, default=8, help="--search-epoch: selection rounds per wave of the leaf-parallel mode")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 600 if (args.impl == "b200" and not args.search_epoch) else 10
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.workload in LARGE_WORKLOADS and args.impl == "b200":
        run_largestate(args)
    elif args.search_epoch:
        run_search_epoch(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
