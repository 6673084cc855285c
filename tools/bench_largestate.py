"""Large-state mode benchmark: L-layer hardware-efficient ansatz (Rot on every wire + CNOT ladder)
on an n-qubit statevector, observable sum Z_i Z_{i+1}  (BASELINE.json config 5).

  python tools/bench_largestate.py --qubits 28 --layers 4                      # one GPU
  torchrun --nproc-per-node 2 tools/bench_largestate.py --qubits 29 ...        # sharded, NCCL swaps

Reports HBM roofline of the gate passes: achieved = passes * 32 * 2^n_local bytes / kernel time,
peak = MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=28)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from qas_b200.largestate import LargeStateSimulator, hea_gate_list, zz_chain_terms
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    n = args.qubits
    gl = hea_gate_list(n, args.layers, seed=0)
    terms = zz_chain_terms(n)
    sim = LargeStateSimulator(n, device=local, tile_bits=args.tile_bits)
    best = None
    for rep in range(args.reps + 1):
        sim.reset()
        sim.backend.passes, sim.backend.kernel_ms, sim.backend.bytes, sim.swaps = 0, 0.0, 0.0, 0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sim.run(gl)
        torch.cuda.synchronize()
        t_run = time.perf_counter() - t0
        e = sim.expval(terms)
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        rec = {"run_s": t_run, "total_s": t_all, "kernel_ms": sim.backend.kernel_ms, "passes": sim.backend.passes,
               "bytes": sim.backend.bytes, "swaps": sim.swaps, "energy": e}
        if rep > 0 and (best is None or rec["total_s"] < best["total_s"]):
            best = rec
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        achieved = best["bytes"] / (best["kernel_ms"] * 1e-3) / 1e9
        print(json.dumps({
            "metric": "large-state HEA energy evaluations/sec", "value": 1.0 / best["total_s"], "unit": "evals/s",
            "n_gpus": world, "config": {"workload": "hea_%dq_L%d" % (n, args.layers), "gates": len(gl),
                                        "n_local": sim.n_local, "tile_bits": args.tile_bits or 12},
            "energy": best["energy"], "passes": best["passes"], "swaps": best["swaps"],
            "run_s": best["run_s"], "total_s": best["total_s"], "gate_kernel_ms": best["kernel_ms"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650",
                         "note": "algorithmic bytes = passes * 32 * 2^n_local per GPU; gate passes only"},
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
