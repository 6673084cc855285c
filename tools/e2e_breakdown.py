"""Where does the end-to-end step go?  Wall time per qas_eval_batch call through the public host API against the
GPU time of the same call (stats.last_kernel_ms: first enqueue to the last kernel), with page-locked and with
pageable result buffers.   python tools/e2e_breakdown.py --workload lih_10q"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="lih_10q")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--pre", default="", help="what bench.py does before its end-to-end leg: 'device' (device-pointer steps), "
                                              "'flush' (+ a 256 MB L2 flush buffer)")
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    import __graft_entry__ as ge
    ge.build()
    model, pool, params, ks = bench.make_workload(args.workload, args.batch, 0)
    p, c = params.shape[0], params.shape[1]
    B = len(ks)
    hk = torch.from_numpy(ks.astype(np.uint8)).pin_memory().numpy()
    hp = torch.from_numpy(params).pin_memory().numpy()
    pinned = {"energy": torch.empty(B, dtype=torch.float64).pin_memory().numpy(),
              "grad_dense": torch.empty((p, c, 3), dtype=torch.float64).pin_memory().numpy()}
    ev = model.evaluator(pool, 3)
    if args.pre:
        dev = torch.device("cuda", 0)
        d_k, d_params = torch.from_numpy(ks).to(dev), torch.from_numpy(params).to(dev)
        d_e = torch.zeros(B, dtype=torch.float64, device=dev)
        d_g = torch.zeros((p, c, 3), dtype=torch.float64, device=dev)
        flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev) if "flush" in args.pre else None
        st = torch.cuda.current_stream()
        for i in range(100):
            if flush is not None:
                flush.zero_()
            ev.evaluate_device(d_k, d_params, d_e, None, d_g if i % 2 == 0 else None, stream=st)
        torch.cuda.synchronize()
    for name, out in (("page-locked results", pinned), ("pageable results (fresh arrays)", None), ("page-locked results", pinned)):
        for _ in range(5):
            model.evaluate_batch(hk, hp, pool, want_grad=True, out=out)
        gpu = 0.0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            model.evaluate_batch(hk, hp, pool, want_grad=True, out=out)
            gpu += ev.stats()["last_kernel_ms"]
        wall = (time.perf_counter() - t0) / args.steps * 1e3
        print("%-34s wall %.3f ms per call (%.1f M evals/s), GPU span %.3f ms, host-side remainder %.3f ms"
              % (name, wall, B / wall / 1e3, gpu / args.steps, wall - gpu / args.steps))


if __name__ == "__main__":
    main()
