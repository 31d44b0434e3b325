"""cfg5 in miniature or in full: a batch of independent ComplexF64 QTT vectors (d=30, n=2, rank 64) through
`y = A*x` (shared ComplexF64 MPO of rank W) followed by `tt_compress!(y, 64)`, sharded over the ranks with no
data-path collective (each rank owns total/world vectors, processed in chunks that fit in HBM).

usage: [torchrun ...] python tools/batch_bench.py --total 512 --chunk 64 [--d 30 --rank 64 --W 4]
Prints one JSON line on rank 0: whole-job vectors/s (max time over ranks).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=64)
    ap.add_argument("--d", type=int, default=30)
    ap.add_argument("--rank", type=int, default=64)
    ap.add_argument("--W", type=int, default=4)
    ap.add_argument("--check", action="store_true", help="(kept for old command lines; parity lives in tests/)")
    ap.add_argument("--profile", action="store_true", help="per-kernel-family CUDA-event timing")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_
    import ttn_b200 as t
    d, r, W = args.d, args.rank, args.W
    rks = [min(2 ** k, 2 ** (d - k), r) for k in range(d + 1)]
    Rk = [min(4 ** k, 4 ** (d - k), W) for k in range(d + 1)]
    rng = np.random.default_rng(7)
    A = t.TToperator(d, [np.asfortranarray((rng.standard_normal((2, 2, Rk[k], Rk[k + 1])) + 1j * rng.standard_normal((2, 2, Rk[k], Rk[k + 1]))) / np.sqrt(2.0 * Rk[k + 1]))
                         for k in range(d)], (2,) * d, Rk)
    Ad = t.DeviceTTO.upload(A)
    mine = args.total // world
    nchunks = max(1, mine // args.chunk)
    chunk = mine // nchunks

    def make_chunk(seed):
        g = np.random.default_rng(100 + seed)
        cores = []
        for k in range(d):
            shp = (2, rks[k], rks[k + 1], chunk)
            c = (g.standard_normal(shp) + 1j * g.standard_normal(shp)) / np.sqrt(4.0 * rks[k + 1])
            cores.append(np.asfortranarray(c))
        return cores

    class Batch:  # duck-typed list of TTvectors sharing stacked storage
        pass

    def upload(cores):
        xs = [t.TTvector(d, [c[..., b] for c in cores], (2,) * d, rks) for b in range(chunk)]
        return t.DeviceTT.upload(xs), xs

    host = make_chunk(rank)
    xd, xs = upload(host)
    # warm-up (twice, results dropped: the timed loop must not grow the stream-ordered memory pool)
    for _ in range(2):
        y = t.tt_compress_(t.apply(Ad, xd), r)
        t.synchronize()
        del y
    if dist is not None:
        dist.barrier()
    t.reset_launch_count()
    if args.profile:
        t.profile(True)
    t0 = time.perf_counter()
    y = None
    for c in range(nchunks):
        del y
        y = t.tt_compress_(t.apply(Ad, xd), r)
    t.synchronize()
    el = time.perf_counter() - t0
    launches = t.launch_count()
    fam = None
    if args.profile:
        fam = {k: round(v[0], 1) for k, v in t.profile_read().items()}
        t.profile(False)
    if dist is not None:
        tm = torch.tensor([el], device="cuda", dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        el = float(tm.item())
    ok = None   # parity of the batched path is covered by tests/test_gpu_tt.py (the oracle is test infrastructure only)
    if rank == 0:
        done = nchunks * chunk * world
        per_vec_gflop = 13.3 * (d / 30.0)
        print(json.dumps({"metric": "batched apply+round vectors/s", "value": done / el, "unit": "vectors/s", "n_gpus": world,
                          "vectors": done, "chunk": chunk, "seconds": el, "d": d, "rank": r, "W": W, "dtype": "c128",
                          "gpu_launches": launches, "parity_rel_distance": ok, "family_ms": fam,
                          "approx_tflops": done * per_vec_gflop / el / 1e3}), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
