"""Host <-> device copy bandwidth of the box from pinned memory (one direction, both directions at once, and while a kernel runs):
the denominator for the end-to-end leg of bench.py (2 x 10.4 GB per cfg5 step at N = 1)."""
import json
import time

import torch


def main():
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist = None
    if world > 1:                      # all ranks copy at the same time: what the host can feed to N GPUs at once
        import torch.distributed as dist
        dist.init_process_group("gloo")
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}

    def timed(fn, reps=10):
        fn()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d()
        d2h()

    out["h2d_GBps"] = n / timed(h2d) / 1e9
    out["d2h_GBps"] = n / timed(d2h) / 1e9
    out["duplex_GBps_each_way"] = n / timed(both) / 1e9
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.float64)

    def both_busy():
        both()
        torch.matmul(a, a)

    tb = timed(both_busy)
    tm = timed(lambda: torch.matmul(a, a))
    out["duplex_with_fp64_gemm_s"] = tb
    out["fp64_gemm_alone_s"] = tm
    out["rank"], out["world"], out["cpus"] = rank, world, len(os.sched_getaffinity(0))
    if dist is not None:
        allo = [None] * world
        dist.all_gather_object(allo, out)
        if rank == 0:
            print(json.dumps({"world": world, "cpus": out["cpus"],
                              "h2d_GBps_per_gpu": [round(o["h2d_GBps"], 1) for o in allo],
                              "d2h_GBps_per_gpu": [round(o["d2h_GBps"], 1) for o in allo],
                              "duplex_GBps_each_way_per_gpu": [round(o["duplex_GBps_each_way"], 1) for o in allo]}))
        dist.destroy_process_group()
        return
    print(json.dumps(out))


if __name__ == "__main__":
    main()
