"""Runs the cfg4 two-site matvec a few times (target of the ncu captures of the DMMA GEMM)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import ttn_b200 as t  # noqa: E402
import bench  # noqa: E402

t.synchronize()
stream = torch.cuda.ExternalStream(t.stream_handle())
peak = json.load(open(bench.FP64_PEAK_FILE))
print(json.dumps(bench.bench_matvec(t, torch, stream, peak, reps=int(sys.argv[1]) if len(sys.argv) > 1 else 3)))
