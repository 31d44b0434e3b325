import torch, time
n = 5*1024**3//8
a = torch.empty(n, dtype=torch.float64, device='cuda'); b = torch.empty(n, dtype=torch.float64, device='cuda')
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
ms = t(lambda: a.fill_(1.0)); print("fill 5GB: %.3f ms -> %.0f GB/s written" % (ms, n*8/ms/1e6))
ms = t(lambda: b.copy_(a)); print("copy 5GB: %.3f ms -> %.0f GB/s read+write" % (ms, 2*n*8/ms/1e6))
ms = t(lambda: a.zero_()); print("memset 5GB: %.3f ms -> %.0f GB/s written" % (ms, n*8/ms/1e6))
