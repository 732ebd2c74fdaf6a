import torch, time
torch.cuda.set_device(0)
n = 1350451200 // 8
x = torch.empty(n, dtype=torch.float64, device='cuda'); y = torch.empty_like(x)
def t(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: x.fill_(1.5)); print("fill_   %.3f ms  %.0f GB/s (write only)" % (ms, n*8/ms/1e6))
ms = t(lambda: x.zero_());    print("zero_   %.3f ms  %.0f GB/s (memset)" % (ms, n*8/ms/1e6))
ms = t(lambda: y.copy_(x));   print("copy_   %.3f ms  %.0f GB/s (read+write)" % (ms, 2*n*8/ms/1e6))
ms = t(lambda: x.sum());      print("sum     %.3f ms  %.0f GB/s (read only)" % (ms, n*8/ms/1e6))
