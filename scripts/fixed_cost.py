import os, sys, torch
sys.path.insert(0, "/root/repo")
import multipinn_b200 as mp
torch.manual_seed(0)
m = mp.FNN(2, 1, [64] * 9).cuda()
e = m.jet_engine
for n in (128, 148 * 128, 50_000, 200_000):
    for order in (0, 2):
        x = torch.rand(n, 2, device="cuda")
        out, stash = e.forward(x, order, True)
        dout = torch.randn_like(out)
        for _ in range(5):
            e.backward(x, order, dout, stash)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            e.backward(x, order, dout, stash)
        b.record(); torch.cuda.synchronize()
        tb = a.elapsed_time(b) / 50
        a.record()
        for _ in range(50):
            e.forward(x, order, True)
        b.record(); torch.cuda.synchronize()
        print(f"n={n:7d} order={order}  bwd {tb*1e3:8.1f} us   fwd {a.elapsed_time(b)/50*1e3:8.1f} us")
