import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import dvode_b200 as B
from dvode_b200 import batches
dev = torch.device("cuda", 0); st = torch.cuda.current_stream()
for lg in (14, 16):
    n = 1 << lg
    s = B.dvode_t().initialize("diffusion48", n, batches.diffusion48_batch(n))
    y0d = torch.zeros(48, n, dtype=torch.float64, device=dev); yd = torch.empty_like(y0d)
    td = torch.zeros(n, dtype=torch.float64, device=dev); isd = torch.ones(n, dtype=torch.int32, device=dev)
    iw = np.zeros(30, dtype=np.int32); iw[5] = 20000
    for r in range(3):
        yd.copy_(y0d); td.zero_(); isd.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        s.solve_schedule_device(48, yd.data_ptr(), td.data_ptr(), batches.diffusion48_touts(), 1, [1e-6], [1e-9], 1, isd.data_ptr(), 1, np.zeros(20), iw, 22, 0, st.cuda_stream)
        e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("diffusion48 mf=22 2^%d: %.1f ms, %.0f systems/s, ok=%d" % (lg, ms, n / ms * 1e3, int((isd == 2).sum())))
    s.close()
