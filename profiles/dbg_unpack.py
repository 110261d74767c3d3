import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, progan_b200
from progan_b200 import _lib
K = progan_b200.get_kernels()
dev = "cuda"
def run(n, cin, cout, reps=10):
    ws = [torch.zeros(9 * cin * cout, device=dev) for _ in range(n)]
    dw = [torch.zeros(cout, cin, 3, 3, device=dev) for _ in range(n)]
    rows = [_lib.UnpackEntry(w.data_ptr(), d.data_ptr(), cin, cout, cin, cout, 9, 0, 0, 0, 1.0, 0.0) for w, d in zip(ws, dw)]
    tab = K._upload(rows, _lib.UnpackEntry, torch.device(dev))
    for _ in range(3):
        K._call("pg_wgrad_unpack_multi", tab.data_ptr(), n, K._stream())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        K._call("pg_wgrad_unpack_multi", tab.data_ptr(), n, K._stream())
    e1.record(); torch.cuda.synchronize()
    print("n=%d %dx%d: %.1f us per launch" % (n, cin, cout, e0.elapsed_time(e1) / reps * 1e3))
run(1, 128, 128); run(12, 128, 128); run(12, 32, 32); run(24, 128, 128)
