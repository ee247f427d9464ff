"""Phase timestamps of k_linear_tc (build with ATMONR_NVCC_EXTRA=-DATM_LIN_TIMING python atmospheric-neural-rendering_b200/build.py):
mean SM-clock cycles per phase of a 128-row tile over 64 CTAs from the middle of the grid, for a 786 432 x 256 x 256 layer."""
import ctypes as C, os, sys, subprocess, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "atmospheric-neural-rendering_b200"))
from atmonr.native import lib as L, ops
L.load()
lib = C.CDLL(str(L.LIB_PATH))
m = 786432
x = torch.randn(m, 256, device="cuda")
w = torch.randn(256, 256, device="cuda") / 16
b = torch.randn(256, device="cuda")
bits = None
for name, fn in (("fwd relu+bits", lambda: ops.linear_forward(x, w, b, True, want_bits=True)),
                 ("dX bits", lambda: ops.linear_forward(x, w, None, False, transpose=True, out_bits=bits)),
                 ("plain", lambda: ops.linear_forward(x, w, None, False))):
    y = fn()
    if name.startswith("fwd"):
        bits = y[1]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    buf = (C.c_longlong * 512)()
    assert lib.atmonr_debug_lin_timing(buf) == 0
    t = torch.tensor(list(buf), dtype=torch.float64).view(64, 8)
    d = (t[:, 1:] - t[:, :-1]).mean(0)
    buf2 = (C.c_longlong * 512)()
    assert lib.atmonr_debug_lin_timing2(buf2) == 0
    t2 = torch.tensor(list(buf2), dtype=torch.float64).view(64, 8)[:, :7]
    d2 = (t2[:, 1:] - t2[:, :-1]).mean(0)
    print("   iteration c=5 (tid 0): wait stage free %.0f | wait loads + refetch %.0f | split+store %.0f | wait MMA(c-1) + issue B %.0f | fence+sync %.0f | wait B %.0f" % tuple(d2.tolist()))
    print(name, "kernel+prep ms %.3f" % e0.elapsed_time(e1), "phase cycles: prologue %.0f | first chunk %.0f | chunks1-3 %.0f | chunks4-7 %.0f | last MMA wait %.0f | epilogue %.0f | dealloc %.0f | total %.0f" % (*d.tolist(), float((t[:, 7] - t[:, 0]).mean())))
