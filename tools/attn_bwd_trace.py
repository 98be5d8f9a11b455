"""Phase stamps (SM clocks, CTA 0, thread 0, last passage of the CTA) of the persistent attention backward; needs the
-DLAKO_TRACE build: tools/build_trace.sh; LAKO_B200_LIB=.../liblako_b200_trace.so python tools/attn_bwd_trace.py"""
import ctypes, os, sys, runpy
sys.argv = [sys.argv[0]]
ns = runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "attn_bwd_time.py"))
lib = ns["lib"]
buf = (ctypes.c_longlong * 64)()
assert lib.lako_debug_eb_trace(buf) == 0
t = list(buf)
t0 = t[3]
rel = lambda x: (x - t0) if x > 0 else -1
print("passage start 0, staged", rel(t[7]), " loops done", rel(t[4]))
for b in range(4):
    print(f"  block {b} (kt={b >> 1}, qt={b & 1}): S/dP ready {rel(t[8 + b * 6])}, P/dS stored {rel(t[8 + b * 6 + 2])}, end {rel(t[8 + b * 6 + 5])}")
print("kernel: start", rel(t[0]), "setup done", rel(t[2]), "flush", rel(t[5]), "end", rel(t[6]))
