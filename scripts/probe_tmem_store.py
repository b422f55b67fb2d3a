"""Development aid: print where tcgen05.st shapes put each thread's registers (see mg_debug_tmem_store)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from musicgan_b200 import build
l = ctypes.CDLL(build.DEBUG_LIB)
l.mg_debug_tmem_store.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
for shape, lane_off, col_off in [(0, 0, 0), (1, 0, 0), (2, 0, 0), (3, 0, 0), (1, 16, 4), (2, 16, 8)]:
    out = th.zeros(128, 32, dtype=th.int32, device="cuda")
    rc = l.mg_debug_tmem_store(out.data_ptr(), shape, lane_off, col_off, None)
    th.cuda.synchronize()
    o = out.cpu()
    print(f"shape {shape} lane_off {lane_off} col_off {col_off} rc {rc}")
    nz = o.nonzero()
    for ln in sorted(set(nz[:, 0].tolist())):
        cols = [(c, int(o[ln, c])) for c in range(32) if o[ln, c] != 0]
        print(f"  lane {ln:3d}: " + " ".join(f"c{c}=t{(v >> 4) & 0xFF}r{v & 0xF}" for c, v in cols))
