"""cta_group::1 UMMA rate (SS operands in smem, no-swizzle K-major): cycles per M128 x N x K16 instruction."""
import sys, ctypes as C
import numpy as np
sys.path.insert(0, ".")
from mapf_gpt_b200 import _lib
L = _lib.lib()
for ctas in (1, 148, 296):
    for N in (48, 80, 128, 160, 256, 1048, 1080, 1128, 1160):   # +1000: TS form (A operand in tensor memory)
        out = np.zeros(2, np.int64)
        rc = L.mg_test_umma_rate(0, N, 2000, ctas, out.ctypes.data_as(C.c_void_p))
        ts = N >= 1000
        N = N % 1000
        if rc:
            print(N, "ERR", L.mg_last_error().decode()); continue
        n = 2000 * 4
        print(f"ctas={ctas:3d} {'TS' if ts else 'SS'} N={N:3d}: {out[0] / n:7.1f} cyc/UMMA to completion, {out[1] / n:7.1f} cyc/UMMA issue loop "
              f"(ideal {N / 2:.0f}); smem operand bytes/UMMA {4096 + N * 32} -> {(4096 + N * 32) / (out[0] / n):.0f} B/clk", flush=True)
