import os, sys, time, numpy as np
sys.path.insert(0, '/root/repo')
os.environ['DUNK_TRACE'] = '1'
import cubesat_apds_b200 as dunk
from cubesat_apds_b200 import _lib
import bench
ctx = dunk.Context(0, 4)
lib = _lib.load()
B = 256
frames = bench.config2_frames(B)
cap = 8192
kps = np.zeros((B, cap), dtype=_lib.KEYPOINT_DTYPE); desc = np.zeros((B, cap, 61), np.uint8); counts = np.zeros(B, np.int32)
for it in range(3):
    t = time.perf_counter()
    _lib.check(lib.dunk_akaze_extract_batch(ctx.handle, frames.ctypes.data, B, 1024, 1024, 1, 1024, 1024 * 1024, 0, kps.ctypes.data, desc.ctypes.data, cap, counts.ctypes.data))
    print('call', it, (time.perf_counter() - t) * 1e3, 'ms', file=sys.stderr)
# raw memcpy rates on this host
a = np.zeros(64 << 20, np.uint8); b = np.ones(64 << 20, np.uint8)
for _ in range(3):
    t = time.perf_counter(); a[:] = b; dt = time.perf_counter() - t
print('numpy 64 MB copy: %.1f GB/s' % (0.064 / dt), file=sys.stderr)
ctx.close()
