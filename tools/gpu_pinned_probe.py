"""What does page-locked memory cost a one-shot caller? (GPU box; not a bench line)
Times cw_host_alloc (cold / from the cache) against np.empty + first touch, and one population-stage call with the
page-locked result / batch buffers against the same call on pageable arrays."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import cogsworth_b200 as cb
from cogsworth_b200 import _lib

ctx = cb.Context(devices=[0])
for mb in (64, 480, 960):
    n = mb * (1 << 20) // 8
    t0 = time.perf_counter(); a = _lib.pinned_empty(n); t1 = time.perf_counter(); a[:] = 0.0; t2 = time.perf_counter()
    del a
    t3 = time.perf_counter(); b = _lib.pinned_empty(n); t4 = time.perf_counter()
    del b
    t5 = time.perf_counter(); c = np.empty(n); c[:] = 0.0; t6 = time.perf_counter()
    del c
    print(f"{mb:5d} MB: cw_host_alloc cold {1e3 * (t1 - t0):7.1f} ms (+ first touch {1e3 * (t2 - t1):6.1f}), from the cache "
          f"{1e3 * (t4 - t3):6.2f} ms; np.empty + first touch {1e3 * (t6 - t5):6.1f} ms", flush=True)
