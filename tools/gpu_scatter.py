import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cogsworth_b200 as cb
ctx = cb.Context(devices=[0])
for thr in (512, 1024, 2048):
    for run in (16, 32, 64, 128, 256, 512):
        # 1.2M streams x 8192 B = 9.8 GB
        print("threads/SM", thr, "run_bytes", run, "GB/s %.0f" % ctx.measure_scatter_write(1_200_000, 8192, run, thr))
