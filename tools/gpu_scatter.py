import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cogsworth_b200 as cb
ctx = cb.Context(devices=[0])
for streams in (500_000, 1_000_000, 6_000_000):
    for run in (64, -64, 128):
        print("streams", streams, "run_bytes", run, "GB/s %.0f" % ctx.measure_scatter_write(streams, 7680, run, 2048), flush=True)
