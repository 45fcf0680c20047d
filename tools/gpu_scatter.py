"""Ceiling of the trajectory write pattern (no arithmetic): `streams` far-apart sequential streams appended in lock-step,
run_bytes at a time; cooperative (run_bytes / 16 adjacent lanes per run, STG.128) or one thread per stream (code 1128:
128-byte runs as four STG.256).  GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cogsworth_b200 as cb
ctx = cb.Context(devices=[0])
for streams, resident in ((6 * 37_888, 256), (6 * 75_776, 512), (6_000_000, 2048), (6_000_000, 512), (6_000_000, 256)):
    for run in (64, 128, 256, 1128):
        print("streams", streams, "resident threads/SM", resident, "run_bytes", run,
              "GB/s %.0f" % ctx.measure_scatter_write(streams, 7680, run, resident), flush=True)
