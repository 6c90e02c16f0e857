"""Stage timeline of the end-to-end host loop of bench.py (pinned host batches, pipelined per-step read-back): run with
RFB200_TRACE=1 and the library prints every stage's start / end on its stream relative to timer_start; this script then
lists the idle gaps of the compute stream.  usage: RFB200_TRACE=1 python tools/trace_e2e.py [steps] 2> trace.txt"""
import os
import re
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from xmipp3_b200._lib import Reconstructor, make_particles


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    dev = torch.device("cuda", 0)
    B, box = 4096, 256
    host = []
    for s in range(2):
        img, cols = bench.synth_batch_torch(B, box, s, dev, ctf=True)
        hbuf = torch.empty(img.shape, dtype=torch.float32, pin_memory=True)
        hbuf.copy_(img)
        host.append((hbuf, make_particles(B, **cols)))
    torch.cuda.synchronize()
    r = Reconstructor(box, use_ctf=True, sampling=1.5, max_batch=B)
    for i in range(2):
        r.insert_host_ptr(host[i][0].data_ptr(), host[i][1])
        r.weight_sum()
    r.sync()
    r.reset()
    r.timer_start()
    t0 = time.perf_counter()
    marks = []
    for i in range(K):
        hbuf, p = host[i & 1]
        ta = time.perf_counter()
        r.insert_host_ptr(hbuf.data_ptr(), p)
        tb = time.perf_counter()
        if i > 0:
            r.weight_sum_end()
        r.weight_sum_begin()
        tc = time.perf_counter()
        marks.append((1e3 * (ta - t0), 1e3 * (tb - t0), 1e3 * (tc - t0)))
    r.weight_sum_end()
    t1 = time.perf_counter()
    r.sync()
    r.timings()
    print("host: %d steps in %.2f ms (%.2f ms per step)" % (K, 1e3 * (t1 - t0), 1e3 * (t1 - t0) / K))
    for i, m in enumerate(marks):
        print("host step %d: insert call %.2f -> %.2f ms, result calls until %.2f ms" % ((i,) + m))


def gaps(path):
    ev = []
    for line in open(path):
        m = re.match(r"\[rfb200 trace\] (\w+)\s+([\d.]+) ->\s+([\d.]+) ms", line)
        if m:
            ev.append((float(m.group(2)), float(m.group(3)), m.group(1)))
    comp = sorted(e for e in ev if e[2] != "h2d")
    idle, last = 0.0, None
    for a, b, name in comp:
        if last is not None and a - last > 0.05:
            print("gap %.3f ms before %-7s at %.3f ms" % (a - last, name, a))
            idle += a - last
        last = max(last or b, b)
    print("compute stream: %d stages, span %.2f ms, idle %.2f ms" % (len(comp), comp[-1][1] - comp[0][0], idle))
    h2d = sorted(e for e in ev if e[2] == "h2d")
    print("h2d: %d copies, %.2f ms busy" % (len(h2d), sum(b - a for a, b, _ in h2d)))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "gaps":
        gaps(sys.argv[2])
    else:
        main()
