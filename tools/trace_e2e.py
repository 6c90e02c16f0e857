import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from xmipp3_b200._lib import Reconstructor, make_particles
dev = torch.device("cuda", 0)
B, box = 4096, 256
img, cols = bench.synth_batch_torch(B, box, 0, dev, ctf=True)
p = make_particles(B, **cols)
hbuf = torch.empty(img.shape, dtype=torch.float32, pin_memory=True); hbuf.copy_(img); torch.cuda.synchronize()
r = Reconstructor(box, use_ctf=True, sampling=1.5, max_batch=1024)
for _ in range(2):
    r.insert_host_ptr(hbuf.data_ptr(), p); r.weight_sum()
r.sync()
r.timer_start()
t0 = time.perf_counter()
r.insert_host_ptr(hbuf.data_ptr(), p)
t1 = time.perf_counter()
r.weight_sum()
t2 = time.perf_counter()
print("insert call %.2f ms, result call %.2f ms" % (1e3*(t1-t0), 1e3*(t2-t1)), file=sys.stderr)
r.sync()
