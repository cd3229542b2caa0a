import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import haf_grasping_b200 as h
from haf_grasping_b200 import synth
import bench
F, R = bench.FEATURES, bench.RANGE
model = bench.model_path(2048)
n_clouds = 256
clouds = [synth.synth_cloud(1234 + i, 100000) for i in range(n_clouds)]
off = np.concatenate([[0], np.cumsum([len(c) for c in clouds])]).astype(np.int64)
host = torch.empty((int(off[-1]), 3), dtype=torch.float32, pin_memory=True)
host.numpy()[:] = np.concatenate(clouds)
dev = host.cuda()
gs = h.GraspSearch(F, R, model, svm_mode=h.HAF_SVM_TENSOR_GUARD)
st = torch.cuda.current_stream(); print("stream handle", st.cuda_stream)
gs.set_stream(st.cuda_stream)
rq = h.make_request()
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print("device-resident call ms", t(lambda: gs.search_batch_packed(dev, off, rq)))
print("host-buffer call ms", t(lambda: gs.search_batch_packed(host.numpy(), off, rq)))
side = torch.cuda.Stream()
def copy_only():
    with torch.cuda.stream(side):
        dev.copy_(host, non_blocking=True)
print("torch async copy only ms", t(copy_only))
def both():
    with torch.cuda.stream(side):
        dev2.copy_(host, non_blocking=True)
    gs.search_batch_packed(dev, off, rq)
dev2 = torch.empty_like(dev)
print("torch side-stream copy + device-resident call ms", t(both))
# host-side enqueue time of the host-buffer call
t0 = time.perf_counter(); gs.search_batch_packed(host.numpy(), off, rq); print("single host-buffer call wall ms", (time.perf_counter()-t0)*1e3)
# non-default stream
s2 = torch.cuda.Stream()
gs.set_stream(s2.cuda_stream)
with torch.cuda.stream(s2):
    print("host-buffer call on a non-default stream ms", t(lambda: gs.search_batch_packed(host.numpy(), off, rq)))
    print("device call on a non-default stream ms", t(lambda: gs.search_batch_packed(dev, off, rq)))
