"""Probe: how much of the cfg2 step is launch gaps?  Times the step eagerly and as a replayed CUDA graph
(the replay bakes host scalars such as the Adam step count, so it is a timing probe only)."""
import sys, time, torch
sys.path.insert(0, ".")
import bench
from lshm_b200 import synthetic as S
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
step, opt = bench.build_step(dev, 0, 1, False)
Np = bench.CFG["baselines_per_gpu"] * bench.CFG["patches_per_baseline"]
x = torch.from_numpy(S.make_patches(Np, 8, seed=1)).to(dev)
uv = torch.from_numpy(S.make_uv(Np, seed=0, per_group=4)).to(dev)
step.set_batch(x, uv, 4, global_patches=Np)
def one_step():
    opt.step(step.closure); step.update_multipliers()
for _ in range(3): one_step()
torch.cuda.synchronize()
def timeit(fn, n=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("eager ms/step", timeit(one_step))
t0 = time.perf_counter(); one_step(); t1 = time.perf_counter()
print("host time to enqueue one step (ms)", (t1 - t0) * 1e3)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    one_step()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        one_step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
print("graph ms/step", timeit(g.replay))
