import torch, time, sys
sys.path.insert(0, ".")
from lanpaint_b200 import hostnoise
dev = torch.device("cuda:0")
for shape in ((1,4,128,128),(8,4,128,128),(128,4,128,128)):
    n=1
    for d in shape: n*=d
    hostnoise._draw(n, 1, dev); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(5): hostnoise._draw(n, 2+k, dev)
    e1.record(); torch.cuda.synchronize()
    t0=time.perf_counter(); hostnoise.torch_cpu_randn(shape, 9, dev); torch.cuda.synchronize(); w=time.perf_counter()-t0
    print(shape, "device ms per draw", e0.elapsed_time(e1)/5, "wall incl. generator bookkeeping ms", 1e3*w)
