"""All-reduce cost of the flat gradient buffer (12.8 MB fp32 at 128^3) under torchrun: eager and inside a CUDA graph."""
import os, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for nfl in (3_200_000, 2_400_000, 800_000, 200_000):
    x = torch.randn(nfl, device="cuda")
    e = timed(lambda: dist.all_reduce(x))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        dist.all_reduce(x)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        dist.all_reduce(x)
    gt = timed(g.replay)
    if dist.get_rank() == 0:
        print(f"all_reduce {nfl * 4 / 1e6:.1f} MB x{dist.get_world_size()}: eager {e:.1f} us, graph {gt:.1f} us", flush=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
