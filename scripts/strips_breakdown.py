#!/usr/bin/env python
"""Where a sort-first strip frame spends its time (run under torchrun): kernels vs gather vs host."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import gorender_b200 as g
from gorender_b200 import parallel, workloads

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H = 3840, 2160
objs, cam = workloads.config_c4(100)
stream = torch.cuda.Stream()
dev = g.Device(local, stream.cuda_stream)
with torch.cuda.stream(stream):
    tfb = parallel.TorchFrameBuffer(W, H, 1, dev, torch.device("cuda", local))
    r = g.Renderer(tfb.fb)
    packed = np.ascontiguousarray(r.pack_objects(objs, [cam]))
    def frame(gather=True):
        parallel.draw_strip(r, packed, H, world, rank)
        if gather and world > 1:
            parallel.gather_strips_to_rank0(tfb.color[0], tfb.depth[0], H)
    for _ in range(5): frame()
    torch.cuda.synchronize()
    N = 40
    res = {}
    for name, gather in (("draw+gather", True), ("draw only", False)):
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(N): frame(gather)
        torch.cuda.synchronize(); dev.synchronize()
        res[name] = (time.perf_counter() - t0) / N * 1e3
    dev.set_kernel_timing(True); dev.kernel_times()
    for _ in range(N): frame(False)
    dev.synchronize()
    kt, _ = dev.kernel_times()
    res["kernels"] = {k: v / N for k, v in kt.items() if v > 0.01}
if rank == 0:
    print(json.dumps({"world": world, "ms": res}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
