"""scripts/probe_symm.py -- run under torchrun: does torch symmetric memory (peer-mapped buffers + barrier) work on this box, and
what does its barrier cost?  (The answer decided bench.py's frame assembly at N > 1: PeerFrames in distributed.py.)"""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    t = symm_mem.empty((1 << 20,), dtype=torch.uint8, device=torch.device("cuda", local))
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast", getattr(hdl, "multicast_ptr", None), flush=True)
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (16,), torch.uint8)
    print(rank, "peer first bytes", peer[:4].tolist(), flush=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5): hdl.barrier()
    e0.record()
    for _ in range(100): hdl.barrier()
    e1.record(); torch.cuda.synchronize()
    print(rank, "barrier us", e0.elapsed_time(e1) * 10, flush=True)
    x = torch.zeros(1, device="cuda")
    for _ in range(5): dist.all_reduce(x)
    e0.record()
    for _ in range(100): dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    print(rank, "nccl 1-elem all_reduce us", e0.elapsed_time(e1) * 10, flush=True)
    a = torch.zeros(6220800 // world, dtype=torch.uint8, device="cuda"); g = torch.zeros(6220800, dtype=torch.uint8, device="cuda")
    for _ in range(5): dist.all_gather_into_tensor(g, a)
    e0.record()
    for _ in range(100): dist.all_gather_into_tensor(g, a)
    e1.record(); torch.cuda.synchronize()
    print(rank, "nccl all_gather 6.2MB us", e0.elapsed_time(e1) * 10, flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "SYMM FAILED", repr(e)[:300], flush=True)
dist.destroy_process_group()
