"""Probe (torchrun, N >= 2): torch symmetric memory on this box — peer-memory writes + stream-ordered signals, no collective."""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
print(rank, "backend", symm.get_backend(dev), "nvshmem", symm.is_nvshmem_available(), flush=True)
n = 1 << 20
buf = symm.empty(world * n, dtype=torch.float32, device=dev)
buf.zero_()
hdl = symm.rendezvous(buf, dist.group.WORLD)
print(rank, "rendezvous ok; signal pad bytes", hdl.signal_pad_size, "world", hdl.world_size, flush=True)
torch.cuda.synchronize(); dist.barrier()
# every rank writes its slice into every peer's buffer, then signals channel 3
src = torch.full((n,), float(rank + 1), device=dev)
t0 = time.perf_counter()
for d in range(world):
    peer = hdl.get_buffer(d, (world * n,), torch.float32, 0)
    peer.narrow(0, rank * n, n).copy_(src)
    hdl.put_signal(d, 3, 10000)
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for s in range(world):
        hdl.wait_signal(s, 3, 10000)
    got = buf.view(world, n)[:, ::4096].clone()
side.synchronize()
torch.cuda.synchronize()
ok = all(bool((got[s] == s + 1).all()) for s in range(world))
print(rank, "peer writes + signals:", "OK" if ok else "MISMATCH", f"{(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
# bandwidth of a large peer copy
big = symm.empty(64 << 20, dtype=torch.float32, device=dev)
hb = symm.rendezvous(big, dist.group.WORLD)
x = torch.randn(64 << 20, device=dev)
peer = hb.get_buffer((rank + 1) % world, (64 << 20,), torch.float32, 0)
for _ in range(2): peer.copy_(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); peer.copy_(x); e1.record(); torch.cuda.synchronize()
print(rank, f"256 MB peer copy: {256e6 / (e0.elapsed_time(e1) * 1e-3) / 1e9:.0f} GB/s", flush=True)
dist.barrier(); dist.destroy_process_group()
