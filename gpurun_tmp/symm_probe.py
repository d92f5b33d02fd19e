import os, sys, time
sys.path.insert(0, ".")
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
from fem_forth_perturb_b200 import _lib as L
n = 1 << 20
t = symm_mem.empty(n, dtype=torch.float64, device=f"cuda:{local}")
t.fill_(-1.0)
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "rendezvous ok", type(hdl).__name__, [m for m in dir(hdl) if not m.startswith("_")][:30], flush=True)
peer = (rank + 1) % world
pb = hdl.get_buffer(peer, (n,), torch.float64)
src = torch.arange(n, dtype=torch.float64, device=f"cuda:{local}") + 1000 * rank
idx = torch.arange(0, 4096, dtype=torch.int32, device=f"cuda:{local}")
hdl.barrier(channel=0)
# push 4096 gathered values into the peer's buffer at offset 128 with MY kernel (peer pointer)
import ctypes
dst = ctypes.c_void_p(pb.data_ptr() + 128 * 8)
L.check(L.lib().mwx_gather_f64(4096, L.ptr(src), L.ptr(idx), dst, L.stream()))
hdl.barrier(channel=0)
torch.cuda.synchronize()
exp = torch.arange(4096, dtype=torch.float64, device=f"cuda:{local}") + 1000 * ((rank - 1) % world)
ok = bool(torch.equal(t[128:128 + 4096], exp)) and float(t[0]) == -1.0
# latency of push + barrier
torch.cuda.synchronize(); dist.barrier()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    L.lib().mwx_gather_f64(4096, L.ptr(src), L.ptr(idx), dst, L.stream())
    hdl.barrier(channel=0)
e1.record(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    L.lib().mwx_gather_f64(4096, L.ptr(src), L.ptr(idx), dst, L.stream())
    hdl.barrier(channel=0)
torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 200
print(f"RANK{rank} symm push ok={ok} us_per_halo_device={e0.elapsed_time(e1) * 1000 / 200:.1f} wall_us={wall * 1e6:.1f}", flush=True)
# compare: NCCL p2p
buf = torch.empty(4096, dtype=torch.float64, device=f"cuda:{local}"); rbuf = torch.empty_like(buf)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(200):
    ops = [dist.P2POp(dist.irecv, rbuf, (rank - 1) % world), dist.P2POp(dist.isend, buf, peer)]
    for w in dist.batch_isend_irecv(ops): w.wait()
torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 200
print(f"RANK{rank} nccl p2p wall_us={wall * 1e6:.1f}", flush=True)
dist.destroy_process_group()
