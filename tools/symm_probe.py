"""Probe: does torch symmetric memory (P2P over NVLink) work on this box? torchrun --nproc-per-node 2."""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
try:
    t = symm.empty(1 << 20, dtype=torch.uint8, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok", type(h).__name__, [hex(p) for p in h.buffer_ptrs][:2], "signal pads", len(h.signal_pad_ptrs), flush=True)
    t.fill_(rank + 1)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (1 << 20,), torch.uint8)
    print(rank, "peer value", int(peer[12345].item()), flush=True)
    # timing of barrier
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(100): h.barrier()
    torch.cuda.synchronize(); print(rank, "barrier us", (time.perf_counter() - t0) * 1e4, flush=True)
    mine = torch.full((1 << 20,), 7, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): peer.copy_(mine)
    e1.record(); torch.cuda.synchronize(); print(rank, "p2p write 1MB us", e0.elapsed_time(e1) / 20 * 1e3, flush=True)
except Exception as e:
    print(rank, "FAILED", type(e).__name__, str(e)[:400], flush=True)
dist.destroy_process_group()
