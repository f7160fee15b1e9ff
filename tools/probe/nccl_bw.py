"""All-gather bandwidth probe (torchrun): what the time-sharded exchange can expect from this box's NCCL transport."""
import os, time, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mb in (8, 64, 400):
    n = mb * 1024 * 1024 // 8
    send = torch.ones(n, dtype=torch.float64, device="cuda")
    out = torch.empty(world * n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        dist.all_gather_into_tensor(out, send)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dist.all_gather_into_tensor(out, send)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if dist.get_rank() == 0:
        print(f"all_gather {mb} MB per rank x {world}: {ms:.2f} ms -> {(world - 1) * mb / 1024 / (ms * 1e-3):.1f} GB/s received per rank", flush=True)
dist.destroy_process_group()
