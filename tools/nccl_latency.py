"""Micro-benchmark (debug aid): latency of the two collectives of a map-sharded solve on this box."""
import os, time, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ws = dist.get_world_size(); dev = torch.device("cuda", local)
N, M, T = 8192, 256, 128
send = torch.randn(N * (M // ws), device=dev); recv = torch.empty_like(send)
part = torch.randn(2 * T + 2, device=dev); gath = torch.empty(ws * (2 * T + 2), device=dev)
def bench(fn, iters=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3, (time.perf_counter() - t0) / iters * 1e6
a = bench(lambda: dist.all_to_all_single(recv, send))
b = bench(lambda: dist.all_gather_into_tensor(gath, part))
if dist.get_rank() == 0:
    print("ws %d: all_to_all_single %d KB/rank: %.1f us (gpu) %.1f us (wall) | all_gather %d B: %.1f us (gpu) %.1f us (wall)"
          % (ws, send.numel() * 4 // 1024, a[0], a[1], part.numel() * 4, b[0], b[1]))
dist.destroy_process_group()
