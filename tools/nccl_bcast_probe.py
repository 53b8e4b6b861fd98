"""2-GPU probe: latency of one 4.2 MB ncclBroadcast (the reconstructed-reference exchange), default vs side stream."""
import os, time, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = torch.zeros(1920 * 1088 * 2, dtype=torch.uint8, device="cuda")
side = torch.cuda.Stream()
for name, st in (("default", torch.cuda.current_stream()), ("side", side)):
    with torch.cuda.stream(st):
        for _ in range(5): dist.broadcast(t, src=0)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter(); e0.record(st)
        for _ in range(20): dist.broadcast(t, src=0)
        e1.record(st); e1.synchronize(); w1 = time.perf_counter()
    if dist.get_rank() == 0: print(f"{name}: {e0.elapsed_time(e1)/20*1e3:.1f} us/bcast (events), {(w1-w0)/20*1e6:.1f} us wall")
dist.destroy_process_group()
