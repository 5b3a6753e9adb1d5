"""Raw pinned device-to-host ceiling of this box: one plain cudaMemcpyAsync of a step's
observation bytes (B x A x 27 648 B) per iteration, per GPU, all ranks at once.  Run it the way bench.py is run:
    python scripts/micro/d2h_ceiling.py                                   (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/micro/d2h_ceiling.py
Prints GB/s per rank (min / max), the aggregate, and what that caps e2e agent-frames/s at."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
try:
    from multi_car_racing_b200.dist import bind_to_gpu_numa_node
    bind_to_gpu_numa_node(lr)
except Exception:
    pass
B, A = 1024, 2
nbytes = B * A * 96 * 96 * 3
src = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
for _ in range(5):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
iters = 50
t0 = time.perf_counter()
for _ in range(iters):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if world > 1:
    dist.barrier()
gbs = nbytes * iters / dt / 1e9
t = torch.tensor([gbs, gbs, -gbs], dtype=torch.float64, device="cuda")
if world > 1:
    mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    lo, hi, agg = -mx[2].item(), mx[1].item(), sm[0].item()
else:
    lo = hi = agg = gbs
if rank == 0:
    print(json.dumps({"n_gpus": world, "bytes_per_copy": nbytes, "d2h_GBs_per_gpu_min": lo, "d2h_GBs_per_gpu_max": hi, "d2h_GBs_aggregate": agg,
                      "e2e_ceiling_agent_frames_per_s": agg * 1e9 / (96 * 96 * 3), "note": "copy alone; a synchronous step adds its compute in front"}))
if world > 1:
    dist.destroy_process_group()
