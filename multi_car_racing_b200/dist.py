"""Multi-GPU plumbing: the env batch shards across ranks with no data-path collective
(every reference env owns its own b2World, reference multi_car_racing.py:138); NCCL/gloo only
gathers throughput counters at report time."""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_envs(total_envs, rank, world):
    """Contiguous block of env indices owned by `rank`: (first, count).  All agents of an env stay
    on one rank (cars of one env interact)."""
    base, rem = divmod(int(total_envs), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def reduce_throughput(frames, elapsed_ms, device=None):
    """Whole-job numbers from per-rank counters: (sum of frames, max of elapsed).  Uses the
    initialised torch.distributed group (nccl on GPUs, gloo in the CPU tests); identity when the
    job is a single process."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(frames), float(elapsed_ms)
    t_sum = torch.tensor([float(frames)], dtype=torch.float64, device=device)
    t_max = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t_sum, op=dist.ReduceOp.SUM)
    dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    return float(t_sum.item()), float(t_max.item())


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """One process per GPU: pin this process to the CPUs of the NUMA node its GPU hangs off, so that
    the pinned staging buffers of step_host (allocated afterwards, first-touch local) and the copy
    threads sit next to the GPU's PCIe root.  With 8 ranks copying 56.6 MB of observations per step
    each, remote-node staging halves the aggregate device-to-host bandwidth.  Returns the node id,
    or None when the topology cannot be read (nothing is changed then).  MCR_NUMA_BIND=0 disables."""
    if os.environ.get("MCR_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = _parse_cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read())
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None
