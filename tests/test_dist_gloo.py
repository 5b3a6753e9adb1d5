"""World-size-2 gloo test of the multi-rank plumbing (CPU): env sharding + counter reduction."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from multi_car_racing_b200.dist import rank_world, shard_envs, reduce_throughput
    rank, world, _ = rank_world()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_envs(2049, rank, world)
    frames, ms = reduce_throughput(count * 2 * 10, 5.0 + rank)
    assert frames == 2049 * 2 * 10 and ms == 6.0, (frames, ms)
    dist.barrier()
    if rank == 0:
        print("GLOO_OK", first, count)
    dist.destroy_process_group()
''') % ROOT


def test_two_rank_counter_reduction(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "GLOO_OK 0 1025" in out.stdout
