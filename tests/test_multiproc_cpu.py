"""world_size-2 gloo tests (CPU) of the N>1 host logic of bench.py: per-rank seeding of the
batch shard, MAX-over-ranks timing, whole-job throughput, and the reference arm's
"rank 0 alone runs" rule."""
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import bench
    from cgg_b200 import synth
    dist.init_process_group('gloo', rank=rank, world_size=world)
    assert bench.dist_env() == (rank, rank, world)
    # each rank draws its own shard of the global batch (per-rank seed), no collective on the data path
    mf, _ = synth.make_inputs(rank, 1, 64, 64)
    ms_local = 10.0 + 5.0 * rank                       # rank 1 is the slow one
    ms = bench.reduce_max(ms_local, world, torch.device('cpu'))
    val = bench.whole_job_value(world, 16, 10, ms)
    q.put((rank, float(mf.sum()), ms, val))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_timing_and_sharding():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, ms0, v0), (r1, s1, ms1, v1) = res
    assert s0 != s1                                   # different shards
    assert ms0 == ms1 == 15.0                         # MAX over ranks
    assert abs(v0 - 2 * 16 * 10 / 15e-3) < 1e-6 and v0 == v1


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK='1', LOCAL_RANK='1', WORLD_SIZE='2')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                        '--warmup', '0'], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_flop_model_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert abs(bench.flops_per_image(100) / 1e9 - 60.57) < 0.05
    assert abs(bench.flops_per_image(200) / 1e9 - 104.4) < 0.1
