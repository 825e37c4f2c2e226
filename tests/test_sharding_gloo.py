"""CPU, world_size 2 over gloo: the host-side multi-GPU logic of bench.py.  The path shards by whole
16-image batches (one arg-min group each) with NO data-path collective; torch.distributed
is only the barrier and the max-over-ranks of the timed region."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert bench.dist_env() == (rank, rank, world)
        # every rank owns its own batches: seeds (hence inputs) are disjoint across ranks
        mine = torch.tensor([bench.batch_seed(rank, b) for b in range(16)])
        allseeds = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allseeds, mine)
        flat = torch.cat(allseeds).tolist()
        assert len(set(flat)) == len(flat)
        x0, rel0, _ = bench.synthetic_batch(2, bench.SCALES, bench.batch_seed(rank, 0))
        g = [torch.zeros_like(rel0[0]) for _ in range(world)]
        dist.all_gather(g, rel0[0])
        assert not torch.equal(g[0], g[1])
        # timing reduction: max over ranks, after a barrier
        bench.dist_barrier()
        t = bench.dist_max(10.0 + rank)
        assert t == 10.0 + world - 1
        # whole-job value = all ranks' units / the max time (weak scaling)
        value = world * 100 * bench.BATCH / (t * 1e-3)
        q.put((rank, value))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_timing_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(2))
    assert res[0] == res[1] == 2 * 100 * 16 / 11e-3


def test_reference_arm_prints_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
    env["RANK"] = "0"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         env=env, capture_output=True, text=True, timeout=300)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "fused_depth_maps_per_sec" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.shared_config("fusion") and line["step"]["images_per_step_per_gpu"] == 16


def test_algorithmic_bytes_match_survey():
    sys.path.insert(0, ROOT)
    import bench
    ab = bench.algorithmic_bytes((8, 16, 32))
    # SURVEY.md 8d "Values, scales {8,16,32}"
    assert ab["pair"] == 679680 and ab["quantize"] == 1101824 and ab["als"] == 349440
    assert ab["decompose"] == 20872 and ab["reconstruct"] == 151516 and ab["path"] == 1623652
    assert bench.algorithmic_bytes((8, 16, 32, 64))["path"] == 6216612
    # the per-kernel split reported in config.kernel_gbs / roofline partitions the same contract bytes
    for scales in ((8, 16, 32), (8, 16, 32, 64)):
        ab = bench.algorithmic_bytes(scales)
        assert ab["sparsify_kernel"] + ab["als_sparse_kernel"] + ab["als_dense_kernel"] == ab["quantize"] + ab["als"]
        assert ab["tail_kernel"] + ab["quantize"] + ab["als"] == ab["path"]
    # both arms print the same `config` object (the driver compares the two lines)
    assert bench.shared_config("fusion") == {"workload": bench.shared_config("fusion")["workload"], "batch": 16, "scales": [8, 16, 32]}
