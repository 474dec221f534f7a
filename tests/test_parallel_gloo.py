"""CPU: world_size-2 gloo test of the N>1 sharding / gather logic (the data path itself has no collective)."""
import os
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cvpr23_lfdm_b200 import parallel as PL
    gb = 5                                                  # ragged: shards of 3 and 2
    torch.manual_seed(0)
    imgs = torch.rand(gb, 3, 4, 4)
    mine = PL.shard(imgs, rank, world)
    nf = PL.sharded_noise_fn(gb, rank, world, seed=7)
    n0 = nf((mine.shape[0], 3, 2, 4, 4), "cpu")
    n1 = nf((mine.shape[0], 3, 2, 4, 4), "cpu")
    # stand-in for the (collective-free) sampling chain: any per-sample function of (input, noise)
    vid = mine[:, :, None] * 2 + n0 - n1
    full = PL.gather_videos(vid)
    if rank == 0:
        ret.put(full)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_match_single_process():
    from cvpr23_lfdm_b200 import parallel as PL
    assert [PL.shard_bounds(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    imgs = torch.rand(5, 3, 4, 4)
    g = torch.Generator().manual_seed(7)
    n0, n1 = torch.randn(5, 3, 2, 4, 4, generator=g), torch.randn(5, 3, 2, 4, 4, generator=g)
    assert torch.equal(got, imgs[:, :, None] * 2 + n0 - n1)
