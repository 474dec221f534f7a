"""Multi-GPU plumbing for the sampling path: one process per GPU, contiguous batch shards, no
collective inside the sampling loop (samples are independent end to end), ONE all-gather of the finished frames.

`torch.distributed` (NCCL over NVLink/NVSwitch on the B200 box, gloo in the CPU tests) is used only for that
gather; rank-consistent noise is obtained by drawing the full-batch noise with the same seed on every rank and
slicing the local shard (`sharded_noise_fn`), which reproduces the single-GPU chain sample for sample."""
import torch
import torch.distributed as dist


def shard_bounds(global_batch, rank, world):
    """contiguous shard [lo, hi) of rank; remainders go to the first ranks"""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t, rank, world, dim=0):
    lo, hi = shard_bounds(t.shape[dim], rank, world)
    return t.narrow(dim, lo, hi - lo)


def sharded_noise_fn(global_batch, rank, world, seed, device_draw="cpu"):
    """noise_fn for GaussianDiffusion: every rank draws the full-batch tensor from an identically seeded generator
    and keeps its shard, so an N-GPU run consumes the same noise per sample as the 1-GPU run."""
    gen = torch.Generator(device=device_draw).manual_seed(seed)
    lo, hi = shard_bounds(global_batch, rank, world)

    def fn(shape, device):
        full = torch.randn((global_batch,) + tuple(shape[1:]), generator=gen, device=device_draw)
        return full[lo:hi].to(device)
    return fn


def gather_videos(local, group=None):
    """all-gather of per-rank (b_local, 3, F, H, W) frames into (sum b_local, 3, F, H, W) on every rank (C4)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    local = local.contiguous()
    sizes = [torch.zeros(1, dtype=torch.long, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], device=local.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    if len(set(sizes)) == 1:
        out = torch.empty((world * sizes[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], 0)
