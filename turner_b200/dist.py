"""Process-per-GPU plumbing for the sample-split render.

The path shards by pixel sample: rank g of G renders samples i = g (mod G) of every pixel with the scene replicated,
into its own W*H*4 float accumulation buffer; ONE reduce(sum) onto rank 0 ends the job. There is no other exchange.
torch.distributed carries the reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import os


def sample_split(rank, world_size):
    """(sample_begin, sample_stride) of this rank"""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return rank, world_size


def local_sample_count(pixel_samples, rank, world_size):
    begin, stride = sample_split(rank, world_size)
    return 0 if begin >= pixel_samples else (pixel_samples - begin + stride - 1) // stride


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend):
    import torch.distributed as dist
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def reduce_accum(accum, root=0):
    """sum the per-rank accumulation buffers onto `root` (in place); no-op for a single process"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=root, op=dist.ReduceOp.SUM)
    return accum
