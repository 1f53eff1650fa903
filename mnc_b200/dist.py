"""Data-parallel sharding over images + one all-gather of per-image output records.

The path shards naturally: images are independent (the reference is batch-1 and has no
cross-image op), so there is no collective inside the forward.  One process
per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch on the B200 box; gloo in the CPU tests);
weights are replicated; the only exchange is a single all-gather of fixed-size records:
    [count | boxes 600x4 | scores 600x21 | masks 600x441]  fp32  (~1.12 MB per image).
"""
import os

import torch
import torch.distributed as dist

from .engine import ROIS_PER_IMAGE, MASK_SIZE, NUM_CLASSES

N_DET = 2 * ROIS_PER_IMAGE
REC_FLOATS = 1 + N_DET * 4 + N_DET * NUM_CLASSES + N_DET * MASK_SIZE * MASK_SIZE


def init_from_env(backend=None):
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).  Returns (rank, world, local)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_range(total, rank, world):
    """Contiguous block partition of `total` images: ranks [0, total % world) get one extra."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_records(boxes, masks, scores, valid):
    """(B,600,4), (B,600,1,21,21), (B,600,21), (B,600) -> (B, REC_FLOATS) fp32 send buffer."""
    B = boxes.shape[0]
    count = valid.view(B, -1).to(torch.float32).sum(dim=1, keepdim=True)
    return torch.cat([count, boxes.reshape(B, -1), scores.reshape(B, -1), masks.reshape(B, -1)],
                     dim=1).contiguous()


def unpack_records(rec):
    B = rec.shape[0]
    o = 1
    count = rec[:, 0].round().to(torch.int64)
    boxes = rec[:, o:o + N_DET * 4].view(B, N_DET, 4)
    o += N_DET * 4
    scores = rec[:, o:o + N_DET * NUM_CLASSES].view(B, N_DET, NUM_CLASSES)
    o += N_DET * NUM_CLASSES
    masks = rec[:, o:].view(B, N_DET, 1, MASK_SIZE, MASK_SIZE)
    return count, boxes, masks, scores


def all_gather_records(rec):
    """Every rank contributes (b, REC) and receives (world*b, REC), rank-major = image order for a
    contiguous shard.  Equal per-rank batch (the bench's weak-scaling layout)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return rec
    world = dist.get_world_size()
    out = torch.empty((world * rec.shape[0], rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec)
    return out
