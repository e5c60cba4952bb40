"""Multi-GPU driver: independent video streams, one process per GPU.

The path is embarrassingly data-parallel, exactly like the reference's evaluation driver, which shards
`idxs[device_id::num_workers]` over one process per GPU with no communication
(REF/evaluation/livesports3kcc/distributed_generate_livecc.py:46-50,107-122). Here the processes are
torchrun ranks; torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is used only for the
launch barrier and the final gather of per-stream statistics — there is no data-path collective to fuse
with any kernel.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Sequence


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_streams(stream_ids: Sequence[int], rank: int, world: int) -> List[int]:
    """Static round-robin partition: stream i -> rank i mod world."""
    return list(stream_ids)[rank::world]


def init_distributed(backend: str = None, device=None):
    import torch
    import torch.distributed as dist

    rank, world, _ = dist_env()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def barrier():
    import torch
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def gather_stats(local: List[Dict]) -> List[Dict]:
    """All ranks receive the concatenated per-stream records (a few hundred bytes over NCCL/gloo)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(local)
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, list(local))
    return [r for part in out for r in part]


def run_streams(stream_ids: Sequence[int], run_one: Callable[[int], Dict]) -> List[Dict]:
    """barrier -> this rank's streams -> barrier -> gathered stats (sorted by stream id)."""
    rank, world, _ = dist_env()
    mine = shard_streams(stream_ids, rank, world)
    barrier()
    local = []
    for sid in mine:
        rec = dict(run_one(sid))
        rec["stream"] = sid
        rec["rank"] = rank
        local.append(rec)
    barrier()
    return sorted(gather_stats(local), key=lambda r: r["stream"])


def summarize(records: List[Dict], wall_s: float) -> Dict:
    """Whole-job aggregate: units summed over all streams / the slowest rank's time."""
    tok = sum(r.get("tokens", 0) for r in records)
    frames = sum(r.get("frames", 0) for r in records)
    return {"streams": len(records), "tokens": tok, "frames": frames, "tokens_per_s": tok / wall_s if wall_s else 0.0,
            "frames_per_s": frames / wall_s if wall_s else 0.0}
