"""Frame sharding for multi-GPU rendering: one process per GPU, contiguous frame blocks, no collective on the
data path (every frame of this path is independent).  Output order is by construction:
rank r owns frames [r*block, (r+1)*block), so concatenating the ranks' outputs in rank order is frame order."""
from __future__ import annotations

from typing import Iterator, List, Tuple


def block_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, end) of the contiguous block owned by `rank`: block = ceil(n_frames / world)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    block = -(-n_frames // world)
    start = min(rank * block, n_frames)
    return start, min(start + block, n_frames)


def batches(start: int, end: int, batch: int) -> Iterator[Tuple[int, int]]:
    """Consecutive [i0, i1) batches covering [start, end); the last one may be short."""
    i = start
    while i < end:
        yield i, min(i + batch, end)
        i += batch


def owner_of(frame: int, n_frames: int, world: int) -> int:
    block = -(-n_frames // world)
    return min(frame // block, world - 1)


def render_block(clip, rank: int, world: int) -> List[Tuple[int, object]]:
    """Render this rank's block of `clip` (any object with num_frames / get_frame) -> [(frame_no, frame), ...]."""
    s, e = block_range(clip.num_frames, rank, world)
    return [(n, clip.get_frame(n)) for n in range(s, e)]
