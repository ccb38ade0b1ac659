"""Sharding of the hot path over the GPUs of one box.

Every (file, chunk) is independent: the reference pads, frames and resamples each chunk on its own and merges
results by sorting on `start` (src/write/worker.py:83-87).  So the multi-GPU plan is a pure partition -- one
process per GPU, no collective on the data path.  Files are dealt to ranks longest-first onto the least-loaded
rank; when there are fewer files than ranks, the chunks of each file are split into contiguous ranges instead
(config 3: one 24 h file = 433 chunks over 8 GPUs)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class WorkItem:
    file_index: int
    chunk_index: int
    chunk: tuple            # (start_s, end_s)


def plan(files_chunklists: list[list[tuple]], world_size: int) -> list[list[WorkItem]]:
    """files_chunklists[i] = chunk list of file i.  Returns the work list of every rank (deterministic)."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    ranks: list[list[WorkItem]] = [[] for _ in range(world_size)]
    load = [0.0] * world_size
    n_files = len(files_chunklists)
    if n_files >= world_size:
        order = sorted(range(n_files), key=lambda i: (-sum(c[1] - c[0] for c in files_chunklists[i]), i))
        for i in order:
            r = min(range(world_size), key=lambda k: (load[k], k))
            for j, c in enumerate(files_chunklists[i]):
                ranks[r].append(WorkItem(i, j, (float(c[0]), float(c[1]))))
            load[r] += sum(c[1] - c[0] for c in files_chunklists[i])
    else:
        # fewer files than GPUs: contiguous chunk ranges of each file, ranks taken round-robin across files
        items = [WorkItem(i, j, (float(c[0]), float(c[1]))) for i, cl in enumerate(files_chunklists)
                 for j, c in enumerate(cl)]
        total = len(items)
        base, extra = divmod(total, world_size)
        pos = 0
        for r in range(world_size):
            take = base + (1 if r < extra else 0)
            ranks[r] = items[pos:pos + take]
            pos += take
    for r in range(world_size):
        ranks[r].sort(key=lambda w: (w.file_index, w.chunk_index))
    return ranks


def merge(per_rank_results: list[list[tuple]]) -> list[tuple]:
    """per_rank_results[r] = [(file_index, start_s, row...)]: what the single writer does -- sort by (file, start)."""
    rows = [row for part in per_rank_results for row in part]
    rows.sort(key=lambda t: (t[0], t[1]))
    return rows
