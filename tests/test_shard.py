"""Multi-GPU path = partition only (no collective).  Host logic tested on CPU, N>1 with world_size-2 gloo."""
import os

import numpy as np
import pytest

from buzzdetect_b200 import shard, stream


def test_chunklength_rounding_matches_reference():
    assert stream.setup_chunklength(200) == 199.68          # src/analyze.py:102-111
    assert stream.setup_chunklength(0.5) == 0.96
    assert stream.setup_chunklength(1200) == 1199.04 or stream.setup_chunklength(1200) == 1200.0


def test_chunklist_and_sample_indexing():
    cl = stream.file_chunklist(3600.0, 199.68)
    assert len(cl) == 19 and cl[0] == (0.0, 199.68) and cl[-1][1] == 3600.0
    a, n = stream.chunk_sample_range(cl[1], 44100)
    assert a == int(199.68 * 44100) and n == int(cl[1][1] * 44100) - a
    # 24 h file -> 433 chunks
    assert len(stream.file_chunklist(86400.0, 199.68)) == 433


def test_resume_processes_only_the_gaps():
    """config 3: a partial result covering [0,6h) u [7h,12h) leaves exactly the gaps."""
    hop = 0.96
    starts = np.concatenate([np.arange(0, 6 * 3600, hop), np.arange(7 * 3600, 12 * 3600, hop)])
    starts = np.round(starts, 2)
    cl = stream.file_chunklist(86400.0, 199.68, covered_starts=starts)
    # float noise (start + 0.96 vs the next rounded start) makes melt_coverage report many hair-line "gaps", exactly as
    # the reference's pandas version does; smooth_gaps' tolerance (framelength/4) is what removes them again
    covered = stream.melt_coverage(starts, hop)
    assert len(covered) > 2
    assert cl[0][0] >= 6 * 3600 - 1 and cl[0][0] <= 6 * 3600 + 1
    assert all(not (c[0] >= 7 * 3600 + 1 and c[1] <= 12 * 3600 - 1) for c in cl)
    assert all(c[1] > c[0] for c in cl)
    total = sum(c[1] - c[0] for c in cl)
    assert abs(total - (3600 + 12 * 3600)) < 2.0


@pytest.mark.parametrize("n_files,world", [(1000, 8), (1000, 2), (3, 4), (1, 8), (8, 8), (0, 4)])
def test_plan_is_a_partition(n_files, world):
    rng = np.random.default_rng(n_files + world)
    files = [stream.file_chunklist(float(rng.integers(600, 7200)), 199.68) for _ in range(n_files)]
    ranks = shard.plan(files, world)
    seen = sorted((w.file_index, w.chunk_index) for r in ranks for w in r)
    want = sorted((i, j) for i, cl in enumerate(files) for j in range(len(cl)))
    assert seen == want                                          # every chunk exactly once
    if n_files >= world and n_files:
        load = [sum(w.chunk[1] - w.chunk[0] for w in r) for r in ranks]
        assert max(load) - min(load) <= 7200.0                   # longest-first dealing keeps ranks within one file
        for r in ranks:                                          # whole files stay on one rank
            assert len({w.file_index for w in r}) * 1 <= len(r)
    if n_files == 1 and world == 8:
        sizes = [len(r) for r in ranks]
        assert max(sizes) - min(sizes) <= 1


def _fake_predict(file_index, chunk):
    """Deterministic stand-in for the GPU path: one row per 0.96 s frame, value depends on (file, start)."""
    n = int(round((chunk[1] - chunk[0]) / 0.96))
    starts = np.round(np.arange(n) * 0.96 + chunk[0], 2)
    return [(file_index, float(s), float(np.sin(file_index + s))) for s in starts]


def _worker(rank, world, port, files, out_q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.plan(files, world)[rank]
    rows = [row for w in mine for row in _fake_predict(w.file_index, w.chunk)]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(rows, gathered, dst=0)                     # results go to the single writer (rank 0)
    if rank == 0:
        out_q.put(shard.merge(gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_equal_single_process():
    import torch.multiprocessing as mp
    files = [stream.file_chunklist(d, 199.68) for d in (1800.0, 950.4, 3600.0, 400.0, 2222.0)]
    single = shard.merge([[row for w in shard.plan(files, 1)[0] for row in _fake_predict(w.file_index, w.chunk)]])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, files, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert merged == single
