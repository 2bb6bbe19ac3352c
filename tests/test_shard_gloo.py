"""Host-side sharding logic of the multi-GPU path (ecseg_b200/shard.py), world_size 2 over gloo on
CPU: image i -> rank i mod N, rows gathered on rank 0, CSV identical to the single-process run."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ecseg_b200 import shard  # noqa: E402


def fake_count(path: str) -> int:
    """Stand-in for the GPU segment call: a deterministic function of the file name only."""
    return sum(map(ord, os.path.basename(path))) % 97


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, paths, out_dir: str):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seen = []

    def process_one(p):
        seen.append(p)
        return fake_count(p)

    rows = shard.run_sharded(paths, process_one, rank, world, lambda r: shard.dist_gather(r, rank, world))
    assert seen == [paths[i] for i in range(rank, len(paths), world)]
    if rank == 0:
        shard.write_csv(os.path.join(out_dir, "ec_quantification.csv"), rows)
    else:
        assert rows is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_indices_partition():
    for n in (0, 1, 7, 1024):
        for world in (1, 2, 4, 8):
            got = sorted(i for r in range(world) for i in shard.shard_indices(n, r, world))
            assert got == list(range(n))
    with pytest.raises(ValueError):
        shard.shard_indices(4, 2, 2)


def test_merge_detects_duplicates_and_gaps():
    with pytest.raises(ValueError):
        shard.merge_rows([[(0, "a", 1)], [(0, "a", 1)]], 1)
    with pytest.raises(ValueError):
        shard.merge_rows([[(0, "a", 1)]], 2)


def test_world2_gloo_csv_equals_single_process(tmp_path):
    paths = [f"/data/run/img_{i:03d}.tif" for i in range(11)] + ['/data/run/odd,name "x".tif']
    single = shard.run_sharded(paths, fake_count, 0, 1, lambda r: [r])
    ref_csv = tmp_path / "single.csv"
    shard.write_csv(str(ref_csv), single)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, paths, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ec_quantification.csv").read_text() == ref_csv.read_text()
    lines = ref_csv.read_text().splitlines()
    assert lines[0] == "image name,# of ec" and len(lines) == 1 + len(paths)
    assert lines[-1].startswith('"odd,name ""x"".tif",')


def test_process_batch_share_and_length_check():
    """The per-rank batch hook (the overlapped file pipeline on a GPU box) sees exactly this rank's share, in order."""
    paths = [f"/d/i{i}.tif" for i in range(9)]
    for world in (1, 2, 4):
        per_rank = []
        for rank in range(world):
            seen = []

            def batch(mine):
                seen.extend(mine)
                return [fake_count(p) for p in mine]

            def gather(rows):
                per_rank.append(rows)
                return None           # what every rank but 0 sees

            assert shard.run_sharded(paths, fake_count, rank, world, gather, batch) is None
            assert seen == paths[rank::world]
        assert shard.merge_rows(per_rank, 9) == [(os.path.basename(p), fake_count(p)) for p in paths]
    with pytest.raises(ValueError):
        shard.run_sharded(paths, fake_count, 0, 1, lambda r: [r], lambda mine: [1])
