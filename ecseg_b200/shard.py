"""Multi-GPU `metaseg`: one process per GPU, images sharded by index, no collective on the data path.

Images are independent in the reference (the loop at src/metaseg.py:42 carries no state except the
result table), so image i goes to rank i mod N; every rank writes its own artefacts
(labels/<stem>.png, labels/<stem>.npy, dapi/<name>) and only the per-image (name, count) rows travel:
they are gathered on rank 0 over the host-side process group and written to ec_quantification.csv
in the reference's listing order, so the CSV does not depend on the GPU count.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        -m ecseg_b200.shard            # reads ./config.yaml like `make metaseg`

NVLink / NCCL are deliberately unused for data: the exchange is a few bytes per image.
"""
from __future__ import annotations

import os
import sys
from typing import Callable, Iterable, Sequence


def shard_indices(n_items: int, rank: int, world: int) -> list[int]:
    """Indices of the images rank `rank` of `world` processes (round robin: i mod world == rank)."""
    if not 0 <= rank < world:
        raise ValueError("rank must be in [0, world)")
    return list(range(rank, n_items, world))


def merge_rows(per_rank: Sequence[Iterable[tuple[int, str, int]]], n_items: int) -> list[tuple[str, int]]:
    """Per-rank [(index, name, count)] -> [(name, count)] in listing order; every index exactly once."""
    out: list = [None] * n_items
    for rows in per_rank:
        for idx, name, count in rows:
            if out[idx] is not None:
                raise ValueError(f"image {idx} was processed twice")
            out[idx] = (name, int(count))
    missing = [i for i, r in enumerate(out) if r is None]
    if missing:
        raise ValueError(f"images {missing[:8]} were not processed by any rank")
    return out


def write_csv(csv_path: str, rows: Sequence[tuple[str, int]]) -> None:
    """ec_quantification.csv exactly as pandas.to_csv(index=False) wrote it (src/metaseg.py:56-57)."""
    with open(csv_path, 'w') as f:
        f.write('image name,# of ec\n')
        for name, n in rows:
            name = '"' + name.replace('"', '""') + '"' if (',' in name or '"' in name) else name
            f.write(f'{name},{n}\n')


def run_sharded(image_paths: Sequence[str], process_one: Callable[[str], int], rank: int, world: int,
                gather: Callable[[list], list | None],
                process_batch: Callable[[list], list] | None = None) -> list[tuple[str, int]] | None:
    """Process this rank's share with `process_one(path) -> ecDNA count` (or, when given, the whole share at once
    with `process_batch(paths) -> counts`, which lets decode / GPU / writes overlap), gather the rows with
    `gather(rows)` (returns the list of all ranks' rows on rank 0, None elsewhere) and merge."""
    idx = shard_indices(len(image_paths), rank, world)
    if process_batch is not None:
        counts = list(process_batch([image_paths[i] for i in idx]))
        if len(counts) != len(idx):
            raise ValueError("process_batch must return one count per path")
    else:
        counts = [process_one(image_paths[i]) for i in idx]
    mine = [(i, os.path.split(image_paths[i])[1], int(n)) for i, n in zip(idx, counts)]
    everyone = gather(mine)
    if everyone is None:
        return None
    return merge_rows(everyone, len(image_paths))


def dist_gather(rows: list, rank: int, world: int):
    """Gather python rows on rank 0 over torch.distributed (gloo: host memory only)."""
    if world == 1:
        return [rows]
    import torch.distributed as dist
    out = [None] * world if rank == 0 else None
    dist.gather_object(rows, out, dst=0)
    return out


def main(argv=None) -> int:
    import numpy as np
    import yaml
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    var = yaml.load(open("config.yaml"), Loader=yaml.FullLoader)['metaseg']
    inpath = var['inpath']
    if not os.path.isdir(inpath):
        if rank == 0:
            print("Input folder does not exist. Exiting...")
        return 2
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
    torch.cuda.set_device(local)

    from . import metaseg as ms
    from .utils import allow_random_weights, get_imgs, load_model, meta_segment

    for sub in ('dapi', 'labels'):
        os.makedirs(os.path.join(inpath, sub), exist_ok=True)
    # one listing for everybody: glob order is the row order of the CSV
    paths = [get_imgs(inpath)] if rank == 0 else [None]
    if world > 1:
        dist.broadcast_object_list(paths, src=0)
    paths = paths[0]
    if not paths:
        raise NameError("name 'path_split' is not defined")   # what the reference does on an empty folder
    model = load_model(ms.MODEL_NAME, var.get('precision'), allow_random_weights(var))

    def process_one(p: str) -> int:
        print(f"[rank {rank}] Processing image: ", p)
        I = meta_segment(model, p)
        d, name = os.path.split(p)
        out = os.path.join(d, 'labels', name[:-4])
        ms.save_overlay(out + '.png', I)
        np.save(out, I)
        return model.last_count

    stats = {"rank": rank, "images": 0, "wall_s": 0.0, "setup_s": 0.0}

    def process_batch(mine: list) -> list:
        """This rank's share: *.tif through the overlapped pipeline on this GPU, anything else one by one."""
        import time
        tifs = [p for p in mine if p.lower().endswith('.tif')]
        counts = {}
        t0 = time.perf_counter()
        if tifs and not os.environ.get("ECSEG_SERIAL"):
            from . import tiffio
            from .pipeline import FilesPipeline
            shapes = [tiffio.probe(p) or ms._shape_of(p) for p in tifs]
            pipe = FilesPipeline(model.weights, model.precision, max(max(s[0] for s in shapes), 256),
                                 max(max(s[1] for s in shapes), 256), device=local, n_ctx=int(var.get('contexts', 2)),
                                 n_readers=int(var.get('readers', 4)), n_writers=int(var.get('writers', 6)),
                                 max_bytes_per_px=max(s[2] * s[3] for s in shapes))
            stats["setup_s"] = time.perf_counter() - t0      # contexts, weight upload, pinned slots
            try:
                t0 = time.perf_counter()
                counts.update(dict(pipe.run(tifs)))
            finally:
                pipe.close()
        out = [counts[p] if p in counts else process_one(p) for p in mine]
        stats["images"], stats["wall_s"] = len(mine), time.perf_counter() - t0
        return out

    if world > 1:
        dist.barrier()          # every rank has its model: the shares start together (the job's rate is images / slowest share)
    rows = run_sharded(paths, process_one, rank, world, lambda r: dist_gather(r, rank, world), process_batch)
    all_stats = dist_gather([stats], rank, world)
    if rows is not None:
        import json
        per_rank = [s[0] for s in all_stats]
        wall = max(s["wall_s"] for s in per_rank)
        print("[ecseg_b200.shard] " + json.dumps({"world": world, "images": len(rows), "slowest_share_s": wall,
                                                   "images_per_s": len(rows) / wall if wall > 0 else None,
                                                   "note": "wall_s = this rank's share through the pipeline (decode, GPU, file "
                                                           "writes); setup_s = contexts + weight upload before it",
                                                   "per_rank": per_rank, "host_cores": os.cpu_count()}))
        csv_path = os.path.join(inpath, 'ec_quantification.csv')
        print("Saving ec quantification to", csv_path)
        write_csv(csv_path, rows)
        if model.synthetic:
            print(f"[ecseg_b200] WARNING: {csv_path} and labels/* were produced with RANDOM-INIT weights (opt-in), "
                  "not with a trained metaseg checkpoint.", file=sys.stderr)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
