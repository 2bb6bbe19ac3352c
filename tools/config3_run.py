#!/usr/bin/env python3
"""BASELINE.json config 3 through the PRODUCT driver: K synthetic 2048x2048 DAPI TIFFs in a folder ->
`python -m torch.distributed.run --nproc-per-node N -m ecseg_b200.shard` (the sharded `make metaseg`,
reference loop src/metaseg.py:42-57) for every N in --gpus, one process per GPU.  Checks that
ec_quantification.csv is the same set of rows whatever N is and that every artefact exists; prints one JSON line.

    python tools/config3_run.py --images 256 --gpus 1,2 [--dir /dev/shm]
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=256)
    ap.add_argument("--distinct", type=int, default=32)
    ap.add_argument("--gpus", default="1,2")
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--port", type=int, default=29611)
    args = ap.parse_args()
    from ecseg_b200 import synth
    base = args.dir if os.path.isdir(args.dir) else tempfile.gettempdir()
    per_image = 4 * 2048 * 2048 // 4 + 37948992 + (1 << 20)      # input tif + the three artefacts
    free = shutil.disk_usage(base).free
    if free < args.images * per_image * 1.3 + (8 << 30):
        fit = max(8, int((free - (8 << 30)) / (per_image * 1.3)))
        print(f"[config3] {base} has {free >> 30} GiB free: {args.images} images do not fit, using {fit}", file=sys.stderr)
        args.images = fit
    work = tempfile.mkdtemp(prefix="ecseg_cfg3_", dir=args.dir if os.path.isdir(args.dir) else None)
    data = os.path.join(work, "data")
    os.mkdir(data)
    try:
        t0 = time.time()
        distinct = [synth.synth_dapi(7000 + s, 2048, 2048) for s in range(args.distinct)]
        for i in range(args.images):
            cv2.imwrite(os.path.join(data, f"img{i:05d}.tif"), distinct[i % args.distinct], [cv2.IMWRITE_TIFF_COMPRESSION, 1])
        gen_s = time.time() - t0
        with open(os.path.join(work, "config.yaml"), "w") as f:
            f.write(f"metaseg:\n  inpath: {data}\n")
        env = dict(os.environ, PYTHONPATH=ROOT, ECSEG_ALLOW_RANDOM_WEIGHTS="1")
        runs, csvs = [], {}
        for n in [int(x) for x in args.gpus.split(",")]:
            for sub in ("dapi", "labels"):
                shutil.rmtree(os.path.join(data, sub), ignore_errors=True)
            if os.path.exists(os.path.join(data, "ec_quantification.csv")):
                os.remove(os.path.join(data, "ec_quantification.csv"))
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                   "--master-port", str(args.port + n), "-m", "ecseg_b200.shard"]
            t0 = time.time()
            r = subprocess.run(cmd, cwd=work, env=env, capture_output=True, text=True, timeout=3000)
            wall = time.time() - t0
            if r.returncode != 0:
                print(r.stdout[-3000:], r.stderr[-3000:], file=sys.stderr)
                raise SystemExit(f"shard driver failed at N={n}")
            m = re.search(r"\[ecseg_b200\.shard\] (\{.*\})", r.stdout)
            st = json.loads(m.group(1)) if m else {}
            rows = open(os.path.join(data, "ec_quantification.csv")).read().strip().splitlines()
            assert rows[0] == "image name,# of ec" and len(rows) == args.images + 1, (rows[0], len(rows))
            csvs[n] = sorted(rows[1:])
            n_npy = len([f for f in os.listdir(os.path.join(data, "labels")) if f.endswith(".npy")])
            n_png = len([f for f in os.listdir(os.path.join(data, "labels")) if f.endswith(".png")])
            n_tif = len(os.listdir(os.path.join(data, "dapi")))
            assert n_npy == n_png == n_tif == args.images, (n_npy, n_png, n_tif)
            runs.append({"gpus": n, "images_per_s": st.get("images_per_s"), "slowest_share_s": st.get("slowest_share_s"),
                         "process_wall_s": wall, "per_rank": st.get("per_rank"), "host_cores": st.get("host_cores")})
        first = csvs[next(iter(csvs))]
        same = all(v == first for v in csvs.values())
        out = {"what": "BASELINE config 3 through the product driver: python -m torch.distributed.run -m ecseg_b200.shard over a folder of "
                       f"{args.images} uncompressed 2048x2048 u8 TIFFs ({args.distinct} distinct) in {args.dir}; per image dapi tif + RGBA png + "
                       "int64 npy are written, rank 0 writes ec_quantification.csv; images_per_s = images / slowest rank's share "
                       "(model load and process start-up excluded, reported as process_wall_s)",
               "images": args.images, "tiff_generation_s": gen_s, "runs": runs, "csv_rows_equal_across_gpu_counts": same,
               "sum_ec": sum(int(r.rsplit(",", 1)[1]) for r in first),
               "bytes_written_per_image": 8 * 2048 * 2048 + 128 + 2048 * 2048 + 128 + 200000}
        print(json.dumps(out))
        assert same, "CSV rows differ between GPU counts"
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
