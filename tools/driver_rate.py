#!/usr/bin/env python
"""End-to-end rate of the call_consensus driver on one GPU: MarginPolish-layout image files on disk ->
helen_b200.models.predict_gpu.predict() -> prediction file.

    python tools/driver_rate.py [--images 16384] [--features 10] [--batch 512] [--out profiles/r02_driver_rate.json]

Arms: the feed (native library + prefetch thread | native library behind DataLoader worker processes | general reader
behind worker processes) x the prediction schema (the reference's, three datasets per image | packed, one group per
batch).  Files are written with the package's own HDF5 layer into /dev/shm; the model is random (the rate does not
depend on the weights).  Prints one JSON object."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=65536)
    ap.add_argument("--files", type=int, default=4)
    ap.add_argument("--features", type=int, default=10)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    from helen_b200 import hdf5
    from helen_b200.models import predict_gpu
    from helen_b200.options import ImageSizeOptions

    tmp = tempfile.mkdtemp(prefix="helen_driver_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    rng = np.random.default_rng(0)
    per_file = args.images // args.files
    paths = []
    position = np.stack([np.arange(1000), np.zeros(1000, np.int64), np.zeros(1000, np.int64)], 1)
    for k in range(args.files):
        path = os.path.join(tmp, "images_%d.h5" % k)
        paths.append(path)
        with hdf5.open_file(path, "w") as f:
            block = rng.integers(0, 256, (per_file, 1000, args.features), dtype=np.uint8)
            for i in range(per_file):
                base = "images/img_%06d/" % i
                f[base + "contig"] = np.array([b"chr20"], dtype="S")
                f[base + "contig_start"] = np.array([(k * per_file + i) * 1000])
                f[base + "contig_end"] = np.array([(k * per_file + i) * 1000 + 1000])
                f[base + "feature_chunk_idx"] = np.array([i])
                f[base + "image"] = block[i]
                f[base + "position"] = position + (k * per_file + i) * 1000
    total = per_file * args.files
    # a random checkpoint in the reference's .pkl layout (ModelHander.py:61-76)
    from helen_b200.models.TransducerModel import TransducerGRU
    ImageSizeOptions.IMAGE_HEIGHT = args.features
    model = TransducerGRU(1, args.features, 1, 128, 5, 11)
    model_path = os.path.join(tmp, "model.pkl")
    torch.save({"model_state_dict": {"module." + k: v for k, v in model.state_dict().items()}, "model_optimizer": {},
                "hidden_size": 128, "gru_layers": 1, "epochs": 1}, model_path)
    result = {"images": total, "features": args.features, "batch": args.batch, "host_cores": os.cpu_count(), "hdf5_backend": hdf5.backend(),
              "gpu": torch.cuda.get_device_name(0), "file_bytes": sum(os.path.getsize(p) for p in paths), "arms": []}
    arms = [("native feed, prefetch thread", {}, 0), ("native feed, DataLoader processes", {"HELEN_B200_FEED_PROCESSES": "1"}, args.workers),
            ("general reader, DataLoader processes", {"HELEN_B200_NATIVE_FEED": "0"}, args.workers)]
    # warm-up: CUDA context, library load, page cache of the image files
    os.environ["HELEN_B200_PACKED_PREDICTIONS"] = "1"
    saved, sys.stderr = sys.stderr, open(os.devnull, "w")
    try:
        predict_gpu.predict(paths[:1], os.path.join(tmp, "warm"), model_path, args.batch, 0, 1, 0)
    finally:
        sys.stderr.close()
        sys.stderr = saved
    os.unlink(os.path.join(tmp, "warm_1.hdf"))
    for packed in ("1", "0"):
        for name, env, workers in arms:
            if packed == "0" and name != arms[0][0]:
                continue                                        # the reference schema is writer-bound: one arm is enough
            for key in ("HELEN_B200_FEED_PROCESSES", "HELEN_B200_NATIVE_FEED"):
                os.environ.pop(key, None)
            os.environ.update(env)
            os.environ["HELEN_B200_PACKED_PREDICTIONS"] = packed
            prefix = os.path.join(tmp, "pred_%s_%d" % (packed, len(result["arms"])))
            sys.stderr.flush()
            devnull = open(os.devnull, "w")
            saved = sys.stderr
            sys.stderr = devnull                                # the driver's per-batch progress lines
            t0 = time.perf_counter()
            try:
                predict_gpu.predict(paths, prefix, model_path, args.batch, workers, 1, 0)   # rank 1: no progress printing
            finally:
                sys.stderr = saved
                devnull.close()
            torch.cuda.synchronize()
            seconds = time.perf_counter() - t0
            out_bytes = os.path.getsize(prefix + "_1.hdf")
            result["arms"].append({"feed": name, "prediction_schema": "packed" if packed == "1" else "reference", "workers": workers,
                                   "seconds": seconds, "windows_per_s": total / seconds, "prediction_file_bytes": out_bytes})
            os.unlink(prefix + "_1.hdf")
    for p in os.listdir(tmp):
        os.unlink(os.path.join(tmp, p))
    os.rmdir(tmp)
    line = json.dumps(result)
    print(line)
    if args.out:
        with open(args.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
