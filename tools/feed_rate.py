#!/usr/bin/env python
"""Host-side feed and writer rates, no GPU needed:
    python tools/feed_rate.py [--images 4096] [--features 90] [--workers 0,4,8] [--out profiles/r02_feed_rate.json]
Writes MarginPolish-layout image files with the package's HDF5 layer, then times
  * the reference-style feed: SequenceDataset (one image per item) + DataLoader collation,
  * the bulk feed: BulkImageBatches (one batch per item) through the general reader and through the native feed library,
  * DataStore.write_predictions in the reference schema and in the packed schema,
and prints one JSON object (windows/s per arm, the backend used, host cores)."""
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
    ap.add_argument("--images", type=int, default=4096)
    ap.add_argument("--files", type=int, default=4)
    ap.add_argument("--features", type=int, default=90)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--workers", default="0,4,8")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    from torch.utils.data import DataLoader
    from helen_b200 import hdf5
    from helen_b200.DataStore import DataStore
    from helen_b200.models.bulk_reader import BulkImageBatches
    from helen_b200.models.dataloader_predict import SequenceDataset

    tmp = tempfile.mkdtemp(prefix="helen_feed_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    rng = np.random.default_rng(0)
    per_file = args.images // args.files
    paths = []
    t0 = time.perf_counter()
    for k in range(args.files):
        path = os.path.join(tmp, "images_%d.h5" % k)
        paths.append(path)
        with hdf5.open_file(path, "w") as f:
            for i in range(per_file):
                base = "images/img_%06d/" % i
                f[base + "contig"] = np.array([b"chr20"], dtype="S")
                f[base + "contig_start"] = np.array([i * 1000])
                f[base + "contig_end"] = np.array([i * 1000 + 1000])
                f[base + "feature_chunk_idx"] = np.array([i])
                f[base + "image"] = rng.integers(0, 256, (1000, args.features), dtype=np.uint8)
                f[base + "position"] = np.stack([np.arange(1000) + i * 1000, np.zeros(1000, np.int64), np.zeros(1000, np.int64)], 1)
    total = per_file * args.files
    result = {"backend": hdf5.backend(), "host_cores": os.cpu_count(), "images": total, "features": args.features,
              "batch": args.batch, "torch": torch.__version__, "file_bytes": sum(os.path.getsize(p) for p in paths),
              "write_image_files_windows_per_s": total / (time.perf_counter() - t0), "feed": []}
    for workers in [int(w) for w in args.workers.split(",")]:
        row = {"workers": workers}
        for name, make in (("item_reader", lambda: DataLoader(SequenceDataset(None, file_list=paths), batch_size=args.batch, shuffle=False, num_workers=workers)),
                           ("bulk_reader", lambda: DataLoader(BulkImageBatches(None, file_list=paths, batch_size=args.batch, native=False), batch_size=None, shuffle=False, num_workers=workers)),
                           ("native_reader_1_thread", lambda: DataLoader(BulkImageBatches(None, file_list=paths, batch_size=args.batch, native=True, threads=1), batch_size=None, shuffle=False, num_workers=workers)),
                           ("native_reader_4_threads", lambda: DataLoader(BulkImageBatches(None, file_list=paths, batch_size=args.batch, native=True, threads=4), batch_size=None, shuffle=False, num_workers=workers)),
                           ("native_reader_8_threads", lambda: DataLoader(BulkImageBatches(None, file_list=paths, batch_size=args.batch, native=True, threads=8), batch_size=None, shuffle=False, num_workers=workers))):
            loader = make()
            t0 = time.perf_counter()
            seen, checksum = 0, 0
            for batch in loader:
                seen += batch[4].shape[0]
                checksum += int(batch[4][:, ::97, ::13].sum())
            row[name + "_windows_per_s"] = seen / (time.perf_counter() - t0)
            row[name + "_checksum"] = checksum
            assert seen == total
        result["feed"].append(row)
    # prediction writer
    position = np.stack([np.arange(1000), np.zeros(1000, np.int64), np.zeros(1000, np.int64)], 1)[None].repeat(args.batch, 0)
    bases = rng.integers(0, 5, (args.batch, 1000)).astype(np.uint8)
    rles = rng.integers(0, 11, (args.batch, 1000)).astype(np.uint8)
    for packed in (False, True):
        store = DataStore(os.path.join(tmp, "pred_%d.hdf" % packed), mode="w", packed=packed)
        t0 = time.perf_counter()
        n = 0
        for b in range(max(1, total // args.batch)):
            ids = np.arange(b * args.batch, (b + 1) * args.batch)
            store.write_predictions(["chr20"] * args.batch, ids * 1000, ids * 1000 + 1000, np.zeros(args.batch, np.int64), position, bases, rles)
            n += args.batch
        store.close()
        result["write_predictions_%s_windows_per_s" % ("packed" if packed else "reference_schema")] = n / (time.perf_counter() - t0)
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
