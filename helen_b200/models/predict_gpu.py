"""GPU inference driver with the reference's entry points
(helen/modules/python/models/predict_gpu.py:38, :186, :207).

One OS process per GPU (mp.spawn), each with its own file list and its own
``<prefix>_<rank>.hdf`` output; no process group and no collective: in the reference the gloo
group only served a DistributedDataParallel wrapper that does nothing under no_grad.  The whole
per-batch loop body of the reference (predict_gpu.py:97-159) is one call into the CUDA library.
"""
import os
import queue
import threading
import sys
import time

import torch
import torch.multiprocessing as mp
from torch.utils.data import DataLoader

from ..DataStore import DataStore
from ..options import ImageSizeOptions
from ..predictor import WindowPredictor
from ..TextColor import TextColor
from .bulk_reader import BulkImageBatches
from .prefetch import ThreadPrefetcher
from .dataloader_predict import SequenceDataset
from .ModelHander import ModelHandler


def _cuda_device(device_id):
    return torch.device("cuda", device_id)


def predict(test_file, output_filename, model_path, batch_size, num_workers, rank, device_id):
    prediction_data_file = DataStore(output_filename + "_" + str(rank) + ".hdf", mode='w')
    checkpoint = ModelHandler.load_checkpoint(model_path)
    state_dict = checkpoint['model_state_dict']
    if checkpoint['gru_layers'] != 1 or checkpoint['hidden_size'] != 128:
        raise ValueError("helen_b200 supports gru_layers=1, hidden_size=128 checkpoints only "
                         f"(got {checkpoint['gru_layers']}, {checkpoint['hidden_size']})")
    torch.cuda.set_device(device_id)
    predictor = WindowPredictor(state_dict, device=device_id)
    if predictor.image_features != ImageSizeOptions.IMAGE_HEIGHT:
        sys.stderr.write(TextColor.YELLOW + "WARN: MODEL EXPECTS " + str(predictor.image_features)
                         + " FEATURES PER COLUMN, MARGINPOLISH IMAGES HAVE " + str(ImageSizeOptions.IMAGE_HEIGHT)
                         + ".\n" + TextColor.END)
    if rank == 0:
        print(output_filename + "_" + str(rank) + ".hdf")
        sys.stderr.write(TextColor.PURPLE + 'Loading data\n' + TextColor.END)

    # pinned batches: the host->device copy is then a plain DMA (no staging memcpy) and can run asynchronously
    if os.environ.get("HELEN_B200_ITEM_READER", "0") not in ("", "0"):
        # the reference's feed: one image per item, collated by the DataLoader (dataloader_predict.py:54-88)
        test_data = SequenceDataset(image_directory=None, file_list=test_file)
        test_loader = DataLoader(test_data, batch_size=batch_size, shuffle=False, num_workers=num_workers, pin_memory=True)
    else:
        # bulk feed: an item is a whole batch read from one file in one pass (models/bulk_reader.py).  Files the native feed
        # library reads are fetched by its threads straight into a ring of page-locked buffers, one batch ahead of the GPU
        # (models/prefetch.py); other files go through DataLoader worker processes as in the reference.
        # Ring of 8: two batches queued, one being filled, the one on the GPU, up to four with the writer thread.
        test_data = BulkImageBatches(image_directory=None, file_list=test_file, batch_size=batch_size, ring=8)
        if test_data.native_for_all() and os.environ.get("HELEN_B200_FEED_PROCESSES", "0") in ("", "0"):
            test_loader = ThreadPrefetcher(test_data, depth=2, held=5)
        else:
            test_data.ring = 0
            test_loader = DataLoader(test_data, batch_size=None, shuffle=False, num_workers=num_workers, pin_memory=True)
    total_batches = len(test_loader)
    windows_done, t_begin = 0, time.time()
    device = _cuda_device(device_id)

    def write_out(item):
        # predict_gpu.py:176-179: one record per image; runs while the GPU works on the next batch
        meta, base_dev, rle_dev, done = item
        done.synchronize()
        base_labels, rle_labels = base_dev.cpu().numpy(), rle_dev.cpu().numpy()
        contig, contig_start, contig_end, chunk_id, position, filename = meta
        prediction_data_file.write_predictions(contig, contig_start, contig_end, chunk_id, position,
                                               base_labels, rle_labels, filename)

    # The prediction file is written by its own thread, at most three batches behind the GPU: the loop below only reads a
    # batch off the feed and launches it (the writer's numpy conversions and file writes release the interpreter lock).
    pending = queue.Queue(maxsize=3)
    writer_error = []

    def writer():
        torch.cuda.set_device(device_id)                # (the current device is a per-thread setting)
        while True:
            item = pending.get()
            if item is None:
                return
            if writer_error:
                continue                                # keep draining so that the producer never blocks
            try:
                write_out(item)
            except BaseException as error:
                writer_error.append(error)

    writer_thread = threading.Thread(target=writer, name="helen-prediction-writer", daemon=True)
    writer_thread.start()
    for batch_iterator, (contig, contig_start, contig_end, chunk_id, images, position, filename) in enumerate(test_loader, 1):
        start_time = time.time()
        # the whole loop body of the reference (predict_gpu.py:97-159) is this one asynchronous call
        base_dev, rle_dev = predictor.predict(images.to(device, non_blocking=True))
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(device))
        pending.put(((contig, contig_start, contig_end, chunk_id, position, filename), base_dev, rle_dev, done))
        if writer_error:
            break
        windows_done += images.size(0)
        if rank == 0:
            eta = (time.time() - start_time) * (total_batches - batch_iterator)
            stamp = "{} HOURS {} MINS {} SECS.".format(int(eta / 3600), int(eta % 3600 / 60), int(eta) % 60)
            sys.stderr.write(TextColor.GREEN + "INFO: BATCHES DONE: " + str(batch_iterator) + "/" + str(total_batches)
                             + ". ESTIMATED TIME LEFT: " + stamp + " ("
                             + str(int(windows_done / max(time.time() - t_begin, 1e-9))) + " WINDOWS/S)\n" + TextColor.END)
    pending.put(None)
    writer_thread.join()
    if writer_error:
        raise writer_error[0]
    prediction_data_file.close()
    predictor.close()


def setup(rank, total_callers, args, all_input_files, all_devices):
    output_filepath, model_path, batch_size, num_workers = args
    predict(all_input_files[rank], output_filepath, model_path, batch_size, num_workers, rank, all_devices[rank])


def predict_gpu(file_chunks, output_filepath, model_path, batch_size, total_callers, devices, num_workers):
    args = (output_filepath, model_path, batch_size, num_workers)
    if total_callers == 1:
        setup(0, 1, args, file_chunks, devices)
        return
    mp.spawn(setup, args=(total_callers, args, file_chunks, devices), nprocs=total_callers, join=True)
