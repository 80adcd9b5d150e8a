"""Data-parallel training, one process per GPU (mirrors helen/modules/python/models/train_distributed.py:184-316).

Each rank runs the loop of models/train.py on its share of the images (DistributedSampler) and the ranks' gradients are
averaged after every chunk step with one NCCL all-reduce over a flat gradient buffer (models/grad_sync.py); rank 0
evaluates, logs and saves.  The reference gets the same arithmetic from DistributedDataParallel over a gloo group
(train_distributed.py:131,189).
"""
import datetime
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from .grad_sync import DataParallelContext
from .train import train


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def setup(rank, device_ids, args, port, backend="nccl"):
    """Body of one rank (train_distributed.py:184-204)."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    # rank 0 evaluates the test set after every epoch while the others wait in a broadcast: the wait must outlast it
    # (NCCL's default watchdog would abort every rank after 10 minutes)
    dist.init_process_group(backend, rank=rank, world_size=len(device_ids), timeout=datetime.timedelta(hours=6))
    try:
        torch.cuda.set_device(device_ids[rank])
        (train_file, test_file, batch_size, epochs, gpu_mode, num_workers, retrain_model, retrain_model_path,
         gru_layers, hidden_size, learning_rate, weight_decay, model_dir, stats_dir, train_mode) = args
        train(train_file, test_file, batch_size, epochs, gpu_mode, num_workers, retrain_model, retrain_model_path,
              gru_layers, hidden_size, learning_rate, weight_decay, model_dir, stats_dir, train_mode,
              dist_ctx=DataParallelContext(rank, len(device_ids)))
    finally:
        dist.destroy_process_group()


def train_distributed(train_file, test_file, batch_size, epochs, gpu_mode, num_workers, retrain_model,
                      retrain_model_path, gru_layers, hidden_size, learning_rate, weight_decay, model_dir,
                      stats_dir, device_ids, total_callers, train_mode):
    args = (train_file, test_file, batch_size, epochs, gpu_mode, num_workers, retrain_model, retrain_model_path,
            gru_layers, hidden_size, learning_rate, weight_decay, model_dir, stats_dir, train_mode)
    if total_callers != len(device_ids):
        raise ValueError("train_distributed: total_callers (%d) must equal the number of device ids (%d): one process per GPU"
                         % (total_callers, len(device_ids)))
    mp.spawn(setup, args=(device_ids, args, _free_port()), nprocs=len(device_ids), join=True)
