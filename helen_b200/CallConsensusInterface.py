"""call_consensus with the reference's signature, checks and file sharding
(helen/modules/python/CallConsensusInterface.py:47-156)."""
import os
import sys
from os import listdir
from os.path import isfile, join

import torch

from . import _native
from .FileManager import FileManager
from .TextColor import TextColor


def get_file_paths_from_directory(directory_path):
    return [os.path.abspath(join(directory_path, file)) for file in listdir(directory_path)
            if isfile(join(directory_path, file)) and file[-2:] == 'h5']


def shard_files(input_files, callers):
    """Round-robin deal of input files over callers, empty callers dropped (:135-145)."""
    file_chunks = [[] for _ in range(callers)]
    for i, path in enumerate(input_files):
        file_chunks[i % callers].append(path)
    return [chunk for chunk in file_chunks if len(chunk) > 0]


def _fail(message):
    sys.stderr.write(TextColor.RED + message + "\n" + TextColor.END)
    exit(1)


def call_consensus(image_dir, model_path, batch_size, num_workers, threads, output_dir, output_prefix, gpu_mode,
                   device_ids, callers):
    if not os.path.isfile(model_path):
        _fail("ERROR: CAN NOT LOCATE MODEL FILE.")
    if not os.path.isdir(image_dir):
        _fail("ERROR: CAN NOT LOCATE IMAGE DIRECTORY.")
    if batch_size <= 0:
        _fail("ERROR: batch_size NEEDS TO BE >0.")
    if num_workers < 0:
        _fail("ERROR: num_workers NEEDS TO BE >=0.")
    if threads <= 0:
        _fail("ERROR: THREAD NEEDS TO BE >=0.")

    output_dir = FileManager.handle_output_directory(output_dir)
    output_filename = os.path.join(output_dir, output_prefix)
    sys.stderr.write(TextColor.GREEN + "INFO: " + TextColor.END + "OUTPUT FILE: " + output_filename + "\n")

    if not gpu_mode:
        _fail("ERROR: helen_b200 IS THE GPU (sm_100a) IMPLEMENTATION OF THIS PATH AND HAS NO CPU FALLBACK. "
              "RUN WITH --gpu_mode, OR USE THE REFERENCE PACKAGE FOR CPU INFERENCE.")
    if not torch.cuda.is_available() or _native.load().hb_device_count() <= 0:
        _fail("ERROR: NO CUDA DEVICE AVAILABLE.")

    if device_ids is None:
        total_gpu_devices = torch.cuda.device_count()
        sys.stderr.write(TextColor.GREEN + "INFO: TOTAL GPU AVAILABLE: " + str(total_gpu_devices) + "\n" + TextColor.END)
        device_ids = list(range(total_gpu_devices))
    else:
        device_ids = [int(i) for i in device_ids.split(',')]
        for device_id in device_ids:
            major, minor = torch.cuda.get_device_capability(device=device_id)
            if major != 10:
                _fail("ERROR: GPU DEVICE: " + str(device_id) + " IS sm_" + str(major) + str(minor)
                      + "; helen_b200 REQUIRES sm_100 (B200).")
            sys.stderr.write(TextColor.GREEN + "INFO: CAPABILITY OF GPU#" + str(device_id) + ":\t" + str(major)
                             + "-" + str(minor) + "\n" + TextColor.END)
    callers = len(device_ids)
    sys.stderr.write(TextColor.GREEN + "INFO: AVAILABLE GPU DEVICES: " + str(device_ids) + "\n" + TextColor.END)

    file_chunks = shard_files(get_file_paths_from_directory(image_dir), callers)
    callers = len(file_chunks)
    if callers == 0:
        _fail("ERROR: NO .h5 IMAGE FILES FOUND IN " + str(image_dir))

    from .models.predict_gpu import predict_gpu
    predict_gpu(file_chunks, output_filename, model_path, batch_size, callers, device_ids, num_workers)
    sys.stderr.write(TextColor.GREEN + "INFO: " + TextColor.END + "PREDICTION GENERATED SUCCESSFULLY.\n")
