"""SequenceDataset with the reference's item contract
(helen/modules/python/models/dataloader_predict.py:11-95).

__getitem__ -> (contig, contig_start, contig_end, chunk_id, image u8[1000, F], position int[1000, 3], path)
Images shorter than SEQ_LENGTH are right-padded with zero columns, positions with (-1, -1, -1).

Items are listed file by file, so consecutive reads hit the same file: the handle of the file read last stays
open (one per dataset copy, i.e. per DataLoader worker) instead of one open/close per image as the reference
does (dataloader_predict.py:61).
"""
import os
import sys

import numpy as np
from torch.utils.data import Dataset

from .. import hdf5
from ..FileManager import FileManager
from ..options import ImageSizeOptions
from ..TextColor import TextColor


def _scalar(dataset_value):
    """MarginPolish stores contig/start/end/chunk index as 1-element datasets."""
    return np.asarray(dataset_value).reshape(-1)[0]


class SequenceDataset(Dataset):
    def __init__(self, image_directory, file_list=None):
        if file_list is not None:
            hdf_files = file_list
        else:
            hdf_files = FileManager.get_file_paths_from_directory(image_directory)
        file_image_pair = []
        for hdf5_file_path in hdf_files:
            with hdf5.open_file(hdf5_file_path, 'r') as hdf5_file:
                if 'images' in hdf5_file:
                    for image_name in list(hdf5_file['images'].keys()):
                        file_image_pair.append((hdf5_file_path, image_name))
                else:
                    sys.stderr.write(TextColor.YELLOW + "WARN: NO IMAGES FOUND IN FILE: "
                                     + hdf5_file_path + "\n" + TextColor.END)
        self.all_images = file_image_pair
        self._open_path, self._open_file, self._open_pid = None, None, None

    def _file(self, path):
        if self._open_pid != os.getpid():           # a handle inherited through fork is not ours to use or close
            self._open_path, self._open_file, self._open_pid = None, None, os.getpid()
        if self._open_path != path:
            self.close()
            self._open_file, self._open_path = hdf5.open_file(path, 'r'), path
        return self._open_file

    def close(self):
        if self._open_file is not None and self._open_pid == os.getpid():
            self._open_file.close()
        self._open_path, self._open_file = None, None

    def __getstate__(self):
        state = dict(self.__dict__)                 # handles do not travel to DataLoader workers
        state["_open_path"], state["_open_file"] = None, None
        return state

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __getitem__(self, index):
        hdf5_filepath, image_name = self.all_images[index]
        hdf5_file = self._file(hdf5_filepath)
        group = hdf5_file['images'][image_name]
        contig = _scalar(group['contig'][()])
        if isinstance(contig, bytes):
            contig = contig.decode()
        contig = str(contig).replace("'", '')
        contig_start = int(_scalar(group['contig_start'][()]))
        contig_end = int(_scalar(group['contig_end'][()]))
        chunk_id = int(_scalar(group['feature_chunk_idx'][()]))
        image = np.asarray(group['image'][()]).astype(np.uint8)
        position = np.asarray(group['position'][()]).astype(np.int64)

        seq = ImageSizeOptions.SEQ_LENGTH
        if image.shape[0] < seq:
            missing = seq - image.shape[0]
            image = np.concatenate([image, np.zeros((missing, image.shape[1]), dtype=np.uint8)], axis=0)
            position = np.concatenate([position, np.full((missing, 3), -1, dtype=np.int64)], axis=0)
        if image.shape[0] < seq or position.shape[0] < seq:
            raise ValueError("IMAGE SIZE ERROR: " + str(hdf5_filepath) + " " + str(image.shape))
        return contig, contig_start, contig_end, chunk_id, image, position, hdf5_filepath

    def __len__(self):
        return len(self.all_images)
