"""Training / evaluation dataset (mirrors helen/modules/python/models/dataloader.py:9-70): every image of every
``*.h5`` file in a directory, returned with its labels as (image u8 [T, F], label_base [T], label_run_length [T])."""
import sys

import numpy as np
from torch.utils.data import Dataset

from .. import hdf5
from ..FileManager import FileManager
from ..TextColor import TextColor


class SequenceDataset(Dataset):
    def __init__(self, image_directory):
        file_image_pair = []
        for hdf5_file_path in FileManager.get_file_paths_from_directory(image_directory):
            with hdf5.open_file(hdf5_file_path, 'r') as hdf5_file:
                if 'images' in hdf5_file:                                   # dataloader.py:35-41
                    for image_name in list(hdf5_file['images'].keys()):
                        file_image_pair.append((hdf5_file_path, image_name))
                else:
                    sys.stderr.write(TextColor.YELLOW + "WARN: NO IMAGES FOUND IN FILE: " + hdf5_file_path + "\n" + TextColor.END)
        self.all_images = file_image_pair

    def __getitem__(self, index):
        hdf5_filepath, image_name = self.all_images[index]
        with hdf5.open_file(hdf5_filepath, 'r') as hdf5_file:                # dataloader.py:58-61
            group = hdf5_file['images'][image_name]
            image = np.array(group['image'][()])                             # (copies: the file is closed on return)
            label_base = np.array(group['label_base'][()]).reshape(-1)
            label_run_length = np.array(group['label_run_length'][()]).reshape(-1)
        return image, label_base, label_run_length

    def __len__(self):
        return len(self.all_images)
