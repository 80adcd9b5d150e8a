"""train_interface / test_interface with the reference's signatures
(helen/modules/python/TrainInterface.py:127-147, TestInterface.py:90-138).

The per-chunk step is the CUDA library's hb_train_step_chunk (models/train_step.py).  One device id trains in this
process (models/train.py); several run one process per GPU with the gradients averaged after every step
(models/train_distributed.py, the reference's DistributedDataParallel mode).  test_interface writes the two confusion
matrices as text instead of matplotlib images (matplotlib is not a dependency here)."""
import os
import sys

import torch

from .FileManager import FileManager
from .options import ImageSizeOptions, TrainOptions
from .TextColor import TextColor


class TrainModule:
    """Holds one training configuration (TrainInterface.py:18-125)."""

    def __init__(self, train_file, test_file, gpu_mode, device_ids, max_epochs, batch_size, num_workers,
                 retrain_model, retrain_model_path, model_dir, stats_dir):
        self.train_file = train_file
        self.test_file = test_file
        self.gpu_mode = gpu_mode
        self.device_ids = device_ids
        self.model_dir = model_dir
        self.epochs = max_epochs
        self.batch_size = batch_size
        self.num_workers = num_workers
        self.retrain_model = retrain_model
        self.retrain_model_path = retrain_model_path
        self.stats_dir = stats_dir
        self.hidden_size = TrainOptions.HIDDEN_SIZE
        self.gru_layers = TrainOptions.GRU_LAYERS
        self.learning_rate = 0.0001          # TrainInterface.py:37-38
        self.weight_decay = 0

    def selected_devices(self):
        """--device_ids (all visible devices when absent), checked like TrainInterface.py:66-92."""
        if not torch.cuda.is_available():
            sys.stderr.write(TextColor.RED + "ERROR: TORCH IS NOT BUILT WITH CUDA.\n" + TextColor.END)
            exit(1)
        if self.device_ids is None:
            device_ids = list(range(torch.cuda.device_count()))
            sys.stderr.write(TextColor.GREEN + "INFO: TOTAL GPU AVAILABLE: " + str(len(device_ids)) + "\n" + TextColor.END)
        else:
            device_ids = [int(i) for i in self.device_ids.split(',')]
        if len(device_ids) == 0:
            sys.stderr.write(TextColor.RED + "ERROR: NO GPU AVAILABLE BUT GPU MODE IS SET\n" + TextColor.END)
            exit()
        return device_ids

    def train_model_gpu(self):
        device_ids = self.selected_devices()
        if len(device_ids) == 1:
            from .models.train import train
            torch.cuda.set_device(device_ids[0])
            train(self.train_file, self.test_file, self.batch_size, self.epochs, self.gpu_mode, self.num_workers,
                  self.retrain_model, self.retrain_model_path, self.gru_layers, self.hidden_size, self.learning_rate,
                  self.weight_decay, self.model_dir, self.stats_dir, not_hyperband=True)
        else:
            from .models.train_distributed import train_distributed
            train_distributed(self.train_file, self.test_file, self.batch_size, self.epochs, self.gpu_mode,
                              self.num_workers, self.retrain_model, self.retrain_model_path, self.gru_layers,
                              self.hidden_size, self.learning_rate, self.weight_decay, self.model_dir, self.stats_dir,
                              device_ids, len(device_ids), train_mode=True)

    def train_model(self):
        sys.stderr.write(TextColor.RED + "ERROR: helen_b200 HAS NO CPU PATH, USE --gpu_mode.\n" + TextColor.END)
        exit(1)


def train_interface(train_dir, test_dir, gpu_mode, device_ids, epoch_size, batch_size, num_workers, output_dir,
                    retrain_model, retrain_model_path):
    model_out_dir, stats_dir = FileManager.handle_train_output_directory(output_dir)
    tm = TrainModule(train_dir, test_dir, gpu_mode, device_ids, epoch_size, batch_size, num_workers,
                     retrain_model, retrain_model_path, model_out_dir, stats_dir)
    if gpu_mode:
        tm.train_model_gpu()
    else:
        tm.train_model()
    return model_out_dir, stats_dir


def write_confusion_matrix(matrix, labels, path):
    """Rows = true label, columns = predicted label (the orientation of TestInterface.py:26-78's plots)."""
    with open(path, 'w') as out:
        out.write("true\\pred\t" + "\t".join(labels) + "\n")
        for label, row in zip(labels, matrix):
            out.write(label + "\t" + "\t".join(str(int(v)) for v in row) + "\n")


def test_interface(test_file, batch_size, gpu_mode, num_workers, model_path, output_directory, print_details):
    from .models.ModelHander import ModelHandler
    from .models.test import test
    sys.stderr.write(TextColor.PURPLE + 'Loading data\n' + TextColor.END)
    output_directory = FileManager.handle_output_directory(output_directory)
    if os.path.isfile(model_path) is False:
        sys.stderr.write(TextColor.RED + "ERROR: INVALID PATH TO MODEL\n")
        exit(1)
    if not gpu_mode:
        sys.stderr.write(TextColor.RED + "ERROR: helen_b200 HAS NO CPU PATH, USE --gpu_mode.\n" + TextColor.END)
        exit(1)
    if print_details:
        sys.stderr.write(TextColor.YELLOW + "WARN: --print_details (per-mismatch dumps, test_debug.py) IS NOT MIRRORED.\n"
                         + TextColor.END)
    sys.stderr.write(TextColor.GREEN + "INFO: MODEL LOADING\n" + TextColor.END)
    checkpoint = ModelHandler.load_checkpoint(model_path)
    image_features = next(int(v.shape[1]) for k, v in checkpoint['model_state_dict'].items()
                          if k.endswith('gru_encoder.weight_ih_l0'))      # with or without the 'module.' prefix
    transducer_model, hidden_size, gru_layers, prev_ite = ModelHandler.load_simple_model(
        model_path, input_channels=ImageSizeOptions.IMAGE_CHANNELS, image_features=image_features,
        seq_len=ImageSizeOptions.SEQ_LENGTH, num_base_classes=ImageSizeOptions.TOTAL_BASE_LABELS,
        num_rle_classes=ImageSizeOptions.TOTAL_RLE_LABELS)
    sys.stderr.write(TextColor.GREEN + "INFO: MODEL LOADED\n" + TextColor.END)
    transducer_model = transducer_model.cuda()
    stats_dictionary = test(test_file, batch_size, gpu_mode, transducer_model, num_workers, gru_layers, hidden_size,
                            num_base_classes=ImageSizeOptions.TOTAL_BASE_LABELS,
                            num_rle_classes=ImageSizeOptions.TOTAL_RLE_LABELS)
    write_confusion_matrix(stats_dictionary['rle_confusion_matrix'].tolist(),
                           [str(i) for i in range(ImageSizeOptions.TOTAL_RLE_LABELS)],
                           os.path.join(output_directory, "RLE_CONFUSION_MATRIX.txt"))
    # label order of Options.py:3 (the reference's plot swaps the G and T tick labels, TestInterface.py:59)
    write_confusion_matrix(stats_dictionary['base_confusion_matrix'].tolist(), ['-', 'A', 'C', 'G', 'T'],
                           os.path.join(output_directory, "BASE_CONFUSION_MATRIX.txt"))
    return stats_dictionary
