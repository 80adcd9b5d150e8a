"""Command line of the training side with the reference's sub-commands and flags (helen/helen_train.py:10-272):
`train`, `test`, `torch_stat`, `version`.  `hyperband` (hyper-parameter search, out of scope by SURVEY.md section 2)
is not mirrored."""
import argparse
import sys

from .TextColor import TextColor
from . import __version__


# flag tables: (names, keyword arguments); defaults and types are the reference's (helen_train.py:10-137)
_GPU_FLAG = (("--gpu_mode",), dict(default=False, action='store_true', help="Run on the GPU (required: helen_b200 has no CPU path)."))
TRAIN_FLAGS = [
    (("--train_image_dir",), dict(type=str, required=True, help="Directory of labelled MarginPolish images to train on.")),
    (("--test_image_dir",), dict(type=str, required=True, help="Directory of labelled images evaluated after every epoch.")),
    (("--batch_size",), dict(type=int, default=100, help="Images per batch, default 100.")),
    (("--epoch_size",), dict(type=int, default=10, help="Number of epochs, default 10.")),
    (("--output_dir",), dict(type=str, default='./model', help="Where trained_models_<stamp>/ is created.")),
    (("--retrain_model",), dict(type=bool, default=False, help="Continue from --retrain_model_path.")),
    (("--retrain_model_path",), dict(type=str, default=False, help="Checkpoint to continue from.")),
    _GPU_FLAG,
    (("-d_ids", "--device_ids"), dict(type=str, default=None, help="Comma-separated device ids; one training process per device.")),
    (("--num_workers",), dict(type=int, default=16, help="Data loader workers, default 16.")),
]
TEST_FLAGS = [
    (("--test_image_dir",), dict(type=str, required=True, help="Directory of labelled MarginPolish images.")),
    (("--batch_size",), dict(type=int, default=100, help="Images per batch, default 100.")),
    (("--model_path",), dict(type=str, default='./model', help="Checkpoint to evaluate.")),
    _GPU_FLAG,
    (("--print_details",), dict(default=False, action='store_true', help="Accepted for compatibility; only warns.")),
    (("--output_dir",), dict(type=str, default='./debug_output', help="Where the confusion matrices are written.")),
    (("--num_workers",), dict(type=int, default=40, help="Data loader workers, default 40.")),
]


def _add_flags(parser, table):
    for names, options in table:
        parser.add_argument(*names, **options)
    return parser


def add_train_arguments(parser):
    return _add_flags(parser, TRAIN_FLAGS)


def add_test_arguments(parser):
    return _add_flags(parser, TEST_FLAGS)


def build_parser():
    parser = argparse.ArgumentParser(description="HELEN training on B200 (helen_b200).",
                                     formatter_class=argparse.RawTextHelpFormatter)
    parser.add_argument("--version", default=False, action='store_true', help="Show version.")
    subparsers = parser.add_subparsers(dest='sub_command')
    add_train_arguments(subparsers.add_parser('train', help="Train a HELEN model. Requires a set of labeled images."))
    add_test_arguments(subparsers.add_parser('test', help="Test a model. Requires a set of labeled images"))
    subparsers.add_parser('torch_stat', help="See PyTorch configuration.")
    subparsers.add_parser('version', help="Show program version.")
    return parser


def main(argv=None):
    parser = build_parser()
    flags, _ = parser.parse_known_args(argv)
    if flags.sub_command == 'train':
        from .TrainInterface import train_interface
        sys.stderr.write(TextColor.GREEN + "INFO: TRAIN MODULE SELECTED\n" + TextColor.END)
        train_interface(flags.train_image_dir, flags.test_image_dir, flags.gpu_mode, flags.device_ids, flags.epoch_size,
                        flags.batch_size, flags.num_workers, flags.output_dir, flags.retrain_model, flags.retrain_model_path)
    elif flags.sub_command == 'test':
        from .TrainInterface import test_interface
        sys.stderr.write(TextColor.GREEN + "INFO: TEST MODULE SELECTED\n" + TextColor.END)
        test_interface(flags.test_image_dir, flags.batch_size, flags.gpu_mode, flags.num_workers, flags.model_path,
                       flags.output_dir, flags.print_details)
    elif flags.sub_command == 'torch_stat':
        import torch
        sys.stderr.write(TextColor.YELLOW + "TORCH VERSION: " + TextColor.END + str(torch.__version__) + "\n")
        sys.stderr.write(TextColor.GREEN + "CUDA AVAILABLE: " + TextColor.END + str(torch.cuda.is_available()) + "\n")
        sys.stderr.write(TextColor.GREEN + "GPU DEVICES: " + TextColor.END + str(torch.cuda.device_count()) + "\n")
    elif flags.sub_command == 'version' or flags.version is True:
        print("HELEN (helen_b200) VERSION: ", __version__)
    else:
        sys.stderr.write(TextColor.RED + "ERROR: NO SUBCOMMAND SELECTED. PLEASE SELECT ONE OF THE AVAILABLE SUB-COMMANDS.\n"
                         + TextColor.END)
        parser.print_help()
        return 1
    return 0


if __name__ == '__main__':
    sys.exit(main())
