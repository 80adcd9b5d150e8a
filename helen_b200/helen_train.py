"""Command line of the training side with the reference's sub-commands and flags (helen/helen_train.py:10-272):
`train`, `test`, `torch_stat`, `version`.  `hyperband` (hyper-parameter search, out of scope by SURVEY.md section 2)
is not mirrored."""
import argparse
import sys

from .TextColor import TextColor
from . import __version__


def add_train_arguments(parser):
    """helen_train.py:10-84"""
    parser.add_argument("--train_image_dir", type=str, required=True, help="Training data directory containing HDF files.")
    parser.add_argument("--test_image_dir", type=str, required=True, help="Training data directory containing HDF files.")
    parser.add_argument("--batch_size", type=int, required=False, default=100, help="Batch size for training, default is 100.")
    parser.add_argument("--epoch_size", type=int, required=False, default=10, help="Epoch size for training iteration.")
    parser.add_argument("--output_dir", type=str, required=False, default='./model', help="Path to the output directory.")
    parser.add_argument("--retrain_model", type=bool, default=False, help="If true then retrain a pre-trained mode.")
    parser.add_argument("--retrain_model_path", type=str, default=False, help="Path to the model that will be retrained.")
    parser.add_argument("--gpu_mode", default=False, action='store_true', help="If set then PyTorch will use GPUs. CUDA required.")
    parser.add_argument("-d_ids", "--device_ids", type=str, required=False, default=None,
                        help="List of gpu device ids to use. helen_b200 trains on the first one.")
    parser.add_argument("--num_workers", type=int, required=False, default=16, help="Number of data loader workers.")
    return parser


def add_test_arguments(parser):
    """helen_train.py:87-137"""
    parser.add_argument("--test_image_dir", type=str, required=True, help="Training data directory containing HDF files.")
    parser.add_argument("--batch_size", type=int, required=False, default=100, help="Batch size for training, default is 100.")
    parser.add_argument("--model_path", type=str, required=False, default='./model', help="Path of the model to load and test.")
    parser.add_argument("--gpu_mode", default=False, action='store_true', help="If set then PyTorch will use GPUs. CUDA required.")
    parser.add_argument("--print_details", default=False, action='store_true', help="Not mirrored: prints a warning.")
    parser.add_argument("--output_dir", type=str, required=False, default='./debug_output', help="Output directory.")
    parser.add_argument("--num_workers", type=int, required=False, default=40, help="Number of data loader workers.")
    return parser


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
