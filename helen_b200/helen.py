"""`helen` command line for the predict path: sub-commands `polish` and `call_consensus` with the
reference's flags and defaults (helen/helen.py:12-185, 302-326)."""
import argparse
import sys

from . import __version__
from .TextColor import TextColor


def _common(parser, threads_default, threads_help):
    parser.add_argument("-i", "--image_dir", type=str, required=True,
                        help="[REQUIRED] Path to a directory where all MarginPolish generated images are.")
    parser.add_argument("-m", "--model_path", type=str, required=True,
                        help="[REQUIRED] Path to a trained model (pkl file).")
    parser.add_argument("-b", "--batch_size", type=int, required=False, default=512,
                        help="Batch size for testing, default is 512.")
    parser.add_argument("-w", "--num_workers", type=int, required=False, default=8,
                        help="Number of workers to assign to the dataloader.")
    parser.add_argument("-t", "--threads", type=int, required=False, default=threads_default, help=threads_help)
    parser.add_argument("-o", "--output_dir", type=str, required=False, default='./output/',
                        help="Path to the output directory.")
    parser.add_argument("-p", "--output_prefix", type=str, required=False, default="HELEN_prediction",
                        help="Prefix for the output file. Default is: HELEN_prediction")
    parser.add_argument("-g", "--gpu_mode", default=False, action='store_true',
                        help="Run inference on GPUs (required: this package has no CPU path).")
    parser.add_argument("-d_ids", "--device_ids", type=str, required=False, default=None,
                        help="Comma separated GPU ids, e.g. 0,1,2. Default: all available devices.")
    parser.add_argument("-c", "--callers", type=int, required=False, default=8,
                        help="Accepted for compatibility; one caller per GPU is used.")
    return parser


def add_polish_arguments(parser):
    return _common(parser, 1, "Number of threads for the stitch step, default is 1.")


def add_call_consensus_arguments(parser):
    return _common(parser, 16, "Total available threads to use.")


def add_stitch_arguments(parser):
    """Arguments of sub-command "stitch" (helen/helen.py:188-222)."""
    parser.add_argument("-i", "--input_dir", type=str, required=True,
                        help="[REQUIRED] Path to a directory containing prediction files call consensus.")
    parser.add_argument("-o", "--output_dir", type=str, required=True, help="[REQUIRED] Path to the output directory.")
    parser.add_argument("-t", "--threads", type=int, required=True, help="[REQUIRED] Number of threads.")
    parser.add_argument("-p", "--output_prefix", type=str, required=False, default="HELEN_consensus",
                        help="Prefix for the output file. Default is: HELEN_consensus")
    return parser


def build_parser():
    parser = argparse.ArgumentParser(description="HELEN consensus calling on B200 (helen_b200).",
                                     formatter_class=argparse.RawTextHelpFormatter)
    subparsers = parser.add_subparsers(dest='sub_command')
    add_polish_arguments(subparsers.add_parser('polish', help="Run call_consensus then stitch."))
    add_call_consensus_arguments(subparsers.add_parser('call_consensus', help="Generate the prediction HDF5 files."))
    add_stitch_arguments(subparsers.add_parser('stitch', help="Stitch prediction files into a polished FASTA."))
    subparsers.add_parser('torch_stat', help="See PyTorch configuration.")
    subparsers.add_parser('version', help="Show program version.")
    return parser


def main(argv=None):
    parser = build_parser()
    flags, _ = parser.parse_known_args(argv)
    if flags.sub_command == 'polish':
        from .PolishInterface import polish_genome
        sys.stderr.write(TextColor.GREEN + "INFO: POLISH MODULE SELECTED\n" + TextColor.END)
        polish_genome(flags.image_dir, flags.model_path, flags.batch_size, flags.num_workers, flags.threads,
                      flags.output_dir, flags.output_prefix, flags.gpu_mode, flags.device_ids, flags.callers)
    elif flags.sub_command == 'call_consensus':
        from .CallConsensusInterface import call_consensus
        sys.stderr.write(TextColor.GREEN + "INFO: CALL CONSENSUS MODULE SELECTED\n" + TextColor.END)
        call_consensus(flags.image_dir, flags.model_path, flags.batch_size, flags.num_workers, flags.threads,
                       flags.output_dir, flags.output_prefix, flags.gpu_mode, flags.device_ids, flags.callers)
    elif flags.sub_command == 'stitch':
        from .StitchInterface import perform_stitch
        sys.stderr.write(TextColor.GREEN + "INFO: STITCH MODULE SELECTED\n" + TextColor.END)
        perform_stitch(flags.input_dir, flags.output_dir, flags.output_prefix, flags.threads)
    elif flags.sub_command == 'torch_stat':
        import torch
        sys.stderr.write(TextColor.YELLOW + "TORCH VERSION: " + TextColor.END + str(torch.__version__) + "\n")
        sys.stderr.write(TextColor.GREEN + "CUDA AVAILABLE: " + TextColor.END + str(torch.cuda.is_available()) + "\n")
        if torch.cuda.is_available():
            sys.stderr.write(TextColor.GREEN + "GPU DEVICES: " + TextColor.END + str(torch.cuda.device_count()) + "\n")
    elif flags.sub_command == 'version':
        print("HELEN (helen_b200) VERSION: ", __version__)
    else:
        sys.stderr.write(TextColor.RED + "ERROR: NO SUBCOMMAND SELECTED. PLEASE SELECT ONE OF THE AVAILABLE SUB-COMMANDS.\n"
                         + TextColor.END)
        parser.print_help()
        return 1
    return 0


if __name__ == '__main__':
    sys.exit(main())
