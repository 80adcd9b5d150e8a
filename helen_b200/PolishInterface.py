"""polish_genome with the reference's signature (helen/modules/python/PolishInterface.py:49-105):
call_consensus into a timestamped prediction directory, then stitch.

Stitch (label decoding, anchored overlap resolution -> FASTA) is
helen_b200's own: StitchInterface.perform_stitch over the host library behind include/helen_stitch.h.
"""
import sys
import time

from .CallConsensusInterface import call_consensus
from .FileManager import FileManager
from .StitchInterface import perform_stitch
from .TextColor import TextColor


def get_elapsed_time_string(start_time, end_time):
    elapsed = end_time - start_time
    return "{} HOURS {} MINS {} SECS.".format(int(elapsed / 3600), int(elapsed % 3600 / 60), int(elapsed % 60))


def polish_genome(image_dir, model_path, batch_size, num_workers, threads, output_dir, output_prefix, gpu_mode,
                  device_ids, callers):
    output_dir = FileManager.handle_output_directory(output_dir)
    timestr = time.strftime("%m%d%Y_%H%M%S")
    prediction_output_directory = FileManager.handle_output_directory(output_dir + "/predictions_" + str(timestr) + "/")
    sys.stderr.write(TextColor.GREEN + "INFO: RUN-ID: " + str(timestr) + "\n" + TextColor.END)
    sys.stderr.write(TextColor.GREEN + "INFO: PREDICTION OUTPUT DIRECTORY: " + str(prediction_output_directory) + "\n" + TextColor.END)

    t0 = time.time()
    sys.stderr.write(TextColor.GREEN + "INFO: CALL CONSENSUS STARTING\n" + TextColor.END)
    call_consensus(image_dir, model_path, batch_size, num_workers, threads, prediction_output_directory,
                   output_prefix, gpu_mode, device_ids, callers)
    t1 = time.time()
    sys.stderr.write(TextColor.GREEN + "INFO: STITCH STARTING\n" + TextColor.END)
    perform_stitch(prediction_output_directory, output_dir, output_prefix, threads)
    t2 = time.time()
    sys.stderr.write(TextColor.GREEN + "INFO: FINISHED PROCESSING.\n" + TextColor.END)
    sys.stderr.write(TextColor.GREEN + "INFO: TOTAL TIME ELAPSED: " + get_elapsed_time_string(t0, t2) + "\n" + TextColor.END)
    sys.stderr.write(TextColor.GREEN + "INFO: PREDICTION TIME: " + get_elapsed_time_string(t0, t1) + "\n" + TextColor.END)
    sys.stderr.write(TextColor.GREEN + "INFO: STITCH TIME: " + get_elapsed_time_string(t1, t2) + "\n" + TextColor.END)
    return prediction_output_directory
