"""perform_stitch with the reference's signature (helen/modules/python/StitchInterface.py:40-105):
every contig found in the prediction files of a directory -> one record of <output_prefix>.fa."""
import os
import sys
from os import listdir
from os.path import isfile, join

from . import _feed_native
from .DataStore import open_predictions, region_reader
from .FileManager import FileManager
from .Stitch import Stitch
from .TextColor import TextColor


def get_file_paths_from_directory(directory_path):
    """Prediction files are those whose name ends in 'hdf' (StitchInterface.py:29-37)."""
    return [os.path.abspath(join(directory_path, file)) for file in listdir(directory_path)
            if isfile(join(directory_path, file)) and file[-3:] == 'hdf']


def perform_stitch(input_directory, output_path, output_prefix, threads):
    all_prediction_files = get_file_paths_from_directory(input_directory)

    # one pass over the files: contig -> [(file, region key, start, end)] (the reference reopens every file per contig)
    regions_of = dict()
    for prediction_file in sorted(all_prediction_files):
        native_reader = region_reader(prediction_file)
        if native_reader is not None:
            # the same listing in one library call per contig (a genome has millions of regions)
            try:
                listed = []
                for contig in native_reader.list_predictions():
                    names, starts, ends = native_reader.list_predictions(contig)
                    order = sorted(range(len(names)), key=names.__getitem__)
                    listed.append((contig, [(prediction_file, names[i], int(starts[i]), int(ends[i])) for i in order]))
                for contig, entries in listed:
                    regions_of.setdefault(contig, []).extend(entries)
                continue
            except (_feed_native.Unsupported, IOError):
                pass                                                  # a packed file, or one outside the library's subset
        with open_predictions(prediction_file) as hdf5_file:
            if 'predictions' not in hdf5_file:
                raise ValueError(TextColor.RED + "ERROR: INVALID HDF5 FILE, FILE DOES NOT CONTAIN predictions KEY.\n"
                                 + TextColor.END)
            predictions = hdf5_file['predictions']
            for contig in predictions.keys():
                regions = regions_of.setdefault(contig, [])
                for chunk_key in sorted(predictions[contig].keys()):
                    regions.append((prediction_file, chunk_key,
                                    predictions[contig][chunk_key]['contig_start'][()],
                                    predictions[contig][chunk_key]['contig_end'][()]))

    output_dir = FileManager.handle_output_directory(output_path)
    output_filename = os.path.join(output_dir, output_prefix + '.fa')
    sys.stderr.write(TextColor.GREEN + "INFO: OUTPUT FILE: " + output_filename + "\n" + TextColor.END)
    with open(output_filename, 'w') as consensus_fasta_file:
        for i, contig in enumerate(sorted(regions_of)):
            log_prefix = "{:04d}".format(i) + "/" + "{:04d}".format(len(regions_of)) + ":"
            sys.stderr.write(TextColor.GREEN + "INFO: " + str(log_prefix) + " PROCESSING CONTIG: " + contig + "\n"
                             + TextColor.END)
            consensus_sequence = Stitch().create_consensus_sequence(contig, regions_of[contig], threads)
            sys.stderr.write(TextColor.BLUE + "INFO: " + str(log_prefix) + " FINISHED PROCESSING " + contig
                             + ", POLISHED SEQUENCE LENGTH: " + str(len(consensus_sequence)) + ".\n" + TextColor.END)
            if consensus_sequence is not None and len(consensus_sequence) > 0:
                consensus_fasta_file.write('>' + contig + "\n")
                consensus_fasta_file.write(consensus_sequence + "\n")
    return output_filename
