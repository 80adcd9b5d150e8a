"""Child process of tests/test_feed_native.py::test_corrupted_files_do_not_crash_the_reader: opens every file it is given with
the native feed library and reads everything it lists; errors are fine, crashes are not (the parent checks the exit code)."""
import sys

from helen_b200 import _feed_native

outcomes = {"ok": 0, "error": 0}
for path in sys.argv[1:]:
    try:
        handle = _feed_native.ImageFile(path)
        n = len(handle)
        handle.names()
        if n:
            handle.read_block(0, min(n, 8), 1000, 2)
        try:                                           # prediction files: the stitch's listing and region reads
            for contig in handle.list_predictions()[:3]:
                names, _, _ = handle.list_predictions(contig)
                for region in names[:4] + ["no-such-region"]:
                    try:
                        handle.read_prediction_region(contig, region)
                    except _feed_native.Unsupported:
                        pass
        except _feed_native.Unsupported:
            pass
        handle.close()
        outcomes["ok"] += 1
    except (IOError, ValueError, _feed_native.Unsupported):
        outcomes["error"] += 1
print(outcomes)
