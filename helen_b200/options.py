"""Constants of the predict and stitch paths (mirrors helen/modules/python/Options.py:1-29)."""


class StitchOptions(object):
    BASE_ERROR_RATE = 0.0      # Options.py:2
    label_decoder = {1: 'A', 2: 'C', 3: 'G', 4: 'T', 0: ''}   # Options.py:3
    MATCH_PENALTY = 4          # Options.py:4  (a score, despite the name)
    MISMATCH_PENALTY = 6       # Options.py:5
    GAP_PENALTY = 8            # Options.py:6
    GAP_EXTEND_PENALTY = 2     # Options.py:7
    MIN_SEQUENCE_REQUIRED_FOR_MULTITHREADING = 2   # Options.py:8
    OVERLAP_THRESHOLD = 8      # Options.py:9  shortest aligned run accepted as an anchor
    KMER_SIZE = 15             # Options.py:10 (unused by the reference)


class ImageSizeOptions(object):
    IMAGE_HEIGHT = 90          # Options.py:14  features per pileup column
    IMAGE_CHANNELS = 1         # Options.py:15
    SEQ_LENGTH = 1000          # Options.py:16  columns per window
    SEQ_OVERLAP = 200          # Options.py:17
    LABEL_LENGTH = SEQ_LENGTH
    TOTAL_BASE_LABELS = 5      # Options.py:20
    TOTAL_RLE_LABELS = 11      # Options.py:21


class TrainOptions(object):
    TRAIN_WINDOW = 100         # Options.py:25  chunk width
    WINDOW_JUMP = 50           # Options.py:26  chunk stride
    GRU_LAYERS = 1             # Options.py:27
    HIDDEN_SIZE = 128          # Options.py:28
    CLASS_WEIGHTS = [0.3, 0.5, 0.5, 0.5, 0.5, 0.8, 0.9, 1.0, 1.0, 1.0, 0.9]   # Options.py:29
