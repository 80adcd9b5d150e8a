"""Constants of the predict, train and stitch paths.  Names and values are the reference's
(helen/modules/python/Options.py:1-29); code elsewhere in this package refers to them by these names."""


class StitchOptions(object):
    """Options.py:1-10.  The four *_PENALTY values are the local-alignment scoring (the match one is a reward)."""
    MATCH_PENALTY, MISMATCH_PENALTY = 4, 6
    GAP_PENALTY, GAP_EXTEND_PENALTY = 8, 2
    OVERLAP_THRESHOLD = 8                              # shortest aligned run accepted as an anchor
    BASE_ERROR_RATE = 0.0                              # extra overlap taken into the alignment, as a fraction
    MIN_SEQUENCE_REQUIRED_FOR_MULTITHREADING = 2       # smallest group of regions handed to one worker
    KMER_SIZE = 15                                     # present in the reference, never read
    label_decoder = dict(enumerate(['', 'A', 'C', 'G', 'T']))      # predicted base label -> letter, 0 = gap


class ImageSizeOptions(object):
    """Options.py:13-21.  One MarginPolish image = SEQ_LENGTH pileup columns of IMAGE_HEIGHT features."""
    IMAGE_HEIGHT, IMAGE_CHANNELS = 90, 1
    SEQ_LENGTH = LABEL_LENGTH = 1000
    SEQ_OVERLAP = 200
    TOTAL_BASE_LABELS, TOTAL_RLE_LABELS = 5, 11


class TrainOptions(object):
    """Options.py:24-29.  The chunk loop: TRAIN_WINDOW columns per model call, advancing by WINDOW_JUMP."""
    TRAIN_WINDOW, WINDOW_JUMP = 100, 50
    GRU_LAYERS, HIDDEN_SIZE = 1, 128
    # weight of each run-length class in the RLE cross-entropy (train.py:121-126)
    CLASS_WEIGHTS = [0.3] + [0.5] * 4 + [0.8, 0.9] + [1.0] * 3 + [0.9]
