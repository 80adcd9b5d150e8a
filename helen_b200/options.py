"""Constants of the predict path (mirrors helen/modules/python/Options.py:13-29)."""


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
