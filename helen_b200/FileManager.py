"""Small filesystem helpers with the reference's behaviour (helen/modules/python/FileManager.py)."""
import os
import time


def _ensure_directory(path):
    """One level only, like the reference's os.mkdir: a missing parent is an error, not something to create."""
    if not os.path.isdir(path):
        os.mkdir(path)
    return path


class FileManager:
    @staticmethod
    def handle_output_directory(output_dir):
        """Create the directory if needed and return its absolute path, without a trailing slash (FileManager.py:10-23)."""
        return os.path.abspath(_ensure_directory(output_dir if output_dir.endswith("/") else output_dir + "/"))

    @staticmethod
    def handle_train_output_directory(output_dir):
        """-> (<output_dir>/trained_models_<stamp>/, <that>/stats_<stamp>/), both created, both ending in a slash
        because the training loop appends file names to them directly (FileManager.py:26-49, train.py:41-43)."""
        stamp = time.strftime("%m%d%Y_%H%M%S")
        root = _ensure_directory(output_dir if output_dir.endswith("/") else output_dir + "/")
        models = _ensure_directory("%strained_models_%s/" % (root, stamp))
        stats = _ensure_directory("%sstats_%s/" % (models, stamp))
        return models, stats

    @staticmethod
    def get_file_paths_from_directory(directory_path):
        """MarginPolish image files are those whose name ends in 'h5' (FileManager.py:52-61)."""
        candidates = (os.path.join(directory_path, name) for name in os.listdir(directory_path) if name.endswith('h5'))
        return [path for path in candidates if os.path.isfile(path)]

    @staticmethod
    def chunks(file_names, threads):
        """Consecutive groups of `threads` items (FileManager.py:62-70; the argument is a group size)."""
        return [file_names[i:i + threads] for i in range(0, len(file_names), threads)]
