"""Small filesystem helpers with the reference's behaviour (helen/modules/python/FileManager.py)."""
import os
import time


class FileManager:
    @staticmethod
    def handle_output_directory(output_dir):
        """Create the directory if needed and return its absolute path (FileManager.py:10-23)."""
        if output_dir[-1] != "/":
            output_dir += "/"
        if not os.path.exists(output_dir):
            os.mkdir(output_dir)
        return os.path.abspath(output_dir)

    @staticmethod
    def handle_train_output_directory(output_dir):
        """<output_dir>/trained_models_<stamp>/ and its stats_<stamp>/ sub-directory (FileManager.py:26-49)."""
        timestr = time.strftime("%m%d%Y_%H%M%S")
        if output_dir[-1] != "/":
            output_dir += "/"
        if not os.path.exists(output_dir):
            os.mkdir(output_dir)
        model_save_dir = output_dir + "trained_models_" + timestr + "/"
        if not os.path.exists(model_save_dir):
            os.mkdir(model_save_dir)
        stats_directory = model_save_dir + "stats_" + timestr + "/"
        if not os.path.exists(stats_directory):
            os.mkdir(stats_directory)
        return model_save_dir, stats_directory

    @staticmethod
    def get_file_paths_from_directory(directory_path):
        """MarginPolish image files are those whose name ends in 'h5' (FileManager.py:52-61)."""
        return [os.path.join(directory_path, name) for name in os.listdir(directory_path)
                if os.path.isfile(os.path.join(directory_path, name)) and name[-2:] == 'h5']

    @staticmethod
    def chunks(file_names, threads):
        """Consecutive groups of `threads` items (FileManager.py:62-70; the argument is a group size)."""
        return [file_names[i:i + threads] for i in range(0, len(file_names), threads)]
