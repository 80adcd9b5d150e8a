"""Single place where an HDF5 implementation is chosen: h5py when it can be imported, else the package's own
pure-NumPy reader / writer for the subset of the format this path uses (helen_b200/minih5.py).  The predict path's
tensors never need HDF5; only the MarginPolish image reader and the prediction writer / stitch reader do.
HELEN_B200_HDF5=minih5 | h5py forces one of the two.  Tests may substitute `open_file`."""
import os


def backend():
    want = os.environ.get("HELEN_B200_HDF5", "")
    if want != "minih5":
        try:
            import h5py  # noqa: F401
            return "h5py"
        except ImportError:
            if want == "h5py":
                raise
    return "minih5"


def open_file(path, mode='r'):
    if backend() == "h5py":
        import h5py
        return h5py.File(path, mode)
    from . import minih5
    return minih5.File(path, mode)
