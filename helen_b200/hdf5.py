"""Single place where h5py is imported (lazily): the predict path's tensors never need it, only
the MarginPolish image reader and the prediction writer do.  Tests substitute `open_file`."""


def open_file(path, mode='r'):
    try:
        import h5py
    except ImportError as exc:
        raise ImportError("h5py is required to read MarginPolish images / write prediction files "
                          "(it is not needed for WindowPredictor or the benchmarks)") from exc
    return h5py.File(path, mode)
