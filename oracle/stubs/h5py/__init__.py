"""Import-time stand-in for h5py (absent in this image). The oracle never reads snapshots."""


class File:
    def __init__(self, *args, **kwargs):
        raise RuntimeError("h5py stub: HDF5 I/O is not available in the oracle build")
