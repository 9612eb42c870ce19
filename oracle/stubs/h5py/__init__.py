"""Stand-in for h5py (absent in this image, no network) -- TEST INFRASTRUCTURE ONLY.

Just enough of the h5py.File interface for the reference's own Gadget-HDF5 reader to run unmodified
(/root/reference/cosmic_profiles/gadget/readgadget.py:34-52, :107-116): a file is a NumPy .npz archive whose keys are
the dataset paths ("PartType1/Coordinates", ...) and "Header.attrs/<name>" for the header attributes; write one with
`write_stub_file`.  Datasets support `[:]` and slices, groups support `.attrs[...]`, files support `in` and close().
Nothing of the product imports this."""
import numpy as np


class _Attrs:
    def __init__(self, data, group):
        self._data, self._prefix = data, group + ".attrs/"

    def __getitem__(self, name):
        v = self._data[self._prefix + name]
        return v[()] if v.ndim == 0 else v

    def __contains__(self, name):
        return (self._prefix + name) in self._data


class _Group:
    def __init__(self, data, name):
        self.attrs = _Attrs(data, name)
        self._data, self._name = data, name

    def __getitem__(self, key):
        return File._lookup(self._data, self._name + "/" + key)

    def __contains__(self, key):
        return File._has(self._data, self._name + "/" + key)


class _Dataset:
    def __init__(self, arr):
        self._arr = arr
        self.shape, self.dtype = arr.shape, arr.dtype

    def __getitem__(self, sl):
        return np.array(self._arr[sl])     # h5py returns a fresh array of the stored dtype

    def __len__(self):
        return len(self._arr)


class File:
    def __init__(self, name, mode="r", **kwargs):
        if mode != "r":
            raise RuntimeError("h5py stub: read-only (write with h5py.write_stub_file)")
        with np.load(name, allow_pickle=False) as z:
            self._data = {k: z[k] for k in z.files}
        self.filename = name

    @staticmethod
    def _has(data, key):
        key = key.strip("/")
        return key in data or any(k.startswith(key + "/") or k.startswith(key + ".attrs/") for k in data)

    @staticmethod
    def _lookup(data, key):
        key = key.strip("/")
        if key in data:
            return _Dataset(data[key])
        if File._has(data, key):
            return _Group(data, key)
        raise KeyError("Unable to open object (object '%s' doesn't exist)" % key)

    def __getitem__(self, key):
        return File._lookup(self._data, key)

    def __contains__(self, key):
        return File._has(self._data, key)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def write_stub_file(name, datasets, header_attrs):
    """datasets: {"PartType1/Coordinates": array, ...}; header_attrs: {"Time": ..., "MassTable": ..., ...}."""
    out = dict(datasets)
    for k, v in header_attrs.items():
        out["Header.attrs/" + k] = np.asarray(v)
    with open(name, "wb") as f:      # np.savez would append ".npz" to a bare file name
        np.savez(f, **out)
