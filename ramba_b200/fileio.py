"""File loaders behind `ramba_b200.load` (reference: ramba/fileio.py:21-168, driver side ramba/ramba.py:8930-8945, worker
side RemoteState.load 3929-3956).

A handler knows three things about one file type: the array's (shape, dtype) without reading it, how to read ONE block of
it (the part of the file a rank's shard holds - distributed loads never bring the whole file to one host), and how to read
all of it (types that cannot be read in parts).  Unlike the reference, `read` RETURNS the block as a host array instead of
filling the worker's container in place: the block then goes host -> device through the same staged copy as `fromarray`.

Types: NumPy `.npy` (memory-mapped, distributed), HDF5 (`h5py`, distributed), netCDF (`netCDF4`, distributed), images
(`PIL`, whole-file; channels first like the reference).  A type whose library is not installed is simply not registered,
as in the reference."""
import numpy as np

loaders = {}


class loader:
    """One file type (ramba/fileio.py:21-31): `getinfo(fname, **kw) -> (shape, dtype)`, `read(fname, index, **kw) -> block`
    (index: tuple of slices in the array's coordinates), `readall(fname, **kw) -> whole array`."""

    def __init__(self, ftype, is_dist, getinfo, read, readall, aliases=()):
        self.ftype = ftype
        self.is_dist = is_dist
        self.getinfo = getinfo
        self.read = read
        self.readall = readall
        loaders[ftype] = self
        for a in aliases:
            loaders[a] = self


def get_load_handler(fname, ftype=None):
    """By explicit type or by the file name's extension (ramba/fileio.py:33-40; an unknown type is an error here, the
    reference falls back to HDF5)."""
    if ftype is None:
        ftype = str(fname).split("/")[-1].split(".")[-1].lower()
    if ftype not in loaders:
        raise ValueError("load: unknown file type %r (known: %s)" % (ftype, ", ".join(sorted(loaders))))
    return loaders[ftype]


# ---- NumPy .npy: the header gives shape and dtype, a memory map gives any block without reading the rest
def _npy_map(fname):
    return np.load(fname, mmap_mode="r", allow_pickle=False)


def _npy_getinfo(fname):
    m = _npy_map(fname)
    return tuple(m.shape), m.dtype


def _npy_read(fname, index):
    return np.array(_npy_map(fname)[index])  # (a copy: the map is read-only and goes away)


def _npy_readall(fname):
    return np.load(fname, allow_pickle=False)


loader("npy", True, _npy_getinfo, _npy_read, _npy_readall)


# ---- HDF5 (ramba/fileio.py:42-72)
try:
    import h5py

    def _hdf5_part(fname, nm):
        p = h5py.File(fname, "r")
        if nm:
            for k in nm.split("/"):
                p = p[k]
        return p

    def _hdf5_getinfo(fname, arr_path):
        p = _hdf5_part(fname, arr_path)
        return tuple(p.shape), p.dtype

    def _hdf5_read(fname, index, arr_path):
        p = _hdf5_part(fname, arr_path)
        out = np.empty([s.stop - s.start for s in index], dtype=p.dtype)
        if out.size:
            p.read_direct(out, index)
        return out

    def _hdf5_readall(fname, arr_path):
        return _hdf5_part(fname, arr_path)[:]

    loader("hdf5", True, _hdf5_getinfo, _hdf5_read, _hdf5_readall, aliases=("h5", "h5ad"))
except ImportError:
    pass


# ---- images (ramba/fileio.py:77-100): whole-file, 8-bit, channels first
try:
    from PIL import Image

    def _pil_getinfo(fname):
        img = Image.open(fname)
        c = len(img.getbands())
        return ((c, img.height, img.width) if c > 1 else (img.height, img.width)), np.dtype(np.uint8)

    def _pil_readall(fname):
        arr = np.array(Image.open(fname))
        if arr.ndim > 2:
            arr = np.transpose(arr, (2, 0, 1))  # HxWxC -> CxHxW
        return np.ascontiguousarray(arr)

    def _pil_read(fname, index):
        return np.ascontiguousarray(_pil_readall(fname)[index])

    loader("pil", False, _pil_getinfo, _pil_read, _pil_readall, aliases=("jpg", "jpeg", "png", "tif", "tiff", "bmp"))
except ImportError:
    pass


# ---- netCDF (ramba/fileio.py:104-166)
try:
    import netCDF4

    def _nc_var(fname, var_select):
        p = netCDF4.Dataset(fname)
        only = set(p.variables.keys()) - set(p.dimensions.keys())
        if var_select is None:
            if len(only) != 1:
                raise ValueError("load: the file holds %d variables, pass var_select" % len(only))
            var_select = only.pop()
        elif var_select not in only:
            raise ValueError("load: no variable %r in the file" % (var_select,))
        return p.variables[var_select]

    def _nc_getinfo(fname, var_select=None):
        v = _nc_var(fname, var_select)
        return tuple(v.shape), np.dtype(v.dtype)

    def _nc_read(fname, index, var_select=None):
        return np.ascontiguousarray(_nc_var(fname, var_select)[index])

    def _nc_readall(fname, var_select=None):
        return np.ascontiguousarray(_nc_var(fname, var_select)[:])

    loader("nc", True, _nc_getinfo, _nc_read, _nc_readall)
except ImportError:
    pass
