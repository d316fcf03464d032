"""The on-disk format either side of the hot path: one annotated ``signals (time, channel)`` array per
(dyad, member, modality, task) file, as the reference builds it in ``_build_dataarray`` (src/export.py:246-288) and
writes it with ``DataArray.to_netcdf`` (src/export.py:606), and as ``EEG_IBI_FFDTF_Pipeline._load_eeg_and_ibi``
(src/eeg_alpha_ibi_ffdtf.py:244-266) and ``load_xarray_from_netcdf`` (src/ncdf.py:5-20) read it back.

What is kept: the schema -- variable ``signals`` with dimensions ``(time, channel)``, coordinate variables ``time``
(float64 seconds from the start of the task) and ``channel`` (names), and the attribute set of export.py:270-286 in that
order, sanitised like ``_sanitize_netcdf_attrs_inplace`` (src/ncdf.py:45-72: None -> "", dict / nested list -> JSON text,
flat list kept, anything else -> str).

What differs, and why: the reference asks xarray for ``engine='netcdf4', format='NETCDF4_CLASSIC'``, an HDF5 container that
needs libhdf5 (netCDF4 / h5py), which this environment does not have.  This writer emits the classic NetCDF-3 container
(64-bit offsets) through ``scipy.io.netcdf_file`` with xarray's own conventions for that container (string coordinates as
``char`` arrays over a ``string<N>`` dimension with ``_Encoding = "utf-8"``), which is what xarray itself writes when
netCDF4 is absent; ``xarray.open_dataarray`` and MATLAB's ``ncread`` / ``ncreadatt`` (matlab_utils/ncdf_test_read_demo.m)
read both containers.  Byte-level parity with the reference's files is therefore NOT claimed.

Host-side only: nothing here touches the GPU."""
from __future__ import annotations

import json
import numbers

import numpy as np

ATTR_ORDER = ("dyad_id", "who", "modality", "units", "sampling_freq", "task_name", "task_start", "task_duration", "time_margin_s",
              "channel_names_csv", "channel_names_json", "metadata_json", "task_event_names_csv", "task_event_names_json",
              "task_events_structure")
_UNITS = {"EEG": "μV", "ECG": "μV", "ET": "px", "IBI": "ms", "RMSSD": "ms"}      # export.py:261-267


def sanitize_attr(value):
    """One attribute value made storable (semantics of src/ncdf.py:45-67)."""
    if value is None:
        return ""
    if isinstance(value, (str, bytes, bool, numbers.Number)):
        return value
    if isinstance(value, dict):
        return json.dumps(value, ensure_ascii=False, default=str)
    if isinstance(value, (list, tuple)):
        if any(isinstance(v, (dict, list, tuple)) for v in value):
            return json.dumps(value, ensure_ascii=False, default=str)
        return ["" if v is None else v for v in value]
    if hasattr(value, "tolist"):
        return sanitize_attr(value.tolist())
    return str(value)


def events_structure(events, ordered_event_names, chunk_start):
    """``_build_events_structure`` (export.py:328-340): name / absolute start / start relative to the chunk / duration."""
    out = []
    for name in ordered_event_names:
        ev = events.get(name, {})
        start_abs = float(ev.get("start", 0.0))
        out.append({"name": name, "start_s": start_abs, "start_rel_s": start_abs - chunk_start, "duration_s": float(ev.get("duration", 0.0))})
    return out


def signal_attrs(dyad_id, member, modality, fs, chunk_name, chunk_start, chunk_end, time_margin, channels, metadata, ordered_events, events):
    """The attribute dict of one exported array, keys and value forms of export.py:270-286, already sanitised."""
    attrs = {
        "dyad_id": dyad_id,
        "who": member,
        "modality": modality,
        "units": _UNITS.get(modality, "unknown"),
        "sampling_freq": float(fs),
        "task_name": chunk_name,
        "task_start": 0.0,
        "task_duration": float(chunk_end - chunk_start),
        "time_margin_s": float(time_margin),
        "channel_names_csv": ",".join(channels),
        "channel_names_json": json.dumps(list(channels), ensure_ascii=True),
        "metadata_json": json.dumps(metadata, ensure_ascii=False, default=str),
        "task_event_names_csv": ",".join(ordered_events),
        "task_event_names_json": json.dumps(list(ordered_events), ensure_ascii=True),
        "task_events_structure": events_structure(events, ordered_events, chunk_start),
    }
    return {k: sanitize_attr(v) for k, v in attrs.items()}


class SignalArray:
    """What the callers use of an ``xarray.DataArray`` read from such a file: ``values`` (time, channel), ``coords['time']`` /
    ``coords['channel']`` with ``.values``, ``attrs``, ``dims``, ``name``."""

    class _Coord:
        def __init__(self, values):
            self.values = values

    def __init__(self, values, time, channels, attrs, name="signals"):
        self.values = np.asarray(values)
        self.coords = {"time": SignalArray._Coord(np.asarray(time, dtype=np.float64)), "channel": SignalArray._Coord(np.asarray(list(channels), dtype=object))}
        self.attrs = dict(attrs)
        self.dims = ("time", "channel")
        self.name = name

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _attr_to_nc(value):
    """NetCDF-3 attribute payload: UTF-8 text, a numeric scalar / 1-D numeric array, or (for a flat list of strings, which the
    classic container cannot hold) its JSON text."""
    if isinstance(value, bytes):
        return value
    if isinstance(value, str):
        return value.encode("utf-8")
    if isinstance(value, bool):
        return np.int8(value)
    if isinstance(value, numbers.Number):
        return value
    if isinstance(value, (list, tuple)):
        if all(isinstance(v, numbers.Number) and not isinstance(v, bool) for v in value) and len(value):
            return np.asarray(value)
        return json.dumps(list(value), ensure_ascii=False, default=str).encode("utf-8")
    return str(value).encode("utf-8")


def write_netcdf3(path, data, time, channels, attrs, name="signals"):
    """Write ``data (n_time, n_channel)`` with its coordinates and attributes as a classic NetCDF file (see the module docstring)."""
    from scipy.io import netcdf_file
    data = np.ascontiguousarray(data, dtype=np.float64)
    time = np.ascontiguousarray(time, dtype=np.float64)
    channels = [str(c) for c in channels]
    if data.ndim != 2 or data.shape != (time.shape[0], len(channels)):
        raise ValueError(f"data must be (n_time, n_channel) = ({time.shape[0]}, {len(channels)}), got {data.shape}")
    enc = [c.encode("utf-8") for c in channels]
    width = max([len(b) for b in enc] + [1])
    with netcdf_file(str(path), "w", version=2) as nc:
        nc.createDimension("time", data.shape[0])
        nc.createDimension("channel", len(channels))
        nc.createDimension(f"string{width}", width)
        vt = nc.createVariable("time", "d", ("time",))
        vt[:] = time
        vc = nc.createVariable("channel", "c", ("channel", f"string{width}"))
        chars = np.zeros((len(channels), width), dtype="S1")
        for i, b in enumerate(enc):
            chars[i, :len(b)] = np.frombuffer(b, dtype="S1")
        vc[:] = chars
        vc._Encoding = b"utf-8"
        vs = nc.createVariable(name, "d", ("time", "channel"))
        vs[:] = data
        for key, value in attrs.items():
            setattr(vs, key, _attr_to_nc(sanitize_attr(value)))
    return str(path)


def _decode(value):
    if isinstance(value, bytes):
        return value.decode("utf-8")
    if isinstance(value, np.ndarray) and value.ndim == 0:
        return value.item()
    if isinstance(value, np.ndarray) and value.size == 1:
        return value.reshape(()).item()
    return value


def read_netcdf3(path, decode_json_attrs=False):
    """Read a file written by ``write_netcdf3`` (or by xarray's scipy engine) into a ``SignalArray``.  ``decode_json_attrs`` parses
    attribute texts that start with ``[`` or ``{`` like ``load_xarray_from_netcdf`` (src/ncdf.py:5-20, :70-88)."""
    from scipy.io import netcdf_file
    with open(str(path), "rb") as fh:
        magic = fh.read(4)
    if magic[:3] != b"CDF":
        raise OSError(f"{path}: not a classic NetCDF-3 file (magic {magic!r}); the reference's NETCDF4_CLASSIC files are HDF5 containers "
                      f"and need xarray with netCDF4 or h5netcdf")
    with netcdf_file(str(path), "r", mmap=False) as nc:
        data_vars = [n for n in nc.variables if n not in nc.dimensions]
        if len(data_vars) != 1:
            raise ValueError(f"{path}: expected exactly one data variable, found {data_vars}")
        var = nc.variables[data_vars[0]]
        values = np.array(var[:], dtype=np.float64)
        time = np.array(nc.variables["time"][:], dtype=np.float64)
        raw = np.array(nc.variables["channel"][:])
        channels = [b"".join(row.tolist()).rstrip(b"\x00").decode("utf-8") for row in raw.reshape(raw.shape[0], -1)]
        attrs = {k: _decode(v) for k, v in var._attributes.items()}
    if decode_json_attrs:
        for k, v in list(attrs.items()):
            if isinstance(v, str) and v.strip()[:1] in ("[", "{"):
                try:
                    attrs[k] = json.loads(v)
                except ValueError:
                    pass
    return SignalArray(values, time, channels, attrs, name=data_vars[0])


def open_dataarray(path):
    """``xarray.open_dataarray`` where xarray (and a backend for the file's container) exists, ``read_netcdf3`` otherwise."""
    try:
        import xarray as xr
    except ImportError:
        return read_netcdf3(path)
    try:
        return xr.open_dataarray(path)
    except (ValueError, OSError, ImportError):
        return read_netcdf3(path)
