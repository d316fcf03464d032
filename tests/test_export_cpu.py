"""The NetCDF container either side of the hot path (hyperscanning_signal_analysis_b200/export.py; reference
src/export.py:246-288, :606, src/ncdf.py:5-88, src/eeg_alpha_ibi_ffdtf.py:244-266).  CPU only."""
import json

import numpy as np
import pytest

from hyperscanning_signal_analysis_b200 import export


def _sample():
    rng = np.random.default_rng(7)
    channels = ["Fp1", "Fp2", "F3", "F4", "Cz", "Oz_long_name"]
    t = np.arange(640) / 128.0
    data = rng.standard_normal((t.size, len(channels))) * 20.0
    events = {"Peppa": {"start": 100.0, "duration": 3.0}, "Brave": {"start": 103.0, "duration": 2.0}}
    meta = {"filtration": {"notch": {"Q": 30, "freq": 50.0, "applied": True}}, "references": "Original reference retained", "bad": None}
    attrs = export.signal_attrs("W_030", "ch", "EEG", 128.0, "movies", 100.0, 105.0, 0.5, channels, meta, ["Peppa", "Brave"], events)
    return data, t, channels, attrs


def test_attribute_schema_matches_the_reference_order_and_forms():
    _, _, channels, attrs = _sample()
    assert tuple(attrs) == export.ATTR_ORDER                                   # export.py:270-286
    assert attrs["units"] == "μV" and attrs["task_start"] == 0.0 and attrs["task_duration"] == 5.0
    assert attrs["channel_names_csv"] == ",".join(channels) and json.loads(attrs["channel_names_json"]) == channels
    ev = json.loads(attrs["task_events_structure"])                             # list of dicts -> JSON text (ncdf.py:52-58)
    assert ev[1] == {"name": "Brave", "start_s": 103.0, "start_rel_s": 3.0, "duration_s": 2.0}
    assert export.sanitize_attr(None) == "" and export.sanitize_attr([1, None, "a"]) == [1, "", "a"]
    assert export.sanitize_attr(np.arange(3)) == [0, 1, 2] and export.sanitize_attr({"a": 1}) == '{"a": 1}'
    assert export.signal_attrs("d", "cg", "diode", 1.0, "t", 0.0, 1.0, 0.0, ["x"], {}, [], {})["units"] == "unknown"


def test_netcdf3_round_trip_and_container_layout(tmp_path):
    data, t, channels, attrs = _sample()
    path = export.write_netcdf3(tmp_path / "W_030_EEG_ch_movies.nc", data, t, channels, attrs)
    with open(path, "rb") as fh:
        assert fh.read(4) == b"CDF\x02"                                         # classic container, 64-bit offsets
    from scipy.io import netcdf_file
    with netcdf_file(path, "r", mmap=False) as nc:                              # the layout xarray expects for a DataArray
        assert set(nc.dimensions) == {"time", "channel", "string12"} and nc.dimensions["time"] == t.size
        assert nc.variables["signals"].dimensions == ("time", "channel") and nc.variables["signals"].typecode() == "d"
        assert nc.variables["channel"].dimensions == ("channel", "string12") and nc.variables["channel"]._Encoding == b"utf-8"
        assert nc.variables["signals"].units.decode("utf-8") == "μV"
        assert list(nc.variables["signals"]._attributes) == list(export.ATTR_ORDER)
    with export.open_dataarray(path) as da:                                     # what _load_eeg_and_ibi does with it
        assert da.dims == ("time", "channel")
        assert np.array_equal(np.asarray(da.values), data) and np.array_equal(np.asarray(da.coords["time"].values), t)
        assert [str(c) for c in da.coords["channel"].values.tolist()] == channels
        assert float(da.attrs["sampling_freq"]) == 128.0 and da.attrs["dyad_id"] == "W_030"
    dec = export.read_netcdf3(path, decode_json_attrs=True)                     # load_xarray_from_netcdf(decode_json_attrs=True)
    assert dec.attrs["task_event_names_json"] == ["Peppa", "Brave"] and dec.attrs["metadata_json"]["filtration"]["notch"]["Q"] == 30


def test_errors(tmp_path):
    data, t, channels, attrs = _sample()
    with pytest.raises(ValueError):
        export.write_netcdf3(tmp_path / "bad.nc", data[:, :3], t, channels, attrs)
    hdf = tmp_path / "hdf.nc"
    hdf.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(OSError):
        export.read_netcdf3(hdf)


def test_pipeline_loader_reads_the_files(tmp_path):
    """EEG_IBI_FFDTF_Pipeline._load_eeg_and_ibi (src/eeg_alpha_ibi_ffdtf.py:244-266) on files in this container: directory scan,
    file matching and the (time, eeg, fs, names, ibi, fs, duration) tuple, without a GPU."""
    from hyperscanning_signal_analysis_b200.eeg_alpha_ibi_ffdtf import EEG_IBI_FFDTF_Pipeline
    rng = np.random.default_rng(3)
    names = ["F3", "F4", "Cz"]
    t = np.arange(256) / 128.0
    for kind, chans, fs in (("EEG", names, 128.0), ("IBI", ["IBI"], 4.0)):
        for role in ("ch", "cg"):
            folder = tmp_path / kind / "W_001"
            folder.mkdir(parents=True, exist_ok=True)
            n = t.size if kind == "EEG" else 8
            attrs = {"sampling_freq": fs, "event_duration_s": 2.0, "who": role}
            export.write_netcdf3(folder / f"W_001_{kind}_{role}_Peppa.nc", rng.standard_normal((n, len(chans))), np.arange(n) / fs, chans, attrs)
    pipe = EEG_IBI_FFDTF_Pipeline(cleaned_signals_folder=tmp_path, output_ffDTF_folder=tmp_path / "out", target_events=["Peppa"])
    assert pipe.dyads_to_process == ["W_001"] and len(pipe.eeg_files) == 2 and len(pipe.ibi_files) == 2
    eeg_file, ok = pipe._find_file(pipe.eeg_files, "W_001", "Peppa", "ch")
    ibi_file, ok2 = pipe._find_file(pipe.ibi_files, "W_001", "Peppa", "ch")
    assert ok and ok2
    time_s, eeg, fs_eeg, got_names, ibi, fs_ibi, dur = pipe._load_eeg_and_ibi(eeg_file, ibi_file, role="Child")
    assert eeg.shape == (3, 256) and ibi.shape == (1, 8) and got_names == names
    assert fs_eeg == 128.0 and fs_ibi == 4.0 and dur == 2.0 and np.array_equal(time_s, t)
