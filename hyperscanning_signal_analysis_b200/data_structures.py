"""The parts of the reference's ``src/data_structures.py`` that sit on the hot path:
``Filtration`` (data_structures.py:24-57, fields unchanged -- they are serialised into the NetCDF
``metadata_json`` by export.py:343-367) and ``MultimodalData._decimate_signals``
(data_structures.py:722-800).  ``MultimodalData`` here carries the same attribute names as the reference
container (data_structures.py:136-177) so objects can be exchanged field by field; everything that is not
on the path (MNE round trips, IBI/RMSSD, events) stays in the reference.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np
import pandas as pd

from . import frontend


@dataclass
class Filtration:
    """Stores information about signal filtration (same keys as the reference)."""
    notch: Dict[str, Any] = field(default_factory=lambda: {"Q": None, "freq": None, "a": None, "b": None, "applied": False})
    low_pass: Dict[str, Any] = field(default_factory=lambda: {"type": None, "cut_f": None, "order": None, "f_type": None,
                                                               "a": None, "b": None, "applied": False})
    high_pass: Dict[str, Any] = field(default_factory=lambda: {"type": None, "cut_f": None, "order": None, "f_type": None,
                                                                "a": None, "b": None, "applied": False})


_DECIMATE_PREFIXES = ("EEG_ch_", "EEG_cg_", "ECG", "IBI", "RMSSD", "ET_ch_", "ET_cg_")


class MultimodalData:
    """Field-compatible stand-in for the reference container (data_structures.py:113-177)."""

    def __init__(self):
        self.data: pd.DataFrame = pd.DataFrame()
        self.fs: Optional[float] = None
        self.id: Optional[str] = None
        self.eeg_channel_names: List[str] = []
        self.eeg_channel_mapping: Dict[str, int] = {}
        self.references: Optional[str] = None
        self.eeg_filtration: Filtration = Filtration()
        self.eeg_channel_names_ch: List[str] = []
        self.eeg_channel_names_cg: List[str] = []
        self.events: Dict[str, Any] = {}
        self.epoch: Optional[List[Any]] = None
        self.paths: Any = None
        self.tasks: Any = None
        self.modalities: List[str] = []
        self.child_info: Any = None
        self.notes: Optional[str] = None

    def eeg_channel_names_all(self):
        return list(self.eeg_channel_names_ch) + list(self.eeg_channel_names_cg)

    def _decimate_signals(self, q=8):
        """Reference ``_decimate_signals`` (data_structures.py:722-800): new object, fs/q, ``[::q]`` on
        time/events/diode, FIR anti-alias decimation of every signal column (all columns in ONE batched
        GPU call), NaN ffill/bfill before and NaN-mask restore after."""
        dec = MultimodalData()
        dec.fs = self.fs / q
        dec.id = self.id
        dec.eeg_channel_names_ch = list(self.eeg_channel_names_ch)
        dec.eeg_channel_names_cg = list(self.eeg_channel_names_cg)
        dec.eeg_channel_mapping = dict(self.eeg_channel_mapping)
        dec.references = self.references
        dec.eeg_filtration = copy.deepcopy(self.eeg_filtration)
        dec.events = copy.deepcopy(self.events)
        dec.paths = copy.deepcopy(self.paths)
        dec.tasks = copy.deepcopy(self.tasks)
        dec.modalities = list(self.modalities)
        dec.child_info = copy.deepcopy(self.child_info)
        dec.notes = self.notes

        dec.data['time'] = self.data['time'].values[::q]
        dec.data['time_idx'] = self.data['time_idx'].values[::q]
        for col in ('events', 'ET_event', 'EEG_events', 'diode'):
            if col in self.data.columns:
                dec.data[col] = self.data[col].values[::q]

        cols = [c for c in self.data.columns if c.startswith(_DECIMATE_PREFIXES)]
        if not cols:
            return dec
        stack = np.empty((len(cols), len(self.data)), dtype=np.float64)
        masks = {}
        for r, col in enumerate(cols):
            series = self.data[col]
            is_nan = series.isnull()
            if is_nan.any():
                print(f'Column {col} contains NaN values, applying forward and backward fill before decimation.')
                filled = series.infer_objects(copy=False).ffill().bfill()
                if filled.isnull().any():
                    filled = filled.fillna(0.0)
                stack[r] = filled.values.astype(float)
                masks[r] = is_nan.values[::q]
            else:
                stack[r] = series.values.astype(float)
        out = frontend.decimate(stack, q)
        for r, col in enumerate(cols):
            y = out[r]
            if r in masks:
                y = y.copy()
                y[masks[r]] = np.nan
            dec.data[col] = y
        return dec
