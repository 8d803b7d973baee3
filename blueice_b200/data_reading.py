"""Resolve file-valued settings (host side, cold path; mirrors blueice/data_reading.py:25-51)."""
import pandas as pd

from .utils import find_file_in_folders, read_pickle

__all__ = ['read_files_in']

_FILE_CACHE = {}      # path -> loaded object, process wide (reference: data_reading.CACHE)
_READERS = {'.pkl': read_pickle, '.csv': pd.read_csv}


def read_files_in(config, data_dirs=1):
    """Copy of `config` where every string value ending in .pkl / .csv is replaced by the file contents."""
    out = {}
    for key, value in config.items():
        reader = None
        if isinstance(value, str):
            for ext, fn in _READERS.items():
                if value.endswith(ext):
                    reader = fn
        if reader is None:
            out[key] = value
            continue
        path = find_file_in_folders(value, data_dirs)
        if path not in _FILE_CACHE:
            _FILE_CACHE[path] = reader(path)
        out[key] = _FILE_CACHE[path]
    return out
