"""CSV log writer with the reference's interface (wtracker/utils/log_utils.py:5-90): header written on open,
rows as dicts or iterables in column order, values rendered by ``str`` exactly as ``csv.DictWriter`` does."""

from __future__ import annotations

import csv
from typing import Iterable


class CSVLogger:
    def __init__(self, path: str, col_names: list[str], mode: str = "w+"):
        self.path = path
        self.col_names = col_names
        self._file = open(self.path, mode, newline="")
        self._writer = csv.DictWriter(self._file, self.col_names, escapechar=",")
        self._writer.writeheader()
        self.flush()

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.close()

    def close(self):
        if not self._file.closed:
            self._file.flush()
            self._file.close()

    def _to_dict(self, items: Iterable) -> dict:
        return dict(zip(self.col_names, items))

    def write(self, row: dict | Iterable):
        assert self._file.writable()
        self._writer.writerow(row if isinstance(row, dict) else self._to_dict(row))

    def writerows(self, rows: list[dict] | list[Iterable]):
        assert self._file.writable()
        assert len(rows) > 0
        if not isinstance(rows[0], dict):
            rows = [self._to_dict(r) for r in rows]
        self._writer.writerows(rows)

    def flush(self):
        self._file.flush()
