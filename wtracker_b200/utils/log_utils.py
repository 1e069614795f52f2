"""Text writer of the tracking log (``bboxes.csv``).  Presents the interface of the reference's ``CSVLogger``
(wtracker/utils/log_utils.py:5-90: header on open, ``write`` / ``writerows`` of dicts or column-ordered iterables,
``flush``, ``close``, context manager) but is built for the batched path: rows are rendered to text in one pass per
call — a whole cycle, or the whole table of a lock-step sweep — and handed to the file as a single string.  Cells are
``str(value)`` joined by commas with ``\\r\\n`` line ends, i.e. the bytes Python's csv module produces for the numeric
and plain-string cells of the log (tests/test_log_cpu.py compares with files written by the unmodified reference)."""

from __future__ import annotations

from typing import Iterable, Mapping, Sequence

_SPECIAL = (",", '"', "\r", "\n")


def _cell(value) -> str:
    text = "" if value is None else str(value)
    if any(ch in text for ch in _SPECIAL):          # minimal quoting, quotes doubled
        text = '"' + text.replace('"', '""') + '"'
    return text


class CSVLogger:
    def __init__(self, path: str, col_names: list[str], mode: str = "w+"):
        self.path = path
        self.col_names = list(col_names)
        self._known = set(self.col_names)
        self._file = open(path, mode, newline="")
        self._file.write(",".join(_cell(c) for c in self.col_names) + "\r\n")
        self._file.flush()

    # ---- rendering ----------------------------------------------------------------------------------------------
    def _line(self, row: Mapping | Iterable) -> str:
        if isinstance(row, Mapping):
            extra = [k for k in row if k not in self._known]
            if extra:
                raise ValueError(f"dict contains fields not in col_names: {extra}")
            cells = [_cell(row.get(c, "")) for c in self.col_names]
        else:
            cells = [_cell(v) for _, v in zip(self.col_names, row)]
            cells += [""] * (len(self.col_names) - len(cells))
        return ",".join(cells) + "\r\n"

    # ---- the reference's surface --------------------------------------------------------------------------------
    def write(self, row: Mapping | Iterable) -> None:
        assert self._file.writable()
        self._file.write(self._line(row))

    def writerows(self, rows: Sequence[Mapping] | Sequence[Iterable]) -> None:
        assert self._file.writable()
        assert len(rows) > 0
        self._file.write("".join(self._line(r) for r in rows))

    def flush(self) -> None:
        self._file.flush()

    def close(self) -> None:
        if not self._file.closed:
            self._file.flush()
            self._file.close()

    def __enter__(self) -> "CSVLogger":
        return self

    def __exit__(self, exc_type, exc_value, traceback) -> None:
        self.close()
