"""Kaldi matrix archives (ark / scp) -- the on-disk format on the input side of the hot path.

The reference reads its features through the external `kaldi_io` package (`kaldi_io.read_mat(rxfile)` per utterance,
U/BatchLoader.py:45-50; `kaldi_io.read_mat_scp`, L/initialize_model.py:58; `kaldi_io.read_mat(lda.mat)`,
L/initialize_model.py:69).  That package is a pip dependency that is not vendored in the reference tree and is absent
from this image, so this is a from-scratch reader/writer of Kaldi's published table format with the same entry points
(`read_mat`, `read_mat_scp`, `read_mat_ark`, `write_mat`).  PARITY UNPINNED: there are neither Kaldi binaries nor
Kaldi-written files here to check against; `tests/test_loader.py` pins it against byte strings assembled by hand from
the format description below and against round trips.

Format (little endian):

    archive  := { key ' ' object }
    scp line := key ' ' path[':' byte-offset-of-object]['[' r0 ':' r1 [',' c0 ':' c1] ']']     (ranges are inclusive)
    binary object  := '\\0' 'B' token ' ' ...
        FM / DM    := '\\4' int32 rows '\\4' int32 cols, rows*cols float32 / float64, row major
        FV / DV    := '\\4' int32 dim, dim float32 / float64                                 (returned as [1, dim])
        CM         := float32 min, float32 range, int32 rows, int32 cols,
                      cols x (4 x uint16 percentiles 0/25/75/100), then uint8 [cols][rows] (column major)
        CM2 / CM3  := the same global header, then uint16 / uint8 [rows][cols]
    text object    := ' [' { newline row } ' ]' newline

Archives are memory-mapped once and kept open (scp files address thousands of utterances inside a few arks in random
order); every matrix is returned as a fresh C-contiguous array that owns its memory.
"""
from __future__ import annotations

import mmap
import os
import re
import struct
import threading
from collections import OrderedDict
from typing import Iterator, Optional, Tuple

import numpy as np

__all__ = ["read_mat", "read_mat_scp", "read_mat_ark", "read_scp", "write_mat", "write_ark_scp", "ArkReader",
           "KaldiFormatError"]


class KaldiFormatError(ValueError):
    pass


_RANGE = re.compile(r"^(.*)\[([0-9]*):([0-9]*)(?:,([0-9]*):([0-9]*))?\]$")
_U16_STEP = np.float32(1.0 / 65535.0)


def _split_rxfilename(rx: str):
    """'path:offset[r0:r1,c0:c1]' -> (path, offset or 0, row range or None, col range or None)."""
    rx = rx.strip()
    if rx.endswith("|") or rx.startswith("|"):
        raise KaldiFormatError("piped rxfilenames are not supported: %r" % rx)
    rows = cols = None
    m = _RANGE.match(rx)
    if m:
        rx = m.group(1)
        if m.group(2) != "" or m.group(3) != "":
            rows = (int(m.group(2) or 0), int(m.group(3)) if m.group(3) != "" else None)
        if m.group(4) is not None and (m.group(4) != "" or m.group(5) != ""):
            cols = (int(m.group(4) or 0), int(m.group(5)) if m.group(5) != "" else None)
    path, offset = rx, 0
    head, sep, tail = rx.rpartition(":")
    if sep and tail.isdigit() and head:
        path, offset = head, int(tail)
    return path, offset, rows, cols


def _apply_range(mat: np.ndarray, rows, cols) -> np.ndarray:
    if rows is not None:
        hi = mat.shape[0] - 1 if rows[1] is None else rows[1]
        if rows[0] > hi or hi >= mat.shape[0]:
            raise KaldiFormatError("row range %r outside a %d-row matrix" % (rows, mat.shape[0]))
        mat = mat[rows[0]:hi + 1]
    if cols is not None:
        hi = mat.shape[1] - 1 if cols[1] is None else cols[1]
        if cols[0] > hi or hi >= mat.shape[1]:
            raise KaldiFormatError("column range %r outside a %d-column matrix" % (cols, mat.shape[1]))
        mat = mat[:, cols[0]:hi + 1]
    return np.ascontiguousarray(mat)


def _sized_int(buf, pos: int) -> Tuple[int, int]:
    if buf[pos] != 4:
        raise KaldiFormatError("expected a 4-byte integer marker at byte %d" % pos)
    return struct.unpack_from("<i", buf, pos + 1)[0], pos + 5


def _decode_cm(buf, pos: int, variant: str) -> Tuple[np.ndarray, int]:
    lo, span, rows, cols = struct.unpack_from("<ffii", buf, pos)
    pos += 16
    lo, span = np.float32(lo), np.float32(span)
    if rows < 0 or cols < 0:
        raise KaldiFormatError("negative size in a compressed matrix header")
    if variant == "CM2":
        q = np.frombuffer(buf, dtype="<u2", count=rows * cols, offset=pos).reshape(rows, cols)
        return lo + (span * _U16_STEP) * q.astype(np.float32), pos + 2 * rows * cols
    if variant == "CM3":
        q = np.frombuffer(buf, dtype=np.uint8, count=rows * cols, offset=pos).reshape(rows, cols)
        return lo + (span * np.float32(1.0 / 255.0)) * q.astype(np.float32), pos + rows * cols
    # CM: per-column quartile headers, piecewise-linear byte code
    pct = np.frombuffer(buf, dtype="<u2", count=4 * cols, offset=pos).reshape(cols, 4)
    pos += 8 * cols
    p = lo + (span * _U16_STEP) * pct.astype(np.float32)                  # [cols, 4]  p0, p25, p75, p100
    q = np.frombuffer(buf, dtype=np.uint8, count=rows * cols, offset=pos).reshape(cols, rows).astype(np.float32)
    p0, p25, p75, p100 = (p[:, i:i + 1] for i in range(4))
    low = p0 + (p25 - p0) * q * np.float32(1.0 / 64.0)
    mid = p25 + (p75 - p25) * (q - 64.0) * np.float32(1.0 / 128.0)
    top = p75 + (p100 - p75) * (q - 192.0) * np.float32(1.0 / 63.0)
    out = np.where(q <= 64.0, low, np.where(q <= 192.0, mid, top)).astype(np.float32)
    return np.ascontiguousarray(out.T), pos + rows * cols


def _decode_text(buf, pos: int, end: int) -> Tuple[np.ndarray, int]:
    open_br = buf.find(b"[", pos, end)
    close_br = buf.find(b"]", pos, end)
    if open_br < 0 or close_br < open_br:
        raise KaldiFormatError("text matrix without '[ ... ]' at byte %d" % pos)
    body = bytes(buf[open_br + 1:close_br]).decode("ascii")
    rows = [np.array(line.split(), dtype=np.float32) for line in body.split("\n") if line.strip()]
    nxt = close_br + 1
    while nxt < end and buf[nxt:nxt + 1] in (b"\n", b" ", b"\r"):
        nxt += 1
    if not rows:
        return np.zeros((0, 0), dtype=np.float32), nxt
    if any(len(r) != len(rows[0]) for r in rows):
        raise KaldiFormatError("ragged text matrix at byte %d" % pos)
    return np.stack(rows), nxt


def _decode_object(buf, pos: int, end: int) -> Tuple[np.ndarray, int]:
    """Decode the matrix object starting at `pos`; -> (matrix, position of the byte after it)."""
    if pos + 2 <= end and buf[pos] == 0 and buf[pos + 1:pos + 2] == b"B":
        sp = buf.find(b" ", pos + 2, min(end, pos + 8))
        if sp < 0:
            raise KaldiFormatError("binary object without a type token at byte %d" % pos)
        token = bytes(buf[pos + 2:sp]).decode("ascii")
        pos = sp + 1
        if token in ("FM", "DM"):
            rows, pos = _sized_int(buf, pos)
            cols, pos = _sized_int(buf, pos)
            dt = "<f4" if token == "FM" else "<f8"
            n = rows * cols
            if rows < 0 or cols < 0 or pos + n * np.dtype(dt).itemsize > end:
                raise KaldiFormatError("matrix %dx%d runs past the end of the file" % (rows, cols))
            mat = np.frombuffer(buf, dtype=dt, count=n, offset=pos).reshape(rows, cols)
            return np.array(mat, dtype=mat.dtype.newbyteorder("=")), pos + n * mat.dtype.itemsize
        if token in ("FV", "DV"):
            dim, pos = _sized_int(buf, pos)
            dt = "<f4" if token == "FV" else "<f8"
            vec = np.frombuffer(buf, dtype=dt, count=dim, offset=pos).reshape(1, dim)
            return np.array(vec, dtype=vec.dtype.newbyteorder("=")), pos + dim * vec.dtype.itemsize
        if token in ("CM", "CM2", "CM3"):
            return _decode_cm(buf, pos, token)
        raise KaldiFormatError("unsupported Kaldi object type %r" % token)
    return _decode_text(buf, pos, end)


class ArkReader:
    """Random access into archives through scp-style rxfilenames; keeps up to `max_open` archives memory-mapped."""

    def __init__(self, max_open: int = 16):
        self.max_open = max_open
        self._lock = threading.Lock()          # loaders read from a background thread (BatchLoader read_ahead)
        self._maps: "OrderedDict[str, tuple]" = OrderedDict()          # path -> (file, mmap, (mtime, size, inode))

    def _map(self, path: str):
        with self._lock:
            return self._map_locked(path)

    def _map_locked(self, path: str):
        st = os.stat(path)
        stamp = (st.st_mtime_ns, st.st_size, st.st_ino)
        hit = self._maps.get(path)
        if hit is not None:
            if hit[2] == stamp:
                self._maps.move_to_end(path)
                return hit[1]
            hit[1].close()                                   # the file was rewritten since it was mapped
            hit[0].close()
            del self._maps[path]
        if st.st_size == 0:
            raise KaldiFormatError("empty file: %s" % path)
        f = open(path, "rb")
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        self._maps[path] = (f, mm, stamp)
        while len(self._maps) > self.max_open:
            _, (old_f, old_mm, _) = self._maps.popitem(last=False)
            old_mm.close()
            old_f.close()
        return mm

    def read_mat(self, rxfilename: str) -> np.ndarray:
        path, offset, rows, cols = _split_rxfilename(rxfilename)
        with self._lock:                       # mapping, possible eviction of another archive and the copy out of it
            mm = self._map_locked(path)
            if offset >= len(mm):
                raise KaldiFormatError("offset %d past the end of %s" % (offset, path))
            mat, _ = _decode_object(mm, offset, len(mm))
        return _apply_range(mat, rows, cols) if (rows or cols) else mat

    def iter_ark(self, path: str) -> Iterator[Tuple[str, np.ndarray]]:
        mm = self._map(path)
        pos, end = 0, len(mm)
        while pos < end:
            while pos < end and mm[pos:pos + 1] in (b"\n", b" ", b"\r"):
                pos += 1
            if pos >= end:
                break
            sp = mm.find(b" ", pos, end)
            if sp < 0:
                raise KaldiFormatError("archive key without an object at byte %d of %s" % (pos, path))
            key = bytes(mm[pos:sp]).decode("utf-8")
            mat, pos = _decode_object(mm, sp + 1, end)
            yield key, mat

    def close(self):
        for f, mm, _ in self._maps.values():
            mm.close()
            f.close()
        self._maps.clear()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


_default_reader: Optional[ArkReader] = None


def _reader() -> ArkReader:
    global _default_reader
    if _default_reader is None:
        _default_reader = ArkReader()
    return _default_reader


def read_mat(file_or_fd) -> np.ndarray:
    """`kaldi_io.read_mat`: an rxfilename ('feats.ark:1234', 'lda.mat') or a binary file object positioned at an object."""
    if isinstance(file_or_fd, (str, os.PathLike)):
        return _reader().read_mat(os.fspath(file_or_fd))
    start = file_or_fd.tell() if hasattr(file_or_fd, "tell") else None
    data = file_or_fd.read()
    mat, used = _decode_object(data, 0, len(data))
    if start is not None and hasattr(file_or_fd, "seek"):
        file_or_fd.seek(start + used)
    return mat


def read_scp(scp_file) -> "OrderedDict[str, str]":
    """key -> rxfilename, in file order (L/train.py:21-27 reads feats.scp the same way)."""
    table: "OrderedDict[str, str]" = OrderedDict()
    with open(scp_file, encoding="utf-8") as f:
        for n, line in enumerate(f, 1):
            parts = line.split(None, 1)
            if not parts:
                continue
            if len(parts) != 2 or not parts[1].strip():
                raise KaldiFormatError("%s line %d: expected 'key rxfilename'" % (scp_file, n))
            table[parts[0]] = parts[1].strip()
    return table


def read_mat_scp(scp_file) -> Iterator[Tuple[str, np.ndarray]]:
    """`kaldi_io.read_mat_scp`: (key, matrix) in scp order."""
    reader = _reader()
    for key, rx in read_scp(scp_file).items():
        yield key, reader.read_mat(rx)


def read_mat_ark(ark_file) -> Iterator[Tuple[str, np.ndarray]]:
    """`kaldi_io.read_mat_ark`: (key, matrix) in archive order."""
    return _reader().iter_ark(os.fspath(ark_file))


def _encode_object(mat: np.ndarray, binary: bool) -> bytes:
    mat = np.asarray(mat)
    if mat.ndim != 2:
        raise ValueError("write_mat expects a 2-D matrix, got shape %r" % (mat.shape,))
    if mat.dtype not in (np.float32, np.float64):
        mat = mat.astype(np.float32)
    if binary:
        token = b"FM " if mat.dtype == np.float32 else b"DM "
        head = b"\0B" + token + b"\4" + struct.pack("<i", mat.shape[0]) + b"\4" + struct.pack("<i", mat.shape[1])
        return head + np.ascontiguousarray(mat).astype(mat.dtype.newbyteorder("<"), copy=False).tobytes()
    if mat.shape[0] == 0:
        return b" [ ]\n"
    lines = ["  " + " ".join(repr(float(v)) for v in row) for row in mat]
    return (" [\n" + "\n".join(lines) + " ]\n").encode("ascii")


def write_mat(file_or_fd, mat, key: str = "", binary: bool = True) -> int:
    """Append one (key, matrix) entry; returns the byte offset of the object (what an scp line points at)."""
    own = isinstance(file_or_fd, (str, os.PathLike))
    f = open(file_or_fd, "ab") if own else file_or_fd
    try:
        if key:
            if any(c.isspace() for c in key):
                raise ValueError("Kaldi keys cannot contain white space: %r" % key)
            f.write(key.encode("utf-8") + b" ")
        offset = f.tell()
        f.write(_encode_object(mat, binary))
        return offset
    finally:
        if own:
            f.close()


def write_ark_scp(ark_file, scp_file, items, binary: bool = True) -> int:
    """Write (key, matrix) pairs to `ark_file` and the matching `scp_file` ('key ark:offset'); returns the entry count."""
    n = 0
    ark_path = os.path.abspath(os.fspath(ark_file))
    with open(ark_path, "wb") as ark, open(scp_file, "w", encoding="utf-8") as scp:
        for key, mat in items:
            offset = write_mat(ark, mat, key=key, binary=binary)
            scp.write("%s %s:%d\n" % (key, ark_path, offset))
            n += 1
    return n
