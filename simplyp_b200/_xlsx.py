"""Minimal stdlib reader for .xlsx workbooks (zip + SpreadsheetML).

The reference reads its parameter workbook and the observation workbooks with
``pandas.read_excel`` (reference ``inputs.py:44-76`` and ``inputs.py:119-147``),
which needs openpyxl/xlrd.  Neither is available where this package has to run,
and the workbook layout used by SimplyP is simple (values only, no formulas we
need to evaluate), so this module parses the handful of XML parts directly.

Only what the SimplyP input path needs is implemented:

* sheet lookup by name (``xl/workbook.xml`` + ``xl/_rels/workbook.xml.rels``),
* shared strings, inline strings, booleans, numbers, cached formula values,
* date detection from the cell style (builtin and custom number formats),
  Excel serial origin 1899-12-30,
* ``read_table`` — an emulation of ``pd.read_excel(..., index_col=0, usecols=...)``
  for the layouts SimplyP uses (header in the first row).
"""
from __future__ import annotations

import datetime as _dt
import re
import zipfile
import xml.etree.ElementTree as ET

import numpy as np
import pandas as pd

_NS = {
    "m": "http://schemas.openxmlformats.org/spreadsheetml/2006/main",
    "r": "http://schemas.openxmlformats.org/officeDocument/2006/relationships",
    "pr": "http://schemas.openxmlformats.org/package/2006/relationships",
}

# Builtin number-format ids that denote dates/times (ECMA-376 18.8.30).
_BUILTIN_DATE_FMTS = set(range(14, 23)) | set(range(27, 37)) | set(range(45, 48)) | set(range(50, 59))

_EXCEL_EPOCH = _dt.datetime(1899, 12, 30)


def _col_to_index(col: str) -> int:
    """'A' -> 0, 'Z' -> 25, 'AA' -> 26."""
    n = 0
    for ch in col:
        n = n * 26 + (ord(ch.upper()) - ord("A") + 1)
    return n - 1


def _split_ref(ref: str):
    m = re.match(r"([A-Za-z]+)(\d+)$", ref)
    if not m:
        raise ValueError("bad cell reference %r" % ref)
    return int(m.group(2)) - 1, _col_to_index(m.group(1))


def parse_usecols(usecols: str):
    """Expand an Excel-style column spec such as ``"B,E:H"`` into 0-based indices."""
    out = []
    for part in usecols.split(","):
        part = part.strip()
        if not part:
            continue
        if ":" in part:
            a, b = part.split(":")
            out.extend(range(_col_to_index(a.strip()), _col_to_index(b.strip()) + 1))
        else:
            out.append(_col_to_index(part))
    return sorted(set(out))


def _is_date_format(code: str) -> bool:
    # strip quoted literals, bracketed colour/locale sections and escaped chars
    code = re.sub(r'"[^"]*"', "", code)
    code = re.sub(r"\[[^\]]*\]", "", code)
    code = re.sub(r"\\.", "", code)
    code = re.sub(r"_.|\*.", "", code)
    if code.lower() in ("general", ""):
        return False
    return bool(re.search(r"[dmyhs]", code, flags=re.I))


class Workbook:
    """A read-only view of one .xlsx file."""

    def __init__(self, path):
        self.path = str(path)
        self._zip = zipfile.ZipFile(self.path)
        self._shared = self._read_shared_strings()
        self._date_styles = self._read_date_styles()
        self._sheets = self._read_sheet_index()

    # ------------------------------------------------------------------ parts
    def _read_shared_strings(self):
        try:
            root = ET.fromstring(self._zip.read("xl/sharedStrings.xml"))
        except KeyError:
            return []
        out = []
        for si in root.findall("m:si", _NS):
            # concatenates rich-text runs; skips phonetic runs (rPh)
            texts = []
            for node in si:
                tag = node.tag.rsplit("}", 1)[-1]
                if tag == "t":
                    texts.append(node.text or "")
                elif tag == "r":
                    for t in node.findall("m:t", _NS):
                        texts.append(t.text or "")
            out.append("".join(texts))
        return out

    def _read_date_styles(self):
        try:
            root = ET.fromstring(self._zip.read("xl/styles.xml"))
        except KeyError:
            return set()
        custom = {}
        nf = root.find("m:numFmts", _NS)
        if nf is not None:
            for f in nf.findall("m:numFmt", _NS):
                custom[int(f.get("numFmtId"))] = f.get("formatCode", "")
        date_styles = set()
        xfs = root.find("m:cellXfs", _NS)
        if xfs is not None:
            for i, xf in enumerate(xfs.findall("m:xf", _NS)):
                fid = int(xf.get("numFmtId", "0"))
                if fid in custom:
                    if _is_date_format(custom[fid]):
                        date_styles.add(i)
                elif fid in _BUILTIN_DATE_FMTS:
                    date_styles.add(i)
        return date_styles

    def _read_sheet_index(self):
        wb = ET.fromstring(self._zip.read("xl/workbook.xml"))
        rels = ET.fromstring(self._zip.read("xl/_rels/workbook.xml.rels"))
        targets = {}
        for rel in rels.findall("pr:Relationship", _NS):
            tgt = rel.get("Target")
            if tgt.startswith("/"):
                tgt = tgt[1:]
            elif not tgt.startswith("xl/"):
                tgt = "xl/" + tgt
            targets[rel.get("Id")] = tgt
        sheets = {}
        order = []
        for sh in wb.find("m:sheets", _NS).findall("m:sheet", _NS):
            rid = sh.get("{%s}id" % _NS["r"])
            sheets[sh.get("name")] = targets[rid]
            order.append(sh.get("name"))
        self.sheet_names = order
        return sheets

    # ------------------------------------------------------------------ cells
    def cells(self, sheet_name):
        """Return ``{(row, col): value}`` (0-based) for every non-empty cell."""
        sheet_name = str(sheet_name)
        if sheet_name not in self._sheets:
            raise KeyError("no sheet named %r in %s (have %s)" % (sheet_name, self.path, self.sheet_names))
        root = ET.fromstring(self._zip.read(self._sheets[sheet_name]))
        data = root.find("m:sheetData", _NS)
        out = {}
        if data is None:
            return out
        for row in data.findall("m:row", _NS):
            for c in row.findall("m:c", _NS):
                ref = c.get("r")
                if ref is None:
                    continue
                t = c.get("t", "n")
                v = c.find("m:v", _NS)
                if t == "inlineStr":
                    is_ = c.find("m:is", _NS)
                    if is_ is None:
                        continue
                    val = "".join(tn.text or "" for tn in is_.iter("{%s}t" % _NS["m"]))
                elif v is None or v.text is None:
                    continue
                elif t == "s":
                    val = self._shared[int(v.text)]
                elif t == "b":
                    val = bool(int(v.text))
                elif t in ("str", "e"):
                    val = v.text
                else:
                    num = float(v.text)
                    style = int(c.get("s", "0"))
                    if style in self._date_styles:
                        val = _EXCEL_EPOCH + _dt.timedelta(days=num)
                    elif num.is_integer() and "." not in v.text and "e" not in v.text.lower():
                        val = int(num)
                    else:
                        val = num
                out[_split_ref(ref)] = val
        return out

    # ------------------------------------------------------------------ tables
    def read_table(self, sheet_name, usecols=None, index_col=0, header_row=0):
        """Emulate ``pd.read_excel(path, sheet_name=..., index_col=0, usecols=...)``.

        The first selected column becomes the index, the header row gives the
        column names, fully empty data rows are dropped at the tail only
        (pandas keeps interior blank rows as all-NaN, which SimplyP's sheets do
        not have).
        """
        cells = self.cells(sheet_name)
        if not cells:
            return pd.DataFrame()
        max_row = max(r for r, _ in cells)
        max_col = max(c for _, c in cells)
        cols = parse_usecols(usecols) if isinstance(usecols, str) else list(range(max_col + 1))
        header = [cells.get((header_row, c)) for c in cols]
        names = []
        for j, h in enumerate(header):
            names.append(h if h is not None else "Unnamed: %d" % j)
        rows = []
        for r in range(header_row + 1, max_row + 1):
            vals = [cells.get((r, c), np.nan) for c in cols]
            rows.append(vals)
        # drop trailing rows that are empty in all selected columns
        def _empty(vs):
            return all((isinstance(x, float) and np.isnan(x)) for x in vs)
        while rows and _empty(rows[-1]):
            rows.pop()
        # pandas drops rows that are entirely blank in the selected columns
        rows = [vs for vs in rows if not _empty(vs)]
        df = pd.DataFrame(rows, columns=names)
        for col in df.columns:
            ser = df[col]
            if ser.dtype == object:
                try:
                    conv = pd.to_numeric(ser)
                except (ValueError, TypeError):
                    continue
                df[col] = conv
        if index_col is not None:
            df = df.set_index(df.columns[index_col])
        return df


def read_table(path, sheet_name, usecols=None, index_col=0):
    return Workbook(path).read_table(sheet_name, usecols=usecols, index_col=index_col)
