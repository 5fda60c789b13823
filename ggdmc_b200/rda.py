"""Reader for R's `.rda` / `.RData` files (gzip'd ``RDX3`` XDR serialisation).

The reference ships its fixtures and checkpoints as `.rda` files written by
``save()`` (reference: README.md:173-178, tests/testthat/Group1/data/*.rda).  R
is not available where this package is built or tested, so this module decodes
the format directly.  It is host-side data-format plumbing only: nothing here
is on the sampling hot path.

The decoder understands exactly the SEXP types those files contain
(SURVEY.md section 8c): pairlists, symbols, back references, character
vectors, logical / integer / real vectors, generic vectors (lists), S4 objects,
environments and the ALTREP classes ``compact_intseq``, ``compact_realseq``,
``wrap_*`` and ``deferred_string``.

R objects are mapped to Python like this:

* atomic vectors -> :class:`RVector` (a ``numpy.ndarray`` subclass with an
  ``attrs`` dict; ``dim`` is applied as a Fortran-order reshape);
* character vectors -> :class:`RList` of ``str`` (``None`` for ``NA``);
* lists -> :class:`RList` (``.names`` gives the names, ``obj["name"]`` works);
* S4 objects -> :class:`RS4` (slots are attributes: ``obj.slot("theta")`` or
  ``obj["theta"]``).
"""
from __future__ import annotations

import gzip
import bz2
import lzma
import struct
from typing import Any, Dict, List, Optional

import numpy as np

# SEXP type codes (R internals, Rinternals.h / serialize.c)
NILSXP, SYMSXP, LISTSXP, CLOSXP, ENVSXP, PROMSXP, LANGSXP = 0, 1, 2, 3, 4, 5, 6
CHARSXP, LGLSXP, INTSXP, REALSXP, CPLXSXP, STRSXP = 9, 10, 13, 14, 15, 16
VECSXP, EXPRSXP, BCODESXP, EXTPTRSXP, RAWSXP, S4SXP = 19, 20, 21, 22, 24, 25
ALTREP_SXP, ATTRLISTSXP, ATTRLANGSXP = 238, 239, 240
BASEENV_SXP, EMPTYENV_SXP = 241, 242
NAMESPACESXP, PACKAGESXP, PERSISTSXP = 249, 250, 247
GLOBALENV_SXP, UNBOUNDVALUE_SXP, MISSINGARG_SXP, BASENAMESPACE_SXP = 253, 252, 251, 248
NILVALUE_SXP, REFSXP = 254, 255

NA_INTEGER = -2147483648


class RVector(np.ndarray):
    """numpy array carrying R attributes (``names``, ``dim``, ``dimnames`` ...)."""

    def __new__(cls, arr, attrs=None):
        obj = np.asarray(arr).view(cls)
        obj.attrs = dict(attrs or {})
        return obj

    def __array_finalize__(self, obj):
        self.attrs = getattr(obj, "attrs", {})

    @property
    def names(self):
        n = self.attrs.get("names")
        return list(n) if n is not None else None


class RList(list):
    """R generic vector / character vector / pairlist with optional names."""

    def __init__(self, items=(), attrs=None):
        super().__init__(items)
        self.attrs = dict(attrs or {})

    @property
    def names(self):
        n = self.attrs.get("names")
        return list(n) if n is not None else None

    def __getitem__(self, key):
        if isinstance(key, str):
            names = self.names or []
            return super().__getitem__(names.index(key))
        return super().__getitem__(key)

    def get(self, key, default=None):
        names = self.names or []
        if key in names:
            return super().__getitem__(names.index(key))
        return default

    def items(self):
        return zip(self.names or [None] * len(self), self)


class RS4:
    """S4 object: slots live in the attribute pairlist, class in ``class``."""

    def __init__(self, attrs):
        self.attrs = dict(attrs or {})

    @property
    def rclass(self):
        c = self.attrs.get("class")
        return c[0] if c else None

    def slot(self, name):
        return self.attrs[name]

    def __getitem__(self, name):
        return self.attrs[name]

    def slots(self):
        return [k for k in self.attrs if k != "class"]

    def __repr__(self):
        return f"<RS4 {self.rclass} slots={self.slots()}>"


class REnv:
    def __init__(self):
        self.frame = {}
        self.attrs = {}


class _Reader:
    def __init__(self, data: bytes):
        self.b = data
        self.p = 0
        self.refs: List[Any] = []

    # -- primitives ------------------------------------------------------
    def i32(self) -> int:
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def raw(self, n: int) -> bytes:
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def length(self) -> int:
        n = self.i32()
        if n == -1:  # long vector: two more ints
            hi = self.i32()
            lo = self.i32()
            n = (hi << 32) + lo
        return n

    # -- items -----------------------------------------------------------
    def attributes(self) -> Dict[str, Any]:
        """Read an attribute pairlist into an ordered dict."""
        pl = self.item()
        out = {}
        if isinstance(pl, RList):
            for k, v in zip(pl.attrs.get("_tags", []), pl):
                out[k] = v
        return out

    def item(self) -> Any:
        flags = self.i32()
        t = flags & 0xFF
        is_obj = bool(flags & 0x100)
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)

        if t == NILVALUE_SXP or t == NILSXP:
            return None
        if t in (EMPTYENV_SXP, BASEENV_SXP, GLOBALENV_SXP, BASENAMESPACE_SXP):
            return REnv()
        if t in (UNBOUNDVALUE_SXP, MISSINGARG_SXP):
            return None
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == SYMSXP:
            name = self.item()  # CHARSXP
            self.refs.append(name)
            return name
        if t in (NAMESPACESXP, PACKAGESXP, PERSISTSXP):
            self.i32()  # 0
            n = self.i32()
            info = [self.item() for _ in range(n)]
            env = REnv()
            env.attrs["_info"] = info
            self.refs.append(env)
            return env
        if t == ENVSXP:
            env = REnv()
            self.refs.append(env)
            self.i32()  # locked
            self.item()  # enclos
            frame = self.item()
            self.item()  # hashtab
            attr = self.item()
            if isinstance(frame, RList):
                for k, v in zip(frame.attrs.get("_tags", []), frame):
                    env.frame[k] = v
            if isinstance(attr, RList):
                env.attrs.update(dict(zip(attr.attrs.get("_tags", []), attr)))
            return env
        if t in (LISTSXP, LANGSXP, CLOSXP, PROMSXP, ATTRLISTSXP, ATTRLANGSXP):
            # iterate the cdr chain instead of recursing
            vals, tags = [], []
            attrs0 = None
            while True:
                if has_attr:
                    a = self.attributes()
                    if attrs0 is None:
                        attrs0 = a
                tag = self.item() if has_tag else None
                car = self.item()
                vals.append(car)
                tags.append(tag)
                # peek next cdr header
                nflags = self.i32()
                nt = nflags & 0xFF
                if nt in (LISTSXP, LANGSXP, ATTRLISTSXP, ATTRLANGSXP):
                    has_attr = bool(nflags & 0x200)
                    has_tag = bool(nflags & 0x400)
                    continue
                if nt in (NILVALUE_SXP, NILSXP):
                    break
                # improper list tail: rewind and read it as an item
                self.p -= 4
                vals.append(self.item())
                tags.append(None)
                break
            out = RList(vals, attrs0)
            out.attrs["_tags"] = tags
            if any(tg is not None for tg in tags):
                out.attrs.setdefault("names", tags)
            return out
        if t == CHARSXP:
            n = self.i32()
            if n == -1:
                return None
            return self.raw(n).decode("utf-8", errors="replace")
        if t == ALTREP_SXP:
            info = self.item()
            state = self.item()
            attr = self.item()
            val = self._altrep(info, state)
            if isinstance(attr, RList):
                a = dict(zip(attr.attrs.get("_tags", []), attr))
                val = self._with_attrs(val, a, False)
            return val

        # vectors
        if t in (LGLSXP, INTSXP):
            n = self.length()
            arr = np.frombuffer(self.raw(4 * n), dtype=">i4").astype(np.int32)
            val: Any = arr
        elif t == REALSXP:
            n = self.length()
            val = np.frombuffer(self.raw(8 * n), dtype=">f8").astype(np.float64)
        elif t == CPLXSXP:
            n = self.length()
            val = np.frombuffer(self.raw(16 * n), dtype=">c16").astype(np.complex128)
        elif t == RAWSXP:
            n = self.length()
            val = np.frombuffer(self.raw(n), dtype=np.uint8).copy()
        elif t == STRSXP:
            n = self.length()
            val = RList([self.item() for _ in range(n)])
        elif t in (VECSXP, EXPRSXP):
            n = self.length()
            val = RList([self.item() for _ in range(n)])
        elif t == S4SXP:
            val = RS4({})
        elif t == EXTPTRSXP:
            self.refs.append(None)
            self.item()
            self.item()
            val = None
        else:
            raise ValueError(f"rda: unsupported SEXP type {t} at byte {self.p}")

        if has_attr:
            a = self.attributes()
            val = self._with_attrs(val, a, t == LGLSXP)
        elif t == LGLSXP:
            val = RVector(val != 0) if not np.any(val == NA_INTEGER) else RVector(val)
        elif isinstance(val, np.ndarray):
            val = RVector(val)
        return val

    @staticmethod
    def _with_attrs(val, a, is_lgl):
        if isinstance(val, RS4):
            val.attrs.update(a)
            return val
        if isinstance(val, RList):
            val.attrs.update(a)
            return val
        if isinstance(val, np.ndarray):
            if is_lgl and not np.any(val == NA_INTEGER):
                val = val != 0
            dim = a.get("dim")
            if dim is not None:
                val = np.asarray(val).reshape(tuple(int(d) for d in dim), order="F")
            return RVector(val, a)
        return val

    @staticmethod
    def _altrep(info, state):
        cls = info[0] if isinstance(info, RList) else None
        if cls == "compact_intseq":
            n, start, step = (int(state[0]), int(state[1]), int(state[2]))
            return np.arange(start, start + n * step, step, dtype=np.int32)[:n]
        if cls == "compact_realseq":
            n, start, step = (int(state[0]), float(state[1]), float(state[2]))
            return start + step * np.arange(n, dtype=np.float64)
        if cls == "deferred_string":
            # state = pairlist(arg, scipen): the vector converted lazily by as.character
            arg = state[0]
            def fmt(v):
                if isinstance(v, (np.floating, float)):
                    return repr(int(v)) if float(v).is_integer() else repr(float(v))
                return str(int(v))
            return RList([fmt(v) for v in np.asarray(arg).ravel()])
        if cls is not None and cls.startswith("wrap_"):
            # state = list(x, meta)
            return state[0]
        raise ValueError(f"rda: unsupported ALTREP class {cls!r}")


def _decompress(path: str) -> bytes:
    with open(path, "rb") as fh:
        head = fh.read(6)
    if head[:2] == b"\x1f\x8b":
        with gzip.open(path, "rb") as fh:
            return fh.read()
    if head[:3] == b"BZh":
        with bz2.open(path, "rb") as fh:
            return fh.read()
    if head[:6] == b"\xfd7zXZ\x00":
        with lzma.open(path, "rb") as fh:
            return fh.read()
    with open(path, "rb") as fh:
        return fh.read()


def read_rda(path: str) -> Dict[str, Any]:
    """Load every object ``save()`` wrote to ``path`` -> ``{name: object}``."""
    data = _decompress(path)
    if data[:5] not in (b"RDX3\n", b"RDX2\n"):
        raise ValueError(f"{path}: not an RDX2/RDX3 file (magic {data[:5]!r})")
    rd = _Reader(data)
    rd.p = 5
    fmt = rd.raw(2)
    if fmt != b"X\n":
        raise ValueError(f"{path}: only XDR serialisation is supported, got {fmt!r}")
    version = rd.i32()
    rd.i32()  # writer version
    rd.i32()  # min reader version
    if version == 3:
        n = rd.i32()
        rd.raw(n)  # native encoding
    top = rd.item()
    if not isinstance(top, RList):
        raise ValueError(f"{path}: top-level object is not a pairlist")
    return dict(zip(top.attrs.get("_tags", []), top))
