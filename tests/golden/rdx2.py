"""Minimal reader for R's `.rda` files (bzip2/gzip/xz of the XDR `RDX2` serialisation, version 2).

Test infrastructure only: it is used by `tests/golden/make_fixtures.py` to pull the
bundled `example_sce` counts / copy-number columns out of
`/root/reference/data/example_sce.rda` WITHOUT R (there is no R in this image).
The format is R's own (src/main/serialize.c in the R sources); nothing here comes from
the clonealign repository.  Only the SEXP types met in the three bundled files are
handled; anything else raises.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct


class RObj:
    """A parsed SEXP: `.type` (int), `.value`, optional `.attr` (dict name -> RObj), `.tag`."""

    __slots__ = ("type", "value", "attr", "tag")

    def __init__(self, type_, value=None, attr=None, tag=None):
        self.type = type_
        self.value = value
        self.attr = attr or {}
        self.tag = tag

    def __repr__(self):
        v = self.value
        if isinstance(v, (list, tuple)) and len(v) > 6:
            v = f"<{len(v)} items>"
        return f"RObj(type={self.type}, value={v!r}, attr={list(self.attr)})"


# SEXP type codes
NILSXP, SYMSXP, LISTSXP, CLOSXP, ENVSXP, PROMSXP, LANGSXP = 0, 1, 2, 3, 4, 5, 6
SPECIALSXP, BUILTINSXP, CHARSXP, LGLSXP, INTSXP, REALSXP, CPLXSXP = 7, 8, 9, 10, 13, 14, 15
STRSXP, DOTSXP, VECSXP, EXPRSXP, BCODESXP, EXTPTRSXP, WEAKREFSXP, RAWSXP, S4SXP = (
    16, 17, 19, 20, 21, 22, 23, 24, 25)
REFSXP, NILVALUE_SXP, GLOBALENV_SXP, UNBOUNDVALUE_SXP, MISSINGARG_SXP = 255, 254, 253, 252, 251
BASENAMESPACE_SXP, NAMESPACESXP, PACKAGESXP, PERSISTSXP = 250, 249, 248, 247
EMPTYENV_SXP, BASEENV_SXP = 242, 241
ATTRLANGSXP, ATTRLISTSXP, BCREPDEF, BCREPREF = 240, 239, 244, 243


class _Reader:
    def __init__(self, data: bytes):
        self.d = data
        self.p = 0
        self.refs: list = []

    # -- primitives -----------------------------------------------------
    def int(self) -> int:
        v = struct.unpack_from(">i", self.d, self.p)[0]
        self.p += 4
        return v

    def length(self) -> int:
        n = self.int()
        if n == -1:
            hi, lo = self.int(), self.int()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def bytes(self, n: int) -> bytes:
        b = self.d[self.p:self.p + n]
        self.p += n
        return b

    # -- items ----------------------------------------------------------
    def attr_list(self) -> dict:
        """Read a pairlist item and flatten it to {tag: value}."""
        out = {}
        node = self.item()
        while node is not None and node.type == LISTSXP:
            car, cdr = node.value
            out[node.tag] = car
            node = cdr
        return out

    def string_vec(self):
        if self.int() != 0:
            raise ValueError("names in persistent strings are not supported")
        n = self.int()
        return [self.item() for _ in range(n)]

    def item(self):
        flags = self.int()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))

        if t == NILVALUE_SXP:
            return None
        if t in (GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP, UNBOUNDVALUE_SXP,
                 MISSINGARG_SXP, BASENAMESPACE_SXP):
            return RObj(t)
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.int()
            return self.refs[idx - 1]
        if t in (NAMESPACESXP, PACKAGESXP, PERSISTSXP):
            o = RObj(t, self.string_vec())
            self.refs.append(o)
            return o
        if t == SYMSXP:
            name = self.item()
            o = RObj(SYMSXP, name.value)
            self.refs.append(o)
            return o
        if t == ENVSXP:
            self.int()  # locked
            o = RObj(ENVSXP, {})
            self.refs.append(o)  # before the children (they may refer back)
            enclos = self.item()
            frame = self.item()
            hashtab = self.item()
            attrib = self.item()
            env = {}

            def walk(node):
                while node is not None and node.type == LISTSXP:
                    car, cdr = node.value
                    env[node.tag] = car
                    node = cdr

            walk(frame)
            if hashtab is not None and hashtab.type == VECSXP:
                for bucket in hashtab.value:
                    walk(bucket)
            o.value = env
            del enclos, attrib
            return o
        if t in (LISTSXP, LANGSXP, CLOSXP, PROMSXP, DOTSXP):
            # iterative over the cdr chain to avoid deep recursion
            head = None
            prev = None
            while True:
                attr = self.attr_list() if has_attr else {}
                tag = None
                if has_tag:
                    tg = self.item()
                    tag = tg.value if tg is not None else None
                car = self.item()
                node = RObj(t, [car, None], attr, tag)
                if head is None:
                    head = node
                else:
                    prev.value[1] = node
                prev = node
                # peek at the cdr
                save = self.p
                flags = self.int()
                t2 = flags & 0xFF
                if t2 in (LISTSXP, LANGSXP, CLOSXP, PROMSXP, DOTSXP):
                    t = t2
                    has_attr = bool(flags & (1 << 9))
                    has_tag = bool(flags & (1 << 10))
                    continue
                self.p = save
                prev.value[1] = self.item()
                return head
        if t in (EXTPTRSXP, WEAKREFSXP):
            o = RObj(t)
            self.refs.append(o)
            if t == EXTPTRSXP:
                self.item()
                self.item()
            if has_attr:
                o.attr = self.attr_list()
            return o
        if t in (SPECIALSXP, BUILTINSXP):
            n = self.int()
            return RObj(t, self.bytes(n).decode())
        if t == CHARSXP:
            n = self.int()
            o = RObj(CHARSXP, None if n == -1 else self.bytes(n).decode("utf-8", "replace"))
        elif t in (LGLSXP, INTSXP):
            n = self.length()
            o = RObj(t, list(struct.unpack_from(f">{n}i", self.d, self.p)))
            self.p += 4 * n
        elif t == REALSXP:
            n = self.length()
            o = RObj(t, list(struct.unpack_from(f">{n}d", self.d, self.p)))
            self.p += 8 * n
        elif t == CPLXSXP:
            n = self.length()
            o = RObj(t, list(struct.unpack_from(f">{2 * n}d", self.d, self.p)))
            self.p += 16 * n
        elif t == STRSXP:
            n = self.length()
            o = RObj(t, [self.item().value for _ in range(n)])
        elif t in (VECSXP, EXPRSXP):
            n = self.length()
            o = RObj(t, [self.item() for _ in range(n)])
        elif t == RAWSXP:
            n = self.length()
            o = RObj(t, self.bytes(n))
        elif t == S4SXP:
            o = RObj(t)
        elif t == BCODESXP:
            nreps = self.int()
            reps = [None] * nreps
            o = RObj(t, self.bc1(reps))
        else:
            raise ValueError(f"unsupported SEXP type {t} at byte {self.p}")
        if has_attr:
            o.attr = self.attr_list()
        return o

    # -- byte code (only needs to be skipped correctly) -------------------
    def bc1(self, reps):
        code = self.item()
        n = self.int()
        consts = []
        for _ in range(n):
            t = self.int()
            if t == BCODESXP:
                consts.append(self.bc1(reps))
            elif t in (LANGSXP, LISTSXP, BCREPDEF, BCREPREF, ATTRLANGSXP, ATTRLISTSXP):
                consts.append(self.bclang(t, reps))
            else:
                consts.append(self.item())
        return (code, consts)

    def bclang(self, t, reps):
        if t == BCREPREF:
            return reps[self.int()]
        if t in (BCREPDEF, LANGSXP, LISTSXP, ATTRLANGSXP, ATTRLISTSXP):
            pos = -1
            has_attr = False
            if t == BCREPDEF:
                pos = self.int()
                t = self.int()
            if t == ATTRLANGSXP:
                t, has_attr = LANGSXP, True
            elif t == ATTRLISTSXP:
                t, has_attr = LISTSXP, True
            node = RObj(t, [None, None])
            if pos >= 0:
                reps[pos] = node
            if has_attr:
                node.attr = self.attr_list()
            tg = self.item()
            node.tag = tg.value if tg is not None else None
            node.value[0] = self.bclang(self.int(), reps)
            node.value[1] = self.bclang(self.int(), reps)
            return node
        return self.item()


def read_rda(path: str) -> dict:
    """Return {object name: RObj} for an `.rda` written by R's save()."""
    raw = open(path, "rb").read()
    if raw[:3] == b"BZh":
        raw = bz2.decompress(raw)
    elif raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    elif raw[:6] == b"\xfd7zXZ\x00":
        raw = lzma.decompress(raw)
    if raw[:5] != b"RDX2\n":
        raise ValueError("not an RDX2 file")
    if raw[5:7] != b"X\n":
        raise ValueError("only the XDR binary flavour is supported")
    r = _Reader(raw)
    r.p = 7
    version, _writer, _minreader = r.int(), r.int(), r.int()
    if version != 2:
        raise ValueError(f"serialisation version {version} not supported")
    top = r.item()
    out = {}
    node = top
    while node is not None and node.type == LISTSXP:
        car, cdr = node.value
        out[node.tag] = car
        node = cdr
    return out
