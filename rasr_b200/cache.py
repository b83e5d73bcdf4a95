"""Feature caches of the reference: Core::Archive containers (file, directory and bundle archives)
holding Flow data streams.

File archive (src/Core/FileArchive.cc:26-80 layout comment, code :166-560; all integers little endian,
src/Core/BinaryStream.hh:65; strings are u32 length + bytes, src/Core/BinaryStream.cc:174-179):

    8 bytes  "SP_ARC1\\0"
    1 byte   != 0: a file info table exists
    entries: u32 0xaa55aa55 | string name | u32 size | u32 compressed size (0 = stored) | u32 checksum (always 0) |
             data | u32 0x55aa55aa                     (an entry with an empty name is a hole left by a removal)
    table:   u32 n, n x (string name, u64 position of the entry's size field, u32 size, u32 compressed) |
             u32 m, m x (u64 position, u32 size) holes | u64 position of the hole table | u64 position of the table
    Without a table the reader scans the entries (scanArchive, :363-413).

Compressed entries are gzip members: a 10 byte gzip header, the raw deflate stream, crc32 and size
(Core::Archive::writeFile, src/Core/Archive.cc:142-215; readFile skips optional gzip header fields, :82-131).

Flow cache (src/Flow/Cache.cc:46-127): the entry named after the segment holds chunks of
    string datatype name ("vector-f32") | u32 n packets | n x packet       (Datatype::writeGatheredData, Datatype.cc:43-53)
with a vector packet = u32 size | size x T | f64 start | f64 end (src/Flow/Vector.hh:88-106, Timestamp.cc:43-53);
the entry "<segment>.attribs" holds <flow-attributes><flow-attribute name= value=/>...</flow-attributes>
(src/Flow/Attributes.hh:67-70,132-138).
"""
import os
import re
import struct
import zlib

import numpy as np

HEADER = b"SP_ARC1\0"
START_TAG, END_TAG = 0xAA55AA55, 0x55AA55AA
_VECTOR_TYPES = {"vector-f32": np.float32, "vector-f64": np.float64, "vector-s32": np.int32, "vector-u32": np.uint32,
                 "vector-s16": np.int16, "vector-u16": np.uint16, "vector-s8": np.int8, "vector-u8": np.uint8}


class ArchiveError(RuntimeError):
    pass


def _gzip_member(data):
    """what Archive::writeFile stores: fixed gzip header, deflate stream of compress2 (zlib header and adler dropped),
    crc32, size"""
    c = zlib.compressobj(zlib.Z_DEFAULT_COMPRESSION, zlib.DEFLATED, -15)
    body = c.compress(data) + c.flush()
    return (bytes([0x1F, 0x8B, 0x08, 0, 0, 0, 0, 0, 0, 0x03]) + body +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data) & 0xFFFFFFFF))


def _gunzip_member(tmp, size):
    """Archive::readFile: skip the gzip header (with its optional fields) and inflate; the trailer is not checked"""
    if len(tmp) < 18 or tmp[0] != 0x1F or tmp[1] != 0x8B:
        raise ArchiveError("compressed entry without gzip header")
    base, flags = 10, tmp[3]
    if flags & 0x04:
        base += 2 + struct.unpack_from("<H", tmp, base)[0]
    if flags & 0x08:
        base = tmp.index(b"\0", base) + 1
    if flags & 0x10:
        base = tmp.index(b"\0", base) + 1
    if flags & 0x02:
        base += 2
    out = zlib.decompressobj(-15).decompress(tmp[base:])
    if len(out) != size:
        raise ArchiveError("entry inflates to %d bytes, %d expected" % (len(out), size))
    return out


class FileArchive:
    """Read / append access to one archive file.  mode "r": existing archive; "w": create or append (the file info
    table is rewritten on close, like FileArchive::~FileArchive)."""

    def __init__(self, path, mode="r", allow_overwrite=False):
        if mode not in ("r", "w"):
            raise ValueError("mode must be 'r' or 'w'")
        self.path, self.mode, self.allow_overwrite = str(path), mode, allow_overwrite
        self.files = {}   # name -> (position of the size field, size, compressed)
        self.holes = []   # (position, size)
        self._changed = False
        exists = os.path.exists(self.path) and os.path.getsize(self.path) > 0
        if not exists:
            if mode == "r":
                raise ArchiveError('Archive file "%s" does not exist.' % self.path)
            d = os.path.dirname(self.path)
            if d:
                os.makedirs(d, exist_ok=True)
            self._f = open(self.path, "w+b")
            self._f.write(HEADER + b"\0")
            self._end = 9
            self._changed = True
        else:
            self._f = open(self.path, "rb" if mode == "r" else "r+b")
            if self._f.read(8) != HEADER:
                raise ArchiveError('No file archive header detected in file "%s".' % self.path)
            self._read_table()

    # -- low level -----------------------------------------------------------------------------------------------
    def _u32(self):
        b = self._f.read(4)
        if len(b) < 4:
            raise EOFError
        return struct.unpack("<I", b)[0]

    def _u64(self):
        b = self._f.read(8)
        if len(b) < 8:
            raise EOFError
        return struct.unpack("<Q", b)[0]

    def _string(self):
        n = self._u32()
        b = self._f.read(n)
        if len(b) < n:
            raise EOFError
        return b.decode("utf-8", "surrogateescape")

    def _read_table(self):
        self._f.seek(8)
        if self._f.read(1) != b"\0":
            self._f.seek(-8, os.SEEK_END)
            pos = self._u64()
            self._end = pos
            self._f.seek(pos)
            try:
                for _ in range(self._u32()):
                    name = self._string()
                    p, size, comp = self._u64(), self._u32(), self._u32()
                    self.files.setdefault(name, (p, size, comp))
                for _ in range(self._u32()):
                    p, size = self._u64(), self._u32()
                    self.holes.append((p, size))
            except EOFError:
                raise ArchiveError('Failed to read file info table from archive "%s".' % self.path)
        else:
            self._scan()

    def _scan(self):
        """scanArchive: walk the entries; a missing start tag skips four bytes and tries again"""
        self._f.seek(9)
        self._end = 9
        while True:
            try:
                if self._u32() != START_TAG:
                    continue
                name = self._string()
                pos = self._f.tell()
                size, comp, _ = self._u32(), self._u32(), self._u32()
                self._f.seek(comp if (comp and name) else size, os.SEEK_CUR)
                tag = self._u32()
            except EOFError:
                break
            if name:
                self.files.setdefault(name, (pos, size, comp))
            else:
                self.holes.append((pos, size))
            if tag == END_TAG:
                self._end = self._f.tell()

    # -- public --------------------------------------------------------------------------------------------------
    def names(self):
        return list(self.files)

    def __contains__(self, name):
        return name in self.files

    def read(self, name):
        if name not in self.files:
            raise KeyError(name)
        pos, size, comp = self.files[name]
        self._f.seek(pos + 12)
        data = self._f.read(comp if comp else size)
        if len(data) != (comp if comp else size):
            raise ArchiveError('entry "%s" is truncated' % name)
        return _gunzip_member(data, size) if comp else data

    def write(self, name, data, compress=False):
        if self.mode != "w":
            raise ArchiveError("archive is opened read-only")
        if os.path.normpath(name) != name or name.startswith("/"):
            raise ArchiveError('Filename "%s" contains special character sequences' % name)
        if name in self.files:
            if not self.allow_overwrite:
                raise ArchiveError("Overwriting is not allowed. Change parameter 'allow-overwrite'.")
            self._remove(name)
        self._set_changed()
        stored = _gzip_member(data) if compress else data
        comp = len(stored) if compress else 0
        bname = name.encode("utf-8", "surrogateescape")
        self._f.seek(self._end)
        self._f.write(struct.pack("<II", START_TAG, len(bname)) + bname)
        pos = self._f.tell()
        self._f.write(struct.pack("<III", len(data), comp, 0) + stored + struct.pack("<I", END_TAG))
        self._end = self._f.tell()
        self.files[name] = (pos, len(data), comp)

    def _remove(self, name):
        """FileArchive::remove: the last entry shrinks the archive, any other becomes a hole"""
        pos, size, comp = self.files.pop(name)
        nbytes = comp if comp else size
        begin = pos - (4 + len(name.encode("utf-8", "surrogateescape")) + 4)
        self._set_changed()
        if pos + 12 + nbytes + 4 == self._end:
            self._end = begin
        else:
            hole = nbytes + len(name.encode("utf-8", "surrogateescape"))
            self._f.seek(begin + 4)
            self._f.write(struct.pack("<IIII", 0, hole, 0, 0))
            self.holes.append((begin + 8, hole))

    def _set_changed(self):
        if not self._changed:
            self._f.seek(8)
            self._f.write(b"\0")
            self._changed = True

    def close(self):
        if self._f is None:
            return
        if self._changed and self.mode == "w":
            self._f.seek(self._end)
            table = self._f.tell()
            self._f.write(struct.pack("<I", len(self.files)))
            for name, (pos, size, comp) in self.files.items():
                b = name.encode("utf-8", "surrogateescape")
                self._f.write(struct.pack("<I", len(b)) + b + struct.pack("<QII", pos, size, comp))
            holes = self._f.tell()
            self._f.write(struct.pack("<I", len(self.holes)))
            for pos, size in self.holes:
                self._f.write(struct.pack("<QI", pos, size))
            self._f.write(struct.pack("<QQ", holes, table))
            self._f.truncate()
            self._f.seek(8)
            self._f.write(b"\1")
        self._f.close()
        self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class DirectoryArchive:
    """src/Core/DirectoryArchive.cc: every entry is a plain file below the archive directory; a compressed entry is a
    gzip file, recognised by its magic bytes, whose original size is its last four bytes (probe, :49-72).  The
    reference does not scan the directory (:39-47): names() walks it here for convenience."""

    def __init__(self, path, mode="r"):
        self.path, self.mode = str(path), mode
        if not os.path.isdir(self.path):
            if mode != "w":
                raise ArchiveError('Directory "%s" does not exist' % self.path)
            os.makedirs(self.path, exist_ok=True)

    def _file(self, name):
        return os.path.join(self.path, name)

    def names(self):
        out = []
        for root, _, files in os.walk(self.path):
            out += [os.path.relpath(os.path.join(root, f), self.path) for f in files]
        return sorted(out)

    def __contains__(self, name):
        return bool(name) and os.path.isfile(self._file(name))

    def read(self, name):
        if name not in self:
            raise KeyError(name)
        with open(self._file(name), "rb") as f:
            data = f.read()
        if data[:2] == b"\x1f\x8b" and len(data) >= 18:
            return _gunzip_member(data, struct.unpack("<I", data[-4:])[0])
        return data

    def write(self, name, data, compress=False):
        if self.mode != "w":
            raise ArchiveError("archive is opened read-only")
        if not name:
            raise ArchiveError("empty entry name")
        os.makedirs(os.path.dirname(self._file(name)) or self.path, exist_ok=True)
        with open(self._file(name), "wb") as f:
            f.write(_gzip_member(data) if compress else data)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class BundleArchive:
    """src/Core/BundleArchive.cc: a text file (suffix .bundle) listing archives, one path per whitespace-separated
    token; read-only.  The entry -> archive map is built from the members' file lists (createIndex, :125-137; a later
    archive wins for a duplicate name) and cached next to the bundle as <bundle>.idx.gz in the reference's format
    (writeIndex / readIndex, :139-172): number of archives, their paths, number of entries, "name index" lines."""

    def __init__(self, path, write_index=True):
        import gzip
        self.path = str(path)
        try:
            with open(self.path) as f:
                self.archives = f.read().split()
        except OSError:
            raise ArchiveError("cannot read bundle archive '%s'" % self.path)
        self._open = {}
        self.map = {}
        idx = self.path + ".idx.gz"
        if not self._read_index(idx):
            self.map = {}
            for i, a in enumerate(self.archives):
                with open_archive(a) as ar:
                    for n in ar.names():
                        self.map[n] = i
            if write_index:
                try:
                    with gzip.open(idx, "wt") as f:
                        f.write("%d\n" % len(self.archives) + "".join(a + "\n" for a in self.archives))
                        f.write("%d\n" % len(self.map) + "".join("%s %d\n" % kv for kv in self.map.items()))
                except OSError:
                    pass

    def _read_index(self, idx):
        import gzip
        try:
            with gzip.open(idx, "rt") as f:
                toks = f.read().split()
        except (OSError, EOFError):
            return False
        try:
            n = int(toks[0])
            if n != len(self.archives) or toks[1:1 + n] != self.archives:
                return False
            m = int(toks[1 + n])
            rest = toks[2 + n:2 + n + 2 * m]
            self.map = dict((rest[2 * i], int(rest[2 * i + 1])) for i in range(m))
            return len(self.map) == m or m == len(rest) // 2
        except (ValueError, IndexError):
            return False

    def names(self):
        return list(self.map)

    def __contains__(self, name):
        return name in self.map

    def read(self, name):
        if name not in self.map:
            raise KeyError(name)
        i = self.map[name]
        if i not in self._open:
            self._open[i] = open_archive(self.archives[i])
        return self._open[i].read(name)

    def write(self, *a, **k):
        raise ArchiveError("bundle archives are read-only")

    def close(self):
        for a in self._open.values():
            a.close()
        self._open = {}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def open_archive(path, mode="r", **kw):
    """Archive::create / Archive::test (src/Core/Archive.cc:287-329): *.bundle -> bundle; an existing directory, or a
    missing path ending in '/' -> directory archive; an existing file with the archive header, or a missing path ->
    file archive."""
    path = str(path)
    if path.endswith(".bundle"):
        return BundleArchive(path)
    if os.path.isdir(path) or (not os.path.exists(path) and path.endswith("/")):
        return DirectoryArchive(path, mode)
    if os.path.isfile(path) and os.path.getsize(path) > 0:
        with open(path, "rb") as f:
            if f.read(8) != HEADER:
                raise ArchiveError("unknown type of archive (requested path: '%s')." % path)
    return FileArchive(path, mode, **kw)


# ---- Flow cache ------------------------------------------------------------------------------------------------------

def decode_stream(data):
    """Entry of a Flow cache -> (datatype name, [packet arrays], times [n, 2] f64).  Chunks of gathered packets follow
    each other until the entry ends (CacheReader::readData is called again whenever a chunk is used up)."""
    pos, packets, times, dname = 0, [], [], None
    while pos < len(data):
        (n,) = struct.unpack_from("<I", data, pos)
        name = data[pos + 4:pos + 4 + n].decode()
        pos += 4 + n
        if name not in _VECTOR_TYPES:
            raise ArchiveError("datatype '%s' is not a Flow vector" % name)
        if dname is not None and name != dname:
            raise ArchiveError("mixed datatypes in one stream: %s, %s" % (dname, name))
        dname = name
        dt = np.dtype(_VECTOR_TYPES[name]).newbyteorder("<")
        (count,) = struct.unpack_from("<I", data, pos)
        pos += 4
        for _ in range(count):
            (size,) = struct.unpack_from("<I", data, pos)
            packets.append(np.frombuffer(data, dt, size, pos + 4).astype(dt.newbyteorder("=")))
            pos += 4 + size * dt.itemsize
            times.append(struct.unpack_from("<dd", data, pos))
            pos += 16
    return dname, packets, np.asarray(times, np.float64).reshape(-1, 2)


def encode_stream(packets, times, datatype="vector-f32", gather=0xFFFFFFFF):
    """CacheWriter::putData / ~CacheWriter: a chunk is flushed once MORE than `gather` packets are collected"""
    dt = np.dtype(_VECTOR_TYPES[datatype]).newbyteorder("<")
    name = datatype.encode()
    out, chunk = [], []

    def flush():
        if chunk:
            out.append(struct.pack("<I", len(name)) + name + struct.pack("<I", len(chunk)) + b"".join(chunk))
            chunk.clear()

    for p, (t0, t1) in zip(packets, times):
        p = np.asarray(p)
        chunk.append(struct.pack("<I", p.size) + p.astype(dt).tobytes() + struct.pack("<dd", t0, t1))
        if len(chunk) > gather:
            flush()
    flush()
    return b"".join(out)


def read_features(archive, segment):
    """(features [T, D] (a list of arrays if the packets differ in size), times [T, 2], attributes dict)"""
    _, packets, times = decode_stream(archive.read(segment))
    feats = np.stack(packets) if packets and all(p.size == packets[0].size for p in packets) else packets
    atts = {}
    if segment + ".attribs" in archive:
        for a in re.findall(rb"<flow-attribute\b([^>]*)/>", archive.read(segment + ".attribs")):
            kv = dict((k.decode(), v.decode()) for k, v in re.findall(rb'(\w+)\s*=\s*"([^"]*)"', a))
            if "name" in kv and "value" in kv:
                atts[kv["name"]] = kv["value"]
    return feats, times, atts


def write_features(archive, segment, feats, times, attributes=None, datatype="vector-f32", gather=0xFFFFFFFF,
                   compress=False):
    if attributes is not None:
        xml = '<?xml version="1.0" encoding="UTF-8"?>\n<flow-attributes>\n' + "".join(
            '  <flow-attribute name="%s" value="%s"/>\n' % (k, v) for k, v in attributes.items()) + "</flow-attributes>\n"
        archive.write(segment + ".attribs", xml.encode(), compress)
    archive.write(segment, encode_stream(feats, times, datatype, gather), compress)
