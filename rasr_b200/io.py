"""On-disk model formats of the reference that feed this engine.

Mixture text file ("PMS", doc/file_formats/mixture_file.rst:24-40; reader src/Mm/MixtureSet.cc:168-208,
Mixture::read src/Mm/Mixture.cc:90-108, Mean::read / DiagonalCovariance::read src/Mm/GaussDensity.cc:32-70):

    #Version: 2.0
    #CovarianceType: DiagonalCovariance
    dim nMixtures nDensities nMeans nCovariances
    per mixture:    nDensities  densityId logWeight  densityId logWeight ...     (version < 2.0: linear weights)
    per density:    meanId covarianceId
    per mean:       dim m1 m2 ...
    per covariance: dim v1 w1 v2 w2 ...                       (the reader stores v * w and resets the weights to 1)

The reference reads the file with `istream >> token`, i.e. any whitespace separates tokens and line breaks carry
no meaning; f32 fields are parsed as f32 (strtof), weights as f64.  Files ending in .gz are gzip streams."""
import ctypes
import gzip
import io
import struct

import numpy as np

_libc = ctypes.CDLL(None)
_libc.strtof.restype = ctypes.c_float
_libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]


def _f32(tok):
    """decimal -> f32 exactly as `istream >> float` (single rounding; float(tok) would round twice)"""
    return _libc.strtof(tok, None)


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def read_mixture_set(path):
    """-> dict with the keys of rasr_b200.mm.MixtureSet (dim, mix_offsets, mix_density, mix_log_weight, dens_mean,
    dens_cov, means, variances)."""
    with _open(path, "rb") as f:
        data = f.read()
    stream = io.BytesIO(data)
    version_line = stream.readline().decode("ascii", "replace")
    cov_line = stream.readline().decode("ascii", "replace")
    if not version_line.startswith("#Version:"):
        raise ValueError("%s: not a mixture text file (first line %r)" % (path, version_line[:40]))
    version = np.float32(_f32(version_line[10:].strip().encode() or b"0"))
    if version > np.float32(2.0):
        raise ValueError('%s: version "%g" not supported' % (path, version))
    covtype = cov_line[17:].strip()
    if covtype != "DiagonalCovariance":
        raise ValueError("%s: no correct Covariance Type set: %s" % (path, covtype))
    tok = stream.read().split()
    pos = 0

    def take(n):
        nonlocal pos
        if pos + n > len(tok):
            raise ValueError("%s: unexpected end of file" % path)
        out = tok[pos:pos + n]
        pos += n
        return out

    dim, n_mix, n_dns, n_mean, n_cov = (int(t) for t in take(5))
    offs, dens, logw = [0], [], []
    for _ in range(n_mix):
        n = int(take(1)[0])
        for _ in range(n):
            d, w = take(2)
            dens.append(int(d))
            w = float(w)
            if version < np.float32(2.0):  # Mixture::addDensity: linear weight
                w = float(np.log(w)) if w > 0 else -np.finfo(np.float64).max
            logw.append(w)
        offs.append(len(dens))
    dens_mean, dens_cov = [], []
    for _ in range(n_dns):
        m, c = take(2)
        dens_mean.append(int(m))
        dens_cov.append(int(c))
    means = np.zeros((n_mean, dim), np.float32)
    for i in range(n_mean):
        d = int(take(1)[0])
        if d != dim:
            raise ValueError("%s: mean %d has dimension %d, the set has %d" % (path, i, d, dim))
        means[i] = [_f32(t) for t in take(d)]
    variances = np.zeros((n_cov, dim), np.float32)
    for i in range(n_cov):
        d = int(take(1)[0])
        if d != dim:
            raise ValueError("%s: covariance %d has dimension %d, the set has %d" % (path, i, d, dim))
        vw = take(2 * d)
        for k in range(d):  # v.push_back(velem * welem): f32 * f64 -> f64 -> f32
            variances[i, k] = np.float32(np.float64(np.float32(_f32(vw[2 * k]))) * float(vw[2 * k + 1]))
    return dict(dim=dim, mix_offsets=np.asarray(offs, np.uint32), mix_density=np.asarray(dens, np.uint32),
                mix_log_weight=np.asarray(logw, np.float64), dens_mean=np.asarray(dens_mean, np.uint32),
                dens_cov=np.asarray(dens_cov, np.uint32), means=means, variances=variances)


def write_mixture_set(path, ms):
    """Version 2.0 text file (MixtureSet::write src/Mm/MixtureSet.cc:141-167).  Numbers are printed with enough
    digits to read back bit-identically (the reference prints 6 significant digits)."""
    g32 = lambda v: "%.9g" % float(v)
    g64 = lambda v: "%.17g" % float(v)
    n_mix = len(ms["mix_offsets"]) - 1
    out = ["#Version: 2.0", "#CovarianceType: DiagonalCovariance",
           "%d %d %d %d %d" % (ms["dim"], n_mix, len(ms["dens_mean"]), len(ms["means"]), len(ms["variances"]))]
    for m in range(n_mix):
        a, b = int(ms["mix_offsets"][m]), int(ms["mix_offsets"][m + 1])
        out.append(" ".join([str(b - a)] + ["%d %s" % (ms["mix_density"][e], g64(ms["mix_log_weight"][e]))
                                            for e in range(a, b)]))
    for m, c in zip(ms["dens_mean"], ms["dens_cov"]):
        out.append("%d %d" % (m, c))
    for row in np.asarray(ms["means"], np.float32).reshape(len(ms["means"]), -1):
        out.append(" ".join([str(row.size)] + [g32(v) for v in row]))
    for row in np.asarray(ms["variances"], np.float32).reshape(len(ms["variances"]), -1):
        out.append(" " + " ".join([str(row.size)] + ["%s 1" % g32(v) for v in row]))
    with _open(path, "wb") as f:
        f.write(("\n".join(out) + "\n").encode("ascii"))


# ----------------------------------------------------------------------------------------------------------------
# Mixture-set accumulator files -- what the reference reads for every mixture file name that does not end in .pms or
# .gz (the `.mix` files of a trained system): MixtureSetReader falls back to MixtureSetEstimatorReader
# (src/Mm/MixtureSetReader.hh:108-121), which loads the ACCUMULATORS of the last training iteration and estimates the
# mixture set from them.  (`MixtureSet::read(Core::BinaryInputStream&)` itself is a stub that raises "reading of
# binary mixture sets is not supported in this version", src/Mm/MixtureSet.cc:219-222.)
#
# Layout (little endian; AbstractMixtureSetEstimator::read, src/Mm/AbstractMixtureSetEstimator.cc:404-478):
#   char magic[8] = "MIXSET\0\0", u32 version, u32 dimension
#   u32 nMeans,       per mean:       u32 size, f64 sum[size],         weight      (VectorAccumulator.hh:80-95)
#   u32 nCovariances, per covariance: u32 size, f64 sumOfSquares[size], weight
#   u32 nDensities,   per density:    u32 meanIndex, u32 covarianceIndex            (GaussDensityEstimator.cc:46-56)
#   u32 nMixtures,    per mixture:    u32 nDensities, per entry u32 densityIndex, weight   (MixtureEstimator.cc:140-160)
#   weight = f64 for version > 0, u32 count for version 0
# Estimation (AbstractMixtureSetEstimator::estimate :305-338 with the parameter defaults of :25-58):
#   - a mixture without observations is an error; per mixture, densities whose weight is below
#     max(minimum-observation-weight = 5, total * minimum-relative-weight = 0) are dropped, except the heaviest one
#     (MixtureEstimator.cc:64-79);
#   - densities, means and covariances are renumbered in the order the mixtures refer to them (:804-817);
#   - mixture weights: log(weight), normalised with logExpNorm (Mixture.cc:63-74, Utilities.hh:43-51);
#   - mean = sum / weight (f64, narrowed to f32; GaussDensityEstimator.cc:148-157);
#   - variance = (sumOfSquares - SUM_means sum_m^2 / weight_m) / weight over ALL means that shared the covariance before
#     densities were dropped (:201-229), then clipped from below at minimum-variance (0: off).
# ----------------------------------------------------------------------------------------------------------------
def read_mixture_estimator(path, min_observation_weight=5.0, min_relative_weight=0.0, min_variance=0.0,
                           normalize_mixture_weights=True):
    """-> mixture-set dict (keys as read_mixture_set) estimated from an accumulator file"""
    with _open(path, "rb") as f:
        data = f.read()
    if data[:8].rstrip(b"\0") != b"MIXSET":
        raise ValueError('%s: mixture set estimator file with magic "%s" could not be read, "MIXSET" expected'
                         % (path, data[:8].rstrip(b"\0").decode("ascii", "replace")))
    pos = 8

    def take(fmt):
        nonlocal pos
        v = struct.unpack_from("<" + fmt, data, pos)
        pos += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    version, dim = take("II")
    weight_fmt = "d" if version > 0 else "I"

    def accumulators():
        out = []
        for _ in range(take("I")):
            n = take("I")
            sums = np.frombuffer(data, "<f8", n, pos).astype(np.float64)
            skip(8 * n)
            out.append((sums, float(take(weight_fmt))))
        return out

    def skip(n):
        nonlocal pos
        pos += n

    mean_acc = accumulators()
    cov_acc = accumulators()
    dens = [take("II") for _ in range(take("I"))]
    mixtures = []
    for _ in range(take("I")):
        entries = []
        for _ in range(take("I")):
            d = take("I")
            entries.append([d, float(take(weight_fmt))])
        mixtures.append(entries)
    if pos > len(data):
        raise ValueError("%s: error reading mixture set estimator (file too short)" % path)

    # means per covariance BEFORE densities are dropped (CovarianceToMeanSetMap over every density a mixture refers to)
    cov_means = {}
    for entries in mixtures:
        for d, _ in entries:
            cov_means.setdefault(dens[d][1], set()).add(dens[d][0])
    for m, entries in enumerate(mixtures):
        total = sum(w for _, w in entries)
        if total == 0:
            raise ValueError("%s: mixture %d has zero weight" % (path, m))
        min_w = max(min_observation_weight, total * min_relative_weight)
        heaviest = max(range(len(entries)), key=lambda i: (entries[i][1], -i))  # first maximum
        keep_id = id(entries[heaviest])
        mixtures[m] = [e for e in entries if e[1] >= min_w or id(e) == keep_id]
    # renumber in order of first reference
    d_map, m_map, c_map = {}, {}, {}
    for entries in mixtures:
        for d, _ in entries:
            m_map.setdefault(dens[d][0], len(m_map))
            c_map.setdefault(dens[d][1], len(c_map))
            d_map.setdefault(d, len(d_map))
    offs, mix_density, logw = [0], [], []
    wmin = -np.finfo(np.float64).max  # Core::Type<f64>::min
    for entries in mixtures:
        lw = np.array([np.log(w) if w > 0 else wmin for _, w in entries], np.float64)
        if normalize_mixture_weights and lw.size:
            k = int(np.argmax(lw))
            acc = 0.0
            for i, v in enumerate(lw):
                if i != k:
                    acc += np.exp(v - lw[k])
            lw = lw - (np.log1p(acc) + lw[k])
        mix_density += [d_map[d] for d, _ in entries]
        logw += list(lw)
        offs.append(len(mix_density))
    dens_mean = np.zeros(len(d_map), np.uint32)
    dens_cov = np.zeros(len(d_map), np.uint32)
    for d, i in d_map.items():
        dens_mean[i], dens_cov[i] = m_map[dens[d][0]], c_map[dens[d][1]]
    means = np.zeros((len(m_map), dim), np.float32)
    for old, i in m_map.items():
        sums, w = mean_acc[old]
        if w != 0:
            means[i] = (sums / w).astype(np.float32)
    variances = np.zeros((len(c_map), dim), np.float32)
    for old, i in c_map.items():
        sq, w = cov_acc[old]
        if w == 0:
            continue
        mean_sq = np.zeros(dim, np.float64)
        for mo in sorted(cov_means[old]):  # the reference walks a std::set ordered by address = creation order
            ms, mw = mean_acc[mo]
            if mw > 0:
                mean_sq += ms * ms / mw
        v = ((sq - mean_sq) / w).astype(np.float32)
        if min_variance != 0:
            v = np.maximum(v, np.float32(min_variance))
        variances[i] = v
    return dict(dim=int(dim), mix_offsets=np.asarray(offs, np.uint32), mix_density=np.asarray(mix_density, np.uint32),
                mix_log_weight=np.asarray(logw, np.float64), dens_mean=dens_mean, dens_cov=dens_cov, means=means,
                variances=variances)


def read_mixture_file(path, **estimator_parameters):
    """What Mm::Module_::readMixtureSet does with a file name (src/Mm/MixtureSetReader.cc:27-34, .hh:108-121): text
    reader for *.pms and *.gz, the accumulator reader + estimation for everything else."""
    ext = "." + str(path).rsplit(".", 1)[-1] if "." in str(path) else ""
    if ext in (".pms", ".gz"):
        return read_mixture_set(path)
    return read_mixture_estimator(path, **estimator_parameters)


# ----------------------------------------------------------------------------------------------------------------
# Math::Matrix / Math::Vector files: the Nn layer parameters, priors, LDA matrices, normalisation vectors.
#
# A file name may carry a format qualifier "bin:" or "xml:" (Core::FormatSet, src/Core/FormatSet.cc:21-49); without
# one the Math formats default to XML (registered with isDefault = true, src/Math/Module.cc:26-40).
#   bin  (Core::BinaryFormat, src/Core/FormatSet.hh:244-256; little endian, src/Core/BinaryStream.hh:65):
#        matrix: u32 nRows, u32 nColumns, then elem_ = vector<Vector<T>>: u32 nRows, per row u32 nColumns + data
#                (src/Math/Matrix.hh:560-575, src/Core/BinaryStream.hh:209-215, src/Math/Vector.hh:286-299)
#        vector: u32 size + data
#   xml  <matrix-f32 nRows=".." nColumns=".."> row after row, whitespace separated </matrix-f32>
#        (src/Core/MatrixParser.hh:76-112, writer src/Math/Matrix.hh:642-647), <vector-f32 size=".."> (size optional,
#        src/Core/VectorParser.hh:78-103); values are read with `istream >> T`, written in scientific notation.
# ----------------------------------------------------------------------------------------------------------------
import re
import struct

_TYPES = {"f32": np.float32, "f64": np.float64, "s32": np.int32, "u32": np.uint32}


def split_qualifier(filename, default="xml"):
    """("bin"|"xml"|..., path) of a FormatSet file name"""
    filename = str(filename)
    pos = filename.find(":")
    return (default, filename) if pos < 0 else (filename[:pos], filename[pos + 1:])


def _parse_values(text, dtype):
    toks = text.split()
    if dtype == np.float32:
        return np.array([_f32(t.encode()) for t in toks], np.float32)  # single rounding, like `istream >> float`
    if dtype == np.float64:
        return np.array([float(t) for t in toks], np.float64)
    return np.array([int(t) for t in toks], dtype)


def _xml_element(data, kind):
    m = re.search(rb"<(%s-([a-z0-9]+))\b([^>]*)>(.*?)</\1\s*>" % kind.encode(), data, re.S)
    if not m:
        raise ValueError("no <%s-*> element found" % kind)
    atts = dict((k.decode(), v.decode()) for k, v in re.findall(rb'(\w+)\s*=\s*"([^"]*)"', m.group(3)))
    tname = m.group(2).decode()
    if tname not in _TYPES:
        raise ValueError("unsupported element type %s-%s" % (kind, tname))
    return _TYPES[tname], atts, m.group(4).decode()


def read_matrix(filename, dtype=np.float32):
    """Math::Matrix<T> from a `bin:` or `xml:` (default) file -> array [nRows, nColumns]"""
    fmt, path = split_qualifier(filename)
    with _open(path, "rb") as f:
        data = f.read()
    if fmt == "bin":
        dt = np.dtype(dtype).newbyteorder("<")
        n_rows, n_cols, n_vec = struct.unpack_from("<III", data, 0)
        if n_vec != n_rows:
            raise ValueError("%s: %d row vectors for a %d x %d matrix" % (path, n_vec, n_rows, n_cols))
        out = np.empty((n_rows, n_cols), dtype)
        pos = 12
        for r in range(n_rows):
            (n,) = struct.unpack_from("<I", data, pos)
            if n != n_cols:
                raise ValueError("%s: row %d has %d elements, expected %d" % (path, r, n, n_cols))
            out[r] = np.frombuffer(data, dt, n, pos + 4)
            pos += 4 + n * dt.itemsize
        return out
    if fmt != "xml":
        raise ValueError("unknown matrix format '%s'" % fmt)
    etype, atts, text = _xml_element(data, "matrix")
    if "nRows" not in atts or "nColumns" not in atts:
        raise ValueError("%s: nRows / nColumns attribute not given" % path)  # MatrixParser.hh:83-91
    n_rows, n_cols = int(atts["nRows"]), int(atts["nColumns"])
    vals = _parse_values(text, etype)
    if vals.size != n_rows * n_cols:
        raise ValueError("%s: %d elements for a %d x %d matrix" % (path, vals.size, n_rows, n_cols))
    return vals.reshape(n_rows, n_cols).astype(dtype, copy=False)


def read_vector(filename, dtype=np.float32):
    fmt, path = split_qualifier(filename)
    with _open(path, "rb") as f:
        data = f.read()
    if fmt == "bin":
        (n,) = struct.unpack_from("<I", data, 0)
        return np.frombuffer(data, np.dtype(dtype).newbyteorder("<"), n, 4).astype(dtype)
    if fmt != "xml":
        raise ValueError("unknown vector format '%s'" % fmt)
    etype, atts, text = _xml_element(data, "vector")
    vals = _parse_values(text, etype)
    if "size" in atts and int(atts["size"]) != vals.size:
        raise ValueError("%s: vector dimension mismatch: %s given and %d read" % (path, atts["size"], vals.size))
    return vals.astype(dtype, copy=False)


def _type_name(a):
    for k, v in _TYPES.items():
        if a.dtype == v:
            return k
    raise ValueError("unsupported dtype %s" % a.dtype)


def _sci(a, precision):
    fmt = "%%.%de" % precision if a.dtype.kind == "f" else "%d"
    return " ".join(fmt % v for v in a)


def write_matrix(filename, m, precision=20):
    """the writer the Nn trainer uses (formats().write(filename, parameters, 20), src/Nn/LinearLayer.cc:284)"""
    fmt, path = split_qualifier(filename)
    m = np.ascontiguousarray(m)
    with _open(path, "wb") as f:
        if fmt == "bin":
            f.write(struct.pack("<III", m.shape[0], m.shape[1], m.shape[0]))
            for row in m:
                f.write(struct.pack("<I", m.shape[1]) + row.astype(m.dtype.newbyteorder("<")).tobytes())
        else:
            t = _type_name(m)
            f.write(('<?xml version="1.0" encoding="UTF-8"?>\n<matrix-%s nRows="%d" nColumns="%d">\n'
                     % (t, m.shape[0], m.shape[1])).encode())
            for row in m:
                f.write((_sci(row, precision) + " \n").encode())
            f.write(("</matrix-%s>\n" % t).encode())


def write_vector(filename, v, precision=20):
    fmt, path = split_qualifier(filename)
    v = np.ascontiguousarray(v)
    with _open(path, "wb") as f:
        if fmt == "bin":
            f.write(struct.pack("<I", v.size) + v.astype(v.dtype.newbyteorder("<")).tobytes())
        else:
            t = _type_name(v)
            f.write(('<?xml version="1.0" encoding="UTF-8"?>\n<vector-%s size="%d">\n%s \n</vector-%s>\n'
                     % (t, v.size, _sci(v, precision), t)).encode())
