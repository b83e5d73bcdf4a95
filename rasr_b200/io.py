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
