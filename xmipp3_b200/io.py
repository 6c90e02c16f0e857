"""Python-side readers/writers of the on-disk formats of the path (tests, tools, dataset synthesis):
Xmipp metadata (.xmd STAR), Spider images/stacks/volumes and MRC stacks/volumes.

The C++ host program has its own readers (csrc/host/metadata.cpp, image_io.cpp); these Python twins
exist so that tests can produce inputs for it and inspect its outputs independently.
"""
import numpy as np


# ------------------------------------------------------------------ metadata
def write_xmd(path, columns, block="noname"):
    """columns: ordered dict label -> sequence (strings or numbers), all of the same length."""
    labels = list(columns.keys())
    n = len(next(iter(columns.values()))) if labels else 0
    with open(path, "w") as f:
        f.write("# XMIPP_STAR_1 * \n# \ndata_%s\nloop_\n" % block)
        for l in labels:
            f.write(" _%s\n" % l)
        for i in range(n):
            row = []
            for l in labels:
                v = columns[l][i]
                row.append(v if isinstance(v, str) else ("%d" % v if isinstance(v, (int, np.integer)) else "%.10g" % v))
            f.write(" " + " ".join(row) + "\n")


def write_ctfparam(path, **values):
    """Non-loop metadata block as used by .ctfparam files (ctfModel indirection, data/ctf.cpp:388-395)."""
    with open(path, "w") as f:
        f.write("# XMIPP_STAR_1 * \n# \ndata_fullMicrograph\n")
        for k, v in values.items():
            f.write(" _%s %.10g\n" % (k, v))


def read_xmd(path):
    labels, rows = [], []
    in_loop = False
    for line in open(path):
        t = line.strip()
        if not t or t[0] in "#;":
            continue
        if t.startswith("data_"):
            continue
        if t == "loop_":
            in_loop = True
            continue
        if t[0] == "_" and not rows:
            parts = t.split()
            labels.append(parts[0][1:])
            if not in_loop:
                rows.append(parts[1:2])
            continue
        rows.append(t.split())
    if not in_loop:
        rows = [[r[0] if r else "" for r in rows]]
    return {l: [r[k] for r in rows] for k, l in enumerate(labels)}


# ------------------------------------------------------------------ Spider
def _spider_header(nx, ny, nz, istack=0, maxim=0, imgnum=0):
    lenbyt = nx * 4
    labrec = 1024 // lenbyt + (1 if 1024 % lenbyt else 0)
    labbyt = labrec * lenbyt
    h = np.zeros(labbyt // 4, dtype="<f4")
    h[0], h[1], h[2], h[4] = nz, ny, ny * nz + labrec, (3 if nz > 1 else 1)
    h[11], h[12], h[21], h[22] = nx, labrec, labbyt, lenbyt
    h[23], h[25], h[26] = istack, maxim, imgnum
    return h


def write_spider_stack(path, images):
    images = np.ascontiguousarray(images, dtype="<f4")
    n, ny, nx = images.shape
    with open(path, "wb") as f:
        f.write(_spider_header(nx, ny, 1, istack=2, maxim=n).tobytes())
        for k in range(n):
            f.write(_spider_header(nx, ny, 1, imgnum=k + 1).tobytes())
            f.write(images[k].tobytes())


def write_spider(path, data):
    data = np.ascontiguousarray(data, dtype="<f4")
    if data.ndim == 2:
        data = data[None]
    nz, ny, nx = data.shape
    with open(path, "wb") as f:
        f.write(_spider_header(nx, ny, nz).tobytes())
        f.write(data.tobytes())


def read_spider(path):
    raw = open(path, "rb").read()
    h = np.frombuffer(raw[:108], dtype="<f4")
    dt = "<f4"
    if not (1 <= h[1] < 1e5 and h[1] == np.floor(h[1])):
        h = np.frombuffer(raw[:108], dtype=">f4")
        dt = ">f4"
    nz, ny, nx, labbyt = int(abs(h[0])), int(h[1]), int(h[11]), int(h[21])
    istack, maxim = int(h[23]), int(h[25])
    if istack > 0:
        out = np.empty((maxim, ny, nx), dtype=np.float32)
        off = labbyt
        for k in range(maxim):
            off += labbyt
            out[k] = np.frombuffer(raw, dtype=dt, count=nx * ny, offset=off).reshape(ny, nx)
            off += nx * ny * 4
        return out
    return np.frombuffer(raw, dtype=dt, count=nx * ny * nz, offset=labbyt).reshape(nz, ny, nx).astype(np.float32)


# ------------------------------------------------------------------ MRC
def write_mrc(path, data):
    data = np.ascontiguousarray(data, dtype="<f4")
    if data.ndim == 2:
        data = data[None]
    nz, ny, nx = data.shape
    h = np.zeros(256, dtype="<i4")
    h[0:4] = (nx, ny, nz, 2)
    h[7:10] = (nx, ny, nz)
    h[10:13] = np.array([nx, ny, nz], dtype="<f4").view("<i4")
    h[13:16] = np.array([90, 90, 90], dtype="<f4").view("<i4")
    h[16:19] = (1, 2, 3)
    b = bytearray(h.tobytes())
    b[208:212] = b"MAP "
    b[212:214] = bytes([0x44, 0x44])
    with open(path, "wb") as f:
        f.write(bytes(b))
        f.write(data.tobytes())


def read_mrc(path):
    raw = open(path, "rb").read()
    h = np.frombuffer(raw[:1024], dtype="<i4")
    nx, ny, nz, mode = (int(v) for v in h[0:4])
    ext = int(h[23])
    dt = {0: "i1", 1: "<i2", 2: "<f4", 6: "<u2", 12: "<f2"}[mode]
    return np.frombuffer(raw, dtype=dt, count=nx * ny * nz, offset=1024 + ext).reshape(nz, ny, nx).astype(np.float32)


def read_volume(path):
    p = path.split(":")[0].lower()
    return read_mrc(path) if p.endswith((".mrc", ".mrcs", ".map")) else read_spider(path)
