"""Reader of the DSPHDMP1 container written by oracle/ref_harness.cpp (format: oracle/ref_dump.h)."""
import numpy as np

_DT = {0: np.uint8, 1: np.uint16, 2: np.uint32, 3: np.float32, 4: np.float64, 5: np.int32}


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    assert data[:8] == b"DSPHDMP1", path
    off = 8
    while off < len(data):
        name = data[off:off + 32].split(b"\0", 1)[0].decode(); off += 32
        dtype, ncomp = np.frombuffer(data, np.uint32, 2, off); off += 8
        count = int(np.frombuffer(data, np.uint64, 1, off)[0]); off += 8
        dt = np.dtype(_DT[int(dtype)])
        n = count * int(ncomp)
        arr = np.frombuffer(data, dt, n, off).copy(); off += n * dt.itemsize
        if ncomp > 1:
            arr = arr.reshape(count, int(ncomp))
        if count == 1 and ncomp == 1:
            arr = arr[0]
        elif count == 1:
            arr = arr[0]
        out[name] = arr
    return out
