"""Reader for the GEX named-array container written by oracle/gex.h.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): used by tests/ and by tests/golden/make_golden.py.
"""
import struct
import numpy as np


def read_gex(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    assert data[:4] == b"GEX1", "not a GEX file"
    o = 4
    while o < len(data):
        (nl,) = struct.unpack_from("<I", data, o); o += 4
        name = data[o:o + nl].decode(); o += nl
        (dl,) = struct.unpack_from("<I", data, o); o += 4
        dt = data[o:o + dl].decode(); o += dl
        (nd,) = struct.unpack_from("<I", data, o); o += 4
        shape = struct.unpack_from("<%dQ" % nd, data, o) if nd else (); o += 8 * nd
        (nb,) = struct.unpack_from("<Q", data, o); o += 8
        arr = np.frombuffer(data, dtype=np.dtype("<" + dt), count=nb // np.dtype(dt).itemsize, offset=o).copy()
        o += nb
        out[name] = arr.reshape(shape) if nd else arr.reshape(())
    return out
