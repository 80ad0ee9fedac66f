"""CPU restatement (numpy) of the reference's MPP wire format — TEST INFRASTRUCTURE ONLY, like the rest of oracle/.

Follows (polardbx-executor/src/main/java/com/alibaba/polardbx/executor/):
  mpp/execution/buffer/PagesSerde.java:57-115 (uncompressed path), PagesSerdeUtil.java:36-48 (writeRawPage / readRawPage),
  :50-67 (SerializedChunk frame: positionCount int, marker byte, uncompressedSize int, sizeInBytes int, payload),
  chunk/IntegerBlockEncoding.java:46-70, LongBlockEncoding.java:47-73, DoubleBlockEncoding.java:46-70
  (positionCount int, NULL bit stream, then only the non-NULL values), chunk/EncoderUtil.java:43-150 (first row of
  every 8 in the most significant bit).  airlift Slice integers are little-endian.
Parity unpinned at the byte level: the reference's only serde test is a round trip (TestFileSingleStreamSpiller.java:74-99)
and holds no golden bytes; this restatement is checked by its own round trip and against the GPU codec byte for byte."""
import struct

import numpy as np

_W = {0: 4, 1: 8, 2: 8}
_DT = {0: "<i4", 1: "<i8", 2: "<f8"}


def _block(data, nulls, t):
    m = len(data)
    nl = np.zeros(m, dtype=bool) if nulls is None else np.asarray(nulls, dtype=bool)
    bits = np.packbits(nl.astype(np.uint8), bitorder="big").tobytes()     # first row -> most significant bit
    vals = np.ascontiguousarray(np.asarray(data)[~nl]).astype(_DT[t]).tobytes()
    return struct.pack("<i", m) + bits + vals


def serialize(cols, types, page_rows):
    """cols: [(values, nulls|None)] -> bytes of consecutive framed pages."""
    n = len(cols[0][0]) if cols else 0
    out = bytearray()
    for r0 in range(0, n, page_rows):
        r1 = min(n, r0 + page_rows)
        raw = struct.pack("<i", len(cols))
        for (d, nl), t in zip(cols, types):
            raw += _block(d[r0:r1], None if nl is None else nl[r0:r1], t)
        out += struct.pack("<ibii", r1 - r0, 0, len(raw), len(raw)) + raw
    return bytes(out)


def deserialize(buf, types):
    """bytes -> [(values, nulls)] (values under a NULL flag are 0, like the reference's fresh arrays)."""
    parts = [([], []) for _ in types]
    pos = 0
    while pos < len(buf):
        m, marker, unc, sz = struct.unpack_from("<ibii", buf, pos)
        assert marker == 0 and unc == sz
        pos += 13
        end = pos + sz
        (nb,) = struct.unpack_from("<i", buf, pos)
        assert nb == len(types)
        pos += 4
        for c, t in enumerate(types):
            (pc,) = struct.unpack_from("<i", buf, pos)
            assert pc == m
            pos += 4
            nbytes = (m + 7) // 8
            nl = np.unpackbits(np.frombuffer(buf, dtype=np.uint8, count=nbytes, offset=pos), bitorder="big")[:m].astype(bool)
            pos += nbytes
            k = int((~nl).sum())
            vals = np.zeros(m, dtype=_DT[t])
            vals[~nl] = np.frombuffer(buf, dtype=_DT[t], count=k, offset=pos)
            pos += k * _W[t]
            parts[c][0].append(vals)
            parts[c][1].append(nl)
        assert pos == end
    return [(np.concatenate(v) if v else np.zeros(0, dtype=_DT[t]), np.concatenate(n) if n else np.zeros(0, dtype=bool))
            for (v, n), t in zip(parts, types)]
