"""Writers for minimal KTX 1.1 and DDS files (test infrastructure): the tests synthesise their
own containers from the public file-format layouts, so what a loader must return is known by
construction (and is additionally compared with the compiled reference when it is present)."""
import struct

KTX_MAGIC = bytes([0xAB, 0x4B, 0x54, 0x58, 0x20, 0x31, 0x31, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A])


def level_dims(width, height, block, levels):
    out = []
    for _ in range(levels):
        out.append((width, height, (width + block - 1) // block, (height + block - 1) // block))
        width >>= 1
        height >>= 1
    return out


def write_ktx(path, gl_internal, gl_format, gl_type, width, height, unit_bytes, block, payloads,
              metadata=b"", big_endian=False, declared_levels=None, bad_size_level=None):
    e = ">" if big_endian else "<"
    n = len(payloads) if declared_levels is None else declared_levels
    hdr = KTX_MAGIC + struct.pack(e + "13I", 0x04030201, gl_type, 1 if gl_type else 0, gl_format, gl_internal,
                                  gl_format, width, height, 0, 0, 1, n, len(metadata))
    body = bytearray(hdr + metadata)
    for i, p in enumerate(payloads):
        size = len(p) + (4 if bad_size_level == i else 0)
        body += struct.pack(e + "I", size) + p
        body += b"\0" * (3 - ((len(p) + 3) % 4))
    open(path, "wb").write(bytes(body))


def write_dds(path, width, height, payloads, four_cc=b"\0\0\0\0", dx10_format=None, pf_flags=0x4, bitcount=0,
              masks=(0, 0, 0, 0), mip_flag=True, resource_dimension=3):
    flags = 0x1007 | (0x20000 if mip_flag else 0)
    pf = struct.pack("<II4sIIIII", 32, pf_flags, four_cc, bitcount, *masks)
    hdr = struct.pack("<7I", 124, flags, height, width, 0, 0, len(payloads)) + b"\0" * 44 + pf + struct.pack("<5I", 0x1000, 0, 0, 0, 0)
    assert len(hdr) == 124
    body = bytearray(b"DDS " + hdr)
    if dx10_format is not None:
        body += struct.pack("<5I", dx10_format, resource_dimension, 0, 1, 0)
    for p in payloads:
        body += p
    open(path, "wb").write(bytes(body))
