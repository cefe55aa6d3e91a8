"""Writers for the reference's output files (src/smc_main.jl:499-526), byte-compatible with what JLD2.jl / HDF5.jl emit:

  * `savepath` (JLD2): keys `cloud` (a `SMC.Cloud`: committed compound type tagged `julia_type = SMC.Cloud` with the nine
    fields of src/particle.jl:31-41, array fields stored as referenced datasets), `w`, `W` (n_parts x n_stages Float64) and,
    for intermediate checkpoints, `j` (Int64) -- src/smc_main.jl:499-507,521-525;
  * `particle_store_path` (HDF5): dataset `smcparams`, n_parts x n_para Float64 -- src/smc_main.jl:514-520.

The layout mirrors the reference's own fixture `test/reference/smc_cloud_fix=true_version=150.jld2` object by object (512-byte
user block, superblock v2, version-2 object headers with Jenkins lookup3 checksums, link-info / group-info / link messages,
`_types/00000001` = JLD2's DataType description, `_types/00000002` = the Cloud compound, one global heap collection holding
the type names, compact layout for scalars, contiguous layout for arrays stored in Julia's column-major order);
tests/test_file_formats.py reads the files back with tools/fixture_reader.py and compares their structure with the fixture.
Host-side code: no device work here.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
BASE = 512


def lookup3(data: bytes, init: int = 0) -> int:
    """Bob Jenkins' lookup3 `hashlittle` (the HDF5 metadata checksum)."""
    M = 0xFFFFFFFF

    def rot(x, k):
        return ((x << k) | (x >> (32 - k))) & M

    length = len(data)
    a = b = c = (0xDEADBEEF + length + init) & M
    p = 0
    while length > 12:
        a = (a + int.from_bytes(data[p:p + 4], "little")) & M
        b = (b + int.from_bytes(data[p + 4:p + 8], "little")) & M
        c = (c + int.from_bytes(data[p + 8:p + 12], "little")) & M
        a = (a - c) & M; a ^= rot(c, 4); c = (c + b) & M
        b = (b - a) & M; b ^= rot(a, 6); a = (a + c) & M
        c = (c - b) & M; c ^= rot(b, 8); b = (b + a) & M
        a = (a - c) & M; a ^= rot(c, 16); c = (c + b) & M
        b = (b - a) & M; b ^= rot(a, 19); a = (a + c) & M
        c = (c - b) & M; c ^= rot(b, 4); b = (b + a) & M
        p += 12
        length -= 12
    if length == 0:
        return c
    tail = data[p:p + length] + b"\0" * (12 - length)
    a = (a + int.from_bytes(tail[0:4], "little")) & M
    b = (b + int.from_bytes(tail[4:8], "little")) & M
    c = (c + int.from_bytes(tail[8:12], "little")) & M
    c ^= b; c = (c - rot(b, 14)) & M
    a ^= c; a = (a - rot(c, 11)) & M
    b ^= a; b = (b - rot(a, 25)) & M
    c ^= b; c = (c - rot(b, 16)) & M
    a ^= c; a = (a - rot(c, 4)) & M
    b ^= a; b = (b - rot(a, 14)) & M
    c ^= b; c = (c - rot(b, 24)) & M
    return c


# ---- datatype message bodies (version 3) -------------------------------------------------------------------------------
DT_F64 = bytes.fromhex("31203f000800000000004000340b0034ff030000")      # IEEE binary64, little endian
DT_I64 = bytes.fromhex("300800000800000000004000")                      # signed 64-bit integer
DT_REF = bytes.fromhex("3700000008000000")                              # object reference (8-byte file offset)


def dt_vlen_string():
    # class 9 (variable length), type = string, padding null-terminate, charset UTF-8; base type: 1-byte fixed string
    return bytes.fromhex("39110100") + struct.pack("<I", 16) + bytes.fromhex("30000000") + struct.pack("<I", 1) + bytes.fromhex("00000800")


def dt_vlen_of(base: bytes):
    return bytes.fromhex("39000000") + struct.pack("<I", 16) + base


def dt_compound(members, size):
    """members: [(name, offset, datatype bytes)]; version 3: offsets take the minimum number of bytes for `size`."""
    osz = 1 if size < 256 else 2 if size < 65536 else 4
    out = bytes([0x36]) + struct.pack("<H", len(members)) + b"\0" + struct.pack("<I", size)
    for name, off, dt in members:
        out += name.encode("utf8") + b"\0" + off.to_bytes(osz, "little") + dt
    return out


def shared_dt(addr):
    return bytes([3, 2]) + struct.pack("<Q", addr)      # shared message v3, type 2: committed datatype at `addr`


def msg(mtype, data, flags=0):
    return bytes([mtype]) + struct.pack("<H", len(data)) + bytes([flags]) + data


def ohdr(messages: bytes) -> bytes:
    """Version-2 object header, one chunk, no times / attribute phase change; chunk-size field as small as it fits."""
    n = len(messages)
    if n < 256:
        head = b"OHDR" + bytes([2, 0x00]) + bytes([n])
    else:
        head = b"OHDR" + bytes([2, 0x01]) + struct.pack("<H", n)
    body = head + messages
    return body + struct.pack("<I", lookup3(body))


FILL_MSG = msg(0x05, bytes([3, 9]))                                      # fill value v3: never written, undefined


def array_object(addr, arr, dt=DT_F64):
    """Dataset holding a Julia Array (Float64 vector / matrix): dataspace dims are the reverse of Julia's, the data follow
    the header in Julia's column-major order.  Returns the bytes to place at `addr`."""
    a = np.asarray(arr, dtype=np.float64)
    jdims = a.shape
    raw = np.asfortranarray(a).tobytes(order="F")
    hd = jdims[::-1]
    space = bytes([2, len(hd), 0, 1]) + b"".join(struct.pack("<Q", d) for d in hd)
    m = FILL_MSG + msg(0x01, space) + msg(0x03, dt, flags=1)
    if len(raw) < 8192:                                     # JLD2 stores small datasets inside the header (compact layout)
        return ohdr(m + msg(0x08, bytes([4, 0]) + struct.pack("<H", len(raw)) + raw))
    layout_len = 4 + 2 + 16
    hdr_len = 4 + 2 + (1 if len(m) + layout_len < 256 else 2) + len(m) + layout_len + 4
    m += msg(0x08, bytes([4, 1]) + struct.pack("<QQ", addr + hdr_len, len(raw)))
    h = ohdr(m)
    assert len(h) == hdr_len
    return h + raw


def scalar_object(payload: bytes, dt_body: bytes, dt_flags: int):
    space = bytes([2, 0, 0, 0])
    m = FILL_MSG + msg(0x01, space) + msg(0x03, dt_body, flags=dt_flags)
    m += msg(0x08, bytes([4, 0]) + struct.pack("<H", len(payload)) + payload)
    return ohdr(m)


def group_object(links):
    """Root / `_types` group: link info, group info, one link message per child, and the 16-byte NIL message JLD2 leaves."""
    m = msg(0x02, bytes([0, 0]) + struct.pack("<QQ", UNDEF, UNDEF)) + msg(0x0A, bytes([0, 0]))
    for name, addr in links:
        nb = name.encode("utf8")
        m += msg(0x06, bytes([1, 0x10, 1, len(nb)]) + nb + struct.pack("<Q", addr))
    m += msg(0x00, b"\0" * 16)
    return ohdr(m)


def gcol(objects, size=4096):
    """Global heap collection (version 1) holding byte strings; returns (bytes, {index: length})."""
    out = b"GCOL" + bytes([1, 0, 0, 0]) + struct.pack("<Q", size)
    for idx, data in enumerate(objects, start=1):
        out += struct.pack("<HHIQ", idx, 1, 0, len(data)) + data + b"\0" * (-len(data) % 8)
    free = size - len(out)
    out += struct.pack("<HHIQ", 0, 0, 0, free) + b"\0" * (free - 16)
    assert len(out) == size
    return out


CLOUD_FIELDS = ("particles", "tempering_schedule", "ESS", "stage_index", "n_Φ", "resamples", "c", "accept", "total_sampling_time")


def write_jld2(path, cloud, w=None, W=None, j=None, julia_version="1.5.0", array_keys=("w", "W")):
    """`jldopen(path, true, true, true, IOStream) do file; write(file, "cloud", cloud); write(file, "w", w); write(file, "W", W)
    [; write(file, "j", j)] end` (src/smc_main.jl:499-507,521-525).  `cloud` is a smc_jl_b200.cloud.Cloud."""
    blob = bytearray(48)                                   # superblock placeholder (offsets below are relative to BASE)

    def here():
        return len(blob)

    def align8():
        blob.extend(b"\0" * (-len(blob) % 8))

    def put(b):
        a = here()
        blob.extend(b)
        return a

    # ---- _types/00000001: how JLD2 stores a Julia DataType {name::String, parameters::Vector{Any}} --------------------------
    heap_addr = 200
    t1_addr = here()
    dt_datatype = dt_compound([("name", 0, dt_vlen_string()), ("parameters", 16, dt_vlen_of(DT_REF))], 32)

    def julia_type_attr(name_index, name_len):
        data = struct.pack("<IQI", name_len, heap_addr, name_index) + struct.pack("<IQI", 0, 0, 0)
        body = bytes([2, 1]) + struct.pack("<HHH", 11, 10, 4) + b"julia_type\0" + shared_dt(t1_addr) + bytes([2, 0, 0, 0]) + data
        return msg(0x0C, body)

    names = [b"Core.DataType", b"SMC.Cloud"]
    put(ohdr(msg(0x03, dt_datatype, flags=0x40) + julia_type_attr(1, len(names[0]))))
    align8()
    assert here() <= heap_addr
    blob.extend(b"\0" * (heap_addr - here()))
    put(gcol(names))
    align8()
    blob.extend(b"\0" * 16)                                # (the fixture leaves 16 bytes before the next committed type)
    # ---- _types/00000002: mutable struct Cloud (src/particle.jl:31-41) --------------------------------------------------------
    t2_addr = here()
    members = [("particles", 0, DT_REF), ("tempering_schedule", 8, DT_REF), ("ESS", 16, DT_REF), ("stage_index", 24, DT_I64),
               ("n_Φ", 32, DT_I64), ("resamples", 40, DT_I64), ("c", 48, DT_F64), ("accept", 56, DT_F64),
               ("total_sampling_time", 64, DT_F64)]
    put(ohdr(msg(0x03, dt_compound(members, 72), flags=0x40) + julia_type_attr(2, len(names[1]))))
    # ---- cloud: the 72-byte struct (compact layout) + its three arrays as referenced datasets ---------------------------------
    cloud_addr = here()
    cloud_len = len(scalar_object(b"\0" * 72, shared_dt(t2_addr), 3))
    a_part = cloud_addr + cloud_len
    part = array_object(a_part, np.asarray(cloud.particles))
    a_sched = a_part + len(part)
    sched = array_object(a_sched, np.asarray(cloud.tempering_schedule, dtype=np.float64).ravel())
    a_ess = a_sched + len(sched)
    ess = array_object(a_ess, np.asarray(cloud.ESS, dtype=np.float64).ravel())
    payload = struct.pack("<QQQqqqddd", a_part, a_sched, a_ess, int(cloud.stage_index), int(cloud.n_Φ), int(cloud.resamples),
                          float(cloud.c), float(cloud.accept), float(cloud.total_sampling_time))
    put(scalar_object(payload, shared_dt(t2_addr), 3))
    put(part); put(sched); put(ess)
    links = [("cloud", cloud_addr)]
    for key, arr in ((array_keys[0], w), (array_keys[1], W)):        # array_keys: the debug dump of check_nan_ess stores two vectors
        if arr is not None:
            a = here()
            put(array_object(a, np.asarray(arr)))
            links.append((key, a))
    if j is not None:
        links.append(("j", put(scalar_object(struct.pack("<q", int(j)), DT_I64, 1))))
    types_addr = put(group_object([("00000001", t1_addr), ("00000002", t2_addr)]))
    links.append(("_types", types_addr))
    root_addr = put(group_object(links))
    eof = BASE + len(blob)
    sb = b"\x89HDF\r\n\x1a\n" + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", BASE, UNDEF, eof, root_addr)
    sb += struct.pack("<I", lookup3(sb))
    blob[0:48] = sb
    user = ("Julia data file (HDF5), version 0.2.0\0 (Julia %s 64-bit LE)\0" % julia_version).encode()
    with open(path, "wb") as fh:
        fh.write(user + b"\0" * (BASE - len(user)))
        fh.write(bytes(blob))


# ---- reader for the same subset (loadpath / continue_intermediate, src/smc_main.jl:246,334-335) ---------------------------------
class _File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        i = self.b.find(b"\x89HDF\r\n\x1a\n")
        if i < 0 or self.b[i + 8] != 2:
            raise ValueError("%s: not a JLD2 / HDF5 (superblock v2) file" % path)
        self.base, _, _, self.root = struct.unpack_from("<QQQQ", self.b, i + 12)

    def messages(self, addr):
        b, p = self.b, addr + self.base
        if b[p:p + 4] != b"OHDR" or b[p + 4] != 2:
            raise ValueError("unsupported object header")
        fl = b[p + 5]
        p += 6
        if fl & 0x20:
            p += 16
        if fl & 0x10:
            p += 4
        n = 1 << (fl & 3)
        size = int.from_bytes(b[p:p + n], "little")
        p += n
        end, out = p + size, []
        while p + 4 <= end:
            t, sz, mf = b[p], struct.unpack_from("<H", b, p + 1)[0], b[p + 3]
            p += 4 + (2 if fl & 4 else 0)
            if t:
                out.append((t, mf, b[p:p + sz]))
            p += sz
        return out

    def links(self, addr):
        out = {}
        for t, _, d in self.messages(addr):
            if t == 6:
                n = d[3]
                out[d[4:4 + n].decode("utf8")] = struct.unpack_from("<Q", d, 4 + n)[0]
        return out

    def dataset(self, addr):
        dims, raw, dt, shared = (), None, None, None
        for t, mf, d in self.messages(addr):
            if t == 1:
                dims = struct.unpack_from("<%dQ" % d[1], d, 4)
            elif t == 3:
                if mf & 2:
                    shared = struct.unpack_from("<Q", d, 2)[0]
                else:
                    dt = bytes(d)
            elif t == 8:
                if d[1] == 0:
                    raw = d[4:4 + struct.unpack_from("<H", d, 2)[0]]
                else:
                    a, n = struct.unpack_from("<QQ", d, 2)
                    raw = self.b[a + self.base:a + self.base + n]
        return dims, raw, dt, shared

    def array(self, addr):
        dims, raw, dt, _ = self.dataset(addr)
        if dt is None or dt[0] & 15 not in (0, 1) or struct.unpack_from("<I", dt, 4)[0] != 8:
            raise ValueError("unsupported element type")
        a = np.frombuffer(raw, "<f8" if dt[0] & 15 == 1 else "<i8")
        if not dims:
            return a[0]
        return np.asfortranarray(a.reshape(dims).T) if len(dims) > 1 else a.copy()


def read_jld2(path):
    """`load(path)` for the keys this package (and the reference) writes: {"cloud": Cloud, "w": ..., "W": ..., "j": ...}."""
    from .cloud import Cloud
    f = _File(path)
    out = {}
    for key, addr in f.links(f.root).items():
        if key == "_types":
            continue
        if key == "cloud":
            _, raw, _, shared = f.dataset(addr)
            tmsg = [d for t, _, d in f.messages(shared) if t == 3][0]
            if len(raw) != 72 or not all(n.encode("utf8") + b"\0" in tmsg for n in CLOUD_FIELDS):
                raise ValueError("`cloud` is not an SMC.Cloud")
            pa, sa, ea, si, nphi, nres, c, acc, tst = struct.unpack("<QQQqqqddd", raw)
            out[key] = Cloud(f.array(pa), f.array(sa), f.array(ea), si, nphi, nres, c, acc, tst)
        else:
            out[key] = f.array(addr)
    return out


def read_h5_matrix(path, name):
    f = _File(path)
    return f.array(f.links(f.root)[name])


# ---- plain HDF5 file with one Float64 matrix (particle_store_path) --------------------------------------------------------------
def write_h5_matrix(path, name, arr):
    """`h5open(path, "w") do file; write(file, name, arr) end` (src/smc_main.jl:514-520: dataset `smcparams`, n_parts x n_para).
    A valid HDF5 file in the same modern encoding as above (superblock v2 at offset 0, version-2 object headers, root links as
    link messages; dims = reversed Julia dims, data in Julia's column-major order), readable by HDF5.jl's `h5read(path, name)` /
    any libhdf5 >= 1.8.  (libhdf5 itself would lay the same content out with a v0 superblock and a symbol-table root group, as in
    the reference's `smcsave*.h5`; readers do not care.)"""
    blob = bytearray(48)
    a = len(blob)
    blob.extend(array_object(a, np.asarray(arr, dtype=np.float64)))
    root_addr = len(blob)
    blob.extend(group_object([(name, a)]))
    sb = b"\x89HDF\r\n\x1a\n" + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", 0, UNDEF, len(blob), root_addr)
    sb += struct.pack("<I", lookup3(sb))
    blob[0:48] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(blob))
