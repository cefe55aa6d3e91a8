"""Pure-numpy reader for the reference's JLD2 / classic-HDF5 test fixtures.

Test tooling only (used by tests/golden/make_golden.py inside the build container,
where /root/reference exists). Not product code.
"""
# Minimal pure-Python reader for the reference's fixtures (no h5py/Julia). Survey tooling, not product code.
#   JLD2 (*.jld2): 512-B user block + HDF5 superblock v2, OHDR v2, layout v4 compact/contiguous, committed datatypes.
#   Classic HDF5 (*.h5 written by HDF5.jl): superblock v0, v1 object headers, root links as link messages, layout v3.
import struct, numpy as np
U = struct.unpack_from

def _link(d):                      # link message (type 6) -> (name, address or None)
    fl = d[1]; p = 2; lt = 0
    if fl & 8: lt = d[p]; p += 1
    if fl & 4: p += 8
    if fl & 16: p += 1
    n = 1 << (fl & 3); ln = int.from_bytes(d[p:p+n], 'little'); p += n
    name = d[p:p+ln].decode(); p += ln
    return name, (U('<Q', d, p)[0] if lt == 0 else None)

class DT:                          # parsed HDF5 datatype
    def __init__(s, cls, size, **kw): s.cls, s.size = cls, size; s.__dict__.update(kw)

def _dtype(b, p):                  # datatype message body -> (DT, next offset); versions 1-3
    cls, ver = b[p] & 15, b[p] >> 4
    bits = b[p+1] | b[p+2] << 8 | b[p+3] << 16
    size = U('<I', b, p+4)[0]; p += 8
    if cls == 0: return DT(0, size, signed=bool(bits & 8)), p + 4          # integer
    if cls == 1: return DT(1, size), p + 12                                 # float
    if cls == 3: return DT(3, size), p                                      # fixed string
    if cls == 4: return DT(4, size), p + 4                                  # bitfield (Bool)
    if cls == 5: return DT(5, size), p + (bits & 0xff)                      # opaque
    if cls == 7: return DT(7, size), p                                      # object reference (8-byte rel. offset)
    if cls == 9: base, p = _dtype(b, p); return DT(9, size, base=base, string=(bits & 15) == 1), p   # vlen
    if cls == 6:                                                            # compound (v3: no padding, minimal offsets)
        mem = []; osz = 1 if size < 256 else 2 if size < 65536 else 4
        for _ in range(bits & 0xffff):
            e = b.index(b'\0', p); name = b[p:e].decode('utf8', 'replace'); p = e + 1
            off = int.from_bytes(b[p:p+osz], 'little'); p += osz
            t, p = _dtype(b, p); mem.append((name, off, t))
        return DT(6, size, members=mem), p
    raise NotImplementedError(cls)

class JLD2:
    def __init__(s, fn):
        s.b = open(fn, 'rb').read(); i = s.b.find(b'\x89HDF\r\n\x1a\n'); assert s.b[i+8] in (2, 3)
        s.base = U('<Q', s.b, i+12)[0]; s.root = U('<Q', s.b, i+36)[0]
    def _msgs(s, addr):
        b = s.b; p = addr + s.base; assert b[p:p+4] == b'OHDR'; fl = b[p+5]; p += 6
        if fl & 0x20: p += 16
        if fl & 0x10: p += 4
        n = 1 << (fl & 3); size = int.from_bytes(b[p:p+n], 'little'); p += n
        chunks = [(p, p + size)]; out = []
        while chunks:
            p, e = chunks.pop(0)
            while p + 4 <= e:
                t, sz, mf = b[p], U('<H', b, p+1)[0], b[p+3]; p += 4 + (2 if fl & 4 else 0)
                d = b[p:p+sz]; p += sz
                if t == 0x10: a, l = U('<QQ', d, 0); chunks.append((a + s.base + 4, a + s.base + l - 4))   # OCHK
                elif t: out.append((t, mf, d))
        return out
    def _obj(s, addr):
        o = {'links': {}}
        for t, mf, d in s._msgs(addr):
            if t == 6: k, a = _link(d); o['links'][k] = a
            elif t == 1:                                    # dataspace v2: ver, rank, flags, type(0 scalar,1 simple,2 null)
                o['dims'] = U('<%dQ' % d[1], d, 4); o['null'] = d[3] == 2
            elif t == 3:                                    # datatype, possibly shared (committed under _types/)
                o['dt'] = s._obj(U('<Q', d, 2)[0])['dt'] if mf & 2 else _dtype(d, 0)[0]
            elif t == 8:                                    # layout v3/v4: class 0 compact, 1 contiguous
                if d[1] == 0: o['data'] = d[4:4 + U('<H', d, 2)[0]]
                elif d[1] == 1: a, n = U('<QQ', d, 2); o['data'] = s.b[a + s.base:a + s.base + n]
        return o
    def _heap(s, ref):                                      # vlen -> bytes from global heap collection
        n, a, idx = struct.unpack('<IQI', ref)
        if a == 0 and idx == 0: return b'', 0
        p = a + s.base; assert s.b[p:p+4] == b'GCOL'; end = p + U('<Q', s.b, p+8)[0]; q = p + 16
        while q < end:
            i, _, _, sz = U('<HHIQ', s.b, q)
            if i == idx: return s.b[q+16:q+16+sz], n
            if i == 0: break
            q += 16 + (sz + 7) // 8 * 8
        return None, n
    def _dec(s, t, buf, depth=0):
        if t.cls == 1: return struct.unpack('<d' if t.size == 8 else '<f', buf[:t.size])[0]
        if t.cls in (0, 4): return int.from_bytes(buf[:t.size], 'little', signed=getattr(t, 'signed', False))
        if t.cls == 3: return buf[:t.size].decode('utf8', 'replace')
        if t.cls == 5: return bytes(buf[:t.size])
        if t.cls == 7: a = U('<Q', buf, 0)[0]; return None if a in (0, 2**64 - 1) else (s.read(a, depth + 1) if depth < 8 else ('ref', a))
        if t.cls == 6: return {n: s._dec(m, buf[o:o + m.size], depth) for n, o, m in t.members}
        if t.cls == 9:
            raw, n = s._heap(buf[:16])
            if raw is None: return None
            return raw[:n].decode('utf8', 'replace') if t.string else [s._dec(t.base, raw[i*t.base.size:], depth) for i in range(n)]
    def read(s, addr, depth=0):
        o = s._obj(addr)
        if 'dt' not in o or o.get('data') is None:
            return {k: s.read(a, depth + 1) for k, a in o['links'].items() if k != '_types'} or None
        t, dims = o['dt'], o.get('dims', ())
        if o.get('null'): return np.zeros(0)
        n = int(np.prod(dims)) if dims else 1
        if t.cls in (0, 1) and t.size == 8:
            a = np.frombuffer(o['data'], '<f8' if t.cls == 1 else '<i8', n)
            return a[0] if not dims else (a.reshape(dims).T if len(dims) > 1 else a)   # HDF5 dims = reversed Julia dims
        v = [s._dec(t, o['data'][i*t.size:(i+1)*t.size], depth) for i in range(n)]
        return v if dims else v[0]
    def keys(s): return [k for k in s._obj(s.root)['links'] if k != '_types']
    def __getitem__(s, k): return s.read(s._obj(s.root)['links'][k])

def read_h5_v0(fn):                # classic HDF5 with float64 contiguous datasets linked from the root header
    b = open(fn, 'rb').read(); assert b[:8] == b'\x89HDF\r\n\x1a\n' and b[8] == 0
    root = U('<Q', b, 0x38 + 8)[0]
    def msgs(a):
        n, hs = U('<H', b, a+2)[0], U('<I', b, a+8)[0]; blocks = [(a+16, a+16+hs)]; out = []
        while blocks and len(out) < n:
            p, e = blocks.pop(0)
            while p + 8 <= e and len(out) < n:
                t, sz = U('<HH', b, p); d = b[p+8:p+8+sz]; p += 8 + sz
                if t == 0x10: ca, cl = U('<QQ', d, 0); blocks.append((ca, ca + cl))
                out.append((t, d))
        return out
    res = {}
    for t, d in msgs(root):
        if t != 6: continue
        name, a = _link(d); dims = addr = size = None
        for t2, d2 in msgs(a):
            if t2 == 1: dims = U('<%dQ' % d2[1], d2, 8 if d2[0] == 1 else 4)
            elif t2 == 8 and d2[0] == 3 and d2[1] == 1: addr, size = U('<QQ', d2, 2)
        arr = np.frombuffer(b[addr:addr+size], '<f8').reshape(dims)
        res[name] = arr.T if len(dims) > 1 else arr
    return res
