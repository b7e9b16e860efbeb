"""Minimal HDF5 WRITER for the tests of imp_release_b200/readers.py -- TEST INFRASTRUCTURE (this image has no HDF5 library).

Writes the structures h5py's defaults use for the reference's pair files (dump/dumper/base_dumper.py:78-111), following the
"HDF5 File Format Specification Version 3.0": superblock version 0, old-style groups (symbol-table message -> v1 B-tree of
SNOD nodes + local heap; leaf K = 4, internal K = 16, so big groups get multi-level trees), version-1 object headers,
contiguous datasets of IEEE floats / signed integers, variable-length ASCII strings in a global heap collection.
Independent of the reader: it shares no parsing code with it.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16


def _pad8(b: bytes) -> bytes:
    return b + b'\0' * (-len(b) % 8)


def _msg(mtype: int, data: bytes) -> bytes:
    data = _pad8(data)
    return struct.pack('<HHB3x', mtype, len(data), 0) + data


def _dtype_msg(dt: np.dtype) -> bytes:
    if dt.kind == 'f':
        size = dt.itemsize
        exp_bits, mant_bits = {4: (8, 23), 8: (11, 52)}[size]
        head = struct.pack('<BBBBI', 0x11, 0x20, size * 8 - 1, 0, size)
        return head + struct.pack('<HHBBBBI', 0, size * 8, mant_bits, exp_bits, 0, mant_bits, (1 << (exp_bits - 1)) - 1)
    if dt.kind in 'iu':
        return struct.pack('<BBBBI', 0x10, 0x08 if dt.kind == 'i' else 0, 0, 0, dt.itemsize) + struct.pack('<HH', 0, dt.itemsize * 8)
    raise ValueError(dt)


VLEN_STR = struct.pack('<BBBBI', 0x19, 0x01, 0, 0, 16) + struct.pack('<BBBBI', 0x13, 0, 0, 0, 1)


class H5Writer:
    def __init__(self, rich: bool = False, superblock: int = 0):
        self.rich, self.superblock = rich, superblock
        self.buf = bytearray(96 if superblock == 0 else 100)            # superblock goes here at the end
        self.gheap = []                     # (index, bytes) of the single global heap collection
        self.gheap_addr = None
        self.gheap_fixups = []              # offsets of the 8-byte collection address inside vlen elements

    def _alloc(self, data: bytes) -> int:
        self.buf += b'\0' * (-len(self.buf) % 8)
        off = len(self.buf)
        self.buf += data
        return off

    def _object_header(self, msgs) -> int:
        if self.rich:
            # what libhdf5 really writes around the three essential messages: a NIL message, a modification time, a fill value,
            # and the LAST message moved into a continuation block elsewhere in the file
            extra = [_msg(0x0000, b'\0' * 8), _msg(0x0012, struct.pack('<B3xI', 1, 1700000000)),
                     _msg(0x0005, struct.pack('<BBBB', 2, 2, 2, 0))]
            tail = msgs[-1]
            cont_addr = self._alloc(tail)
            msgs = extra[:1] + msgs[:-1] + extra[1:] + [_msg(0x0010, struct.pack('<QQ', cont_addr, len(tail)))]
            body = b''.join(msgs)
            return self._alloc(struct.pack('<BBHII4x', 1, 0, len(msgs) + 1, 1, len(body)) + body)
        body = b''.join(msgs)
        return self._alloc(struct.pack('<BBHII4x', 1, 0, len(msgs), 1, len(body)) + body)

    def dataset(self, arr) -> int:
        """numpy array -> object header address (contiguous layout)."""
        arr = np.asarray(arr)
        arr = arr if arr.flags.c_contiguous else arr.copy()      # (np.ascontiguousarray would turn a scalar into shape (1,))
        dt = arr.dtype.newbyteorder('<')
        data_addr = self._alloc(arr.astype(dt).tobytes()) if arr.size else UNDEF
        space = struct.pack('<BBB5x', 1, arr.ndim, 0) + b''.join(struct.pack('<Q', s) for s in arr.shape)
        layout = struct.pack('<BBQQ', 3, 1, data_addr, arr.nbytes)
        return self._object_header([_msg(0x0001, space), _msg(0x0003, _dtype_msg(dt)), _msg(0x0008, layout)])

    def compact_dataset(self, arr) -> int:
        """small array stored inside the object header (layout class 0)."""
        arr = np.asarray(arr)
        dt = arr.dtype.newbyteorder('<')
        raw = arr.astype(dt).tobytes()
        space = struct.pack('<BBB5x', 1, arr.ndim, 0) + b''.join(struct.pack('<Q', s) for s in arr.shape)
        layout = struct.pack('<BBH', 3, 0, len(raw)) + raw
        return self._object_header([_msg(0x0001, space), _msg(0x0003, _dtype_msg(dt)), _msg(0x0008, layout)])

    def chunked_dataset(self, arr, chunk) -> int:
        """unfiltered chunked layout: one level-0 v1 B-tree node (type 1) over all chunks, edge chunks stored at full size."""
        arr = np.asarray(arr)
        dt = arr.dtype.newbyteorder('<')
        nd = arr.ndim
        grid = [range(0, s, c) for s, c in zip(arr.shape, chunk)]
        import itertools
        entries = []
        for offs in itertools.product(*grid):
            blk = np.zeros(chunk, dt)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, arr.shape))
            blk[tuple(slice(0, x.stop - x.start) for x in sl)] = arr[sl]
            entries.append((offs, self._alloc(blk.tobytes()), blk.nbytes))
        assert len(entries) <= 64
        body = b''
        for offs, addr, nbytes in entries:
            body += struct.pack('<II', nbytes, 0) + b''.join(struct.pack('<Q', o) for o in offs) + struct.pack('<Q', 0) + struct.pack('<Q', addr)
        body += struct.pack('<II', 0, 0) + b''.join(struct.pack('<Q', s) for s in arr.shape) + struct.pack('<Q', 0)   # final key
        btree = self._alloc(b'TREE' + struct.pack('<BBHQQ', 1, 0, len(entries), UNDEF, UNDEF) + body)
        space = struct.pack('<BBB5x', 1, nd, 0) + b''.join(struct.pack('<Q', s) for s in arr.shape)
        layout = struct.pack('<BBB', 3, 2, nd + 1) + struct.pack('<Q', btree) + b''.join(struct.pack('<I', c) for c in chunk) + struct.pack('<I', dt.itemsize)
        return self._object_header([_msg(0x0001, space), _msg(0x0003, _dtype_msg(dt)), _msg(0x0008, layout)])

    def string_dataset(self, strings) -> int:
        """list of bytes -> 1-D dataset of variable-length ASCII strings (h5py.string_dtype(encoding='ascii'))."""
        elems = b''
        elem_offsets = []
        for s in strings:
            idx = len(self.gheap) + 1
            self.gheap.append((idx, bytes(s)))
            elem_offsets.append(len(elems) + 4)
            elems += struct.pack('<IQI', len(s), 0, idx)
        data_addr = self._alloc(elems)
        self.gheap_fixups += [data_addr + o for o in elem_offsets]
        space = struct.pack('<BBB5x', 1, 1, 0) + struct.pack('<Q', len(strings))
        layout = struct.pack('<BBQQ', 3, 1, data_addr, len(elems))
        return self._object_header([_msg(0x0001, space), _msg(0x0003, VLEN_STR), _msg(0x0008, layout)])

    def group(self, entries: dict):
        """{name: object header address} -> (object header address, B-tree address, local heap address)."""
        names = sorted(entries, key=lambda n: n.encode())
        heap_data = bytearray(8)            # offset 0: the empty string
        name_off = {}
        for n in names:
            name_off[n] = len(heap_data)
            heap_data += _pad8(n.encode() + b'\0')
        heap_data_addr = self._alloc(bytes(heap_data))
        heap_addr = self._alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap_data), UNDEF, heap_data_addr))
        # symbol table nodes
        level = []                           # (address, heap offset of the largest name below)
        per = 2 * LEAF_K
        for i in range(0, max(len(names), 1), per):
            chunk = names[i:i + per]
            body = b''.join(struct.pack('<QQII16x', name_off[n], entries[n], 0, 0) for n in chunk)
            body += b'\0' * (40 * (per - len(chunk)))
            addr = self._alloc(b'SNOD' + struct.pack('<BBH', 1, 0, len(chunk)) + body)
            level.append((addr, name_off[chunk[-1]] if chunk else 0))
        # B-tree levels
        lvl = 0
        per = 2 * INTERNAL_K
        while True:
            nodes = []
            sib = [level[i:i + per] for i in range(0, len(level), per)]
            addrs = []
            for ch in sib:                   # reserve the node addresses first (sibling pointers)
                self.buf += b'\0' * (-len(self.buf) % 8)
                addrs.append(len(self.buf))
                self.buf += b'\0' * (24 + (2 * per + 1) * 8)
            for j, ch in enumerate(sib):
                body = struct.pack('<Q', 0 if j == 0 else sib[j - 1][-1][1])      # key 0: largest name of everything to the left
                for a, k in ch:
                    body += struct.pack('<QQ', a, k)
                body += b'\0' * ((2 * per + 1) * 8 - len(body))
                node = b'TREE' + struct.pack('<BBHQQ', 0, lvl, len(ch), addrs[j - 1] if j > 0 else UNDEF,
                                             addrs[j + 1] if j + 1 < len(sib) else UNDEF) + body
                self.buf[addrs[j]:addrs[j] + len(node)] = node
                nodes.append((addrs[j], ch[-1][1]))
            if len(nodes) == 1:
                btree = nodes[0][0]
                break
            level, lvl = nodes, lvl + 1
        ohdr = self._object_header([_msg(0x0011, struct.pack('<QQ', btree, heap_addr))])
        return ohdr, btree, heap_addr

    def finish(self, root_entries: dict) -> bytes:
        if self.gheap:
            body = b''
            for idx, s in self.gheap:
                body += struct.pack('<HHIQ', idx, 1, 0, len(s)) + _pad8(s)
            size = max(4096, 16 + len(body) + 16)
            free = size - 16 - len(body)
            body += struct.pack('<HHIQ', 0, 0, 0, free) + b'\0' * (free - 16)
            self.gheap_addr = self._alloc(b'GCOL' + struct.pack('<B3xQ', 1, size) + body)
            for off in self.gheap_fixups:
                self.buf[off:off + 8] = struct.pack('<Q', self.gheap_addr)
        ohdr, btree, heap = self.group(root_entries)
        sb = b'\x89HDF\r\n\x1a\n' + struct.pack('<BBBBBBBBHHI', self.superblock, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        if self.superblock == 1:
            sb += struct.pack('<HH', 32, 0)         # indexed-storage internal node K, reserved
        sb += struct.pack('<QQQQ', 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack('<QQII', 0, ohdr, 1, 0) + struct.pack('<QQ', btree, heap)
        assert len(sb) == (96 if self.superblock == 0 else 100)
        self.buf[0:len(sb)] = sb
        return bytes(self.buf)


def write_pair_file(path: str, pairs) -> None:
    """pairs: list of dicts with the fields of dump/dumper/base_dumper.py:87-111 (K1 K2 R T e f img_path1 img_path2 desc1 desc2
    kpt1 kpt2) -> file with one group per field and one dataset per pair index."""
    w = H5Writer()
    root = {}
    for field in ('K1', 'K2', 'R', 'T', 'e', 'f', 'desc1', 'desc2', 'kpt1', 'kpt2'):
        root[field] = w.group({str(i): w.dataset(np.asarray(p[field])) for i, p in enumerate(pairs)})[0]
    for field in ('img_path1', 'img_path2'):
        root[field] = w.group({str(i): w.string_dataset([p[field].encode('ascii')]) for i, p in enumerate(pairs)})[0]
    with open(path, 'wb') as f:
        f.write(w.finish(root))
