"""Pair-dataset reader with the reference's interface (components/readers.py:8-39; file layout written by
dump/dumper/base_dumper.py:78-111) -- SURVEY.md 8(f) rank 3, the on-disk format in front of the matcher.

The reference opens the ``*.hdf5`` pair file with h5py.  This module brings its own reader, ``H5Lite``, for exactly the subset of
HDF5 that h5py's defaults produce for that file (superblock 0/1, old-style groups = symbol table + v1 B-tree + local heap,
version-1 object headers, contiguous / compact / unfiltered chunked datasets of fixed-point and IEEE float types, fixed and
variable-length strings through the global heap).  Datasets come back as zero-copy ``numpy.memmap`` views, which is what the
staging side wants: ``PairBatcher`` packs B pairs straight from the page cache into the pinned buffers of
``imp_release_b200.feeder.PairFeeder`` (one H2D copy per batch, overlapped with the matcher) instead of materialising
per-pair arrays and tensors.  If ``h5py`` is installed it is used instead (set IMP_READER=h5lite to force the built-in one).

PARITY NOTE: this image has no HDF5 library, so ``H5Lite`` could not be checked against files written by libhdf5 here; it
follows the published "HDF5 File Format Specification Version 3.0", and the tests read files produced by an independent
in-tree writer of the same structures (tests/h5_writer.py).  Unsupported features fail loudly (``H5LiteError``).
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List, Optional

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b'\x89HDF\r\n\x1a\n'


class H5LiteError(RuntimeError):
    pass


class _Dataset:
    """One dataset: shape, dtype and where the bytes are.  ``ds[()]`` / ``np.asarray(ds)`` like h5py."""

    def __init__(self, f: 'H5Lite', shape, dtype, layout, vlen_str=False):
        self.file, self.shape, self.dtype, self.layout, self.vlen_str = f, tuple(shape), dtype, layout, vlen_str

    def _raw(self) -> np.ndarray:
        kind = self.layout[0]
        n = int(np.prod(self.shape)) if self.shape else 1
        if kind == 'contiguous':
            addr = self.layout[1]
            if addr == UNDEF or n == 0:
                return np.zeros(self.shape, self.dtype)
            return np.frombuffer(self.file.buf, self.dtype, n, self.file.base + addr).reshape(self.shape)
        if kind == 'compact':
            return np.frombuffer(self.layout[1], self.dtype, n).reshape(self.shape)
        if kind == 'chunked':
            _, btree, chunk = self.layout
            out = np.zeros(self.shape, self.dtype)
            if btree != UNDEF:
                for offs, addr, nbytes in self.file._chunks(btree, len(chunk)):
                    blk = np.frombuffer(self.file.buf, self.dtype, nbytes // self.dtype.itemsize, self.file.base + addr).reshape(chunk)
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, self.shape))
                    out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
            return out
        raise H5LiteError(f'unsupported layout {kind}')

    def __getitem__(self, key):
        a = self._raw()
        if self.vlen_str:
            flat = [self.file._vlen_bytes(bytes(e)) for e in a.reshape(-1)]
            out = np.empty(len(flat), dtype=object)
            out[:] = flat
            a = out.reshape(self.shape)
        return a[key] if key != () else (a if a.shape else a[()])

    def __array__(self, dtype=None, copy=None):
        a = self[()]
        return np.asarray(a, dtype=dtype) if dtype is not None else np.asarray(a)

    def __len__(self):
        return self.shape[0]


class _Group:
    def __init__(self, f: 'H5Lite', btree: int, heap: int):
        self.file, self.btree, self.heap = f, btree, heap
        self._entries: Optional[Dict[str, int]] = None

    def _load(self) -> Dict[str, int]:
        if self._entries is None:
            self._entries = self.file._group_entries(self.btree, self.heap)
        return self._entries

    def keys(self):
        return self._load().keys()

    def __len__(self):
        return len(self._load())

    def __contains__(self, name):
        return name in self._load()

    def __getitem__(self, name: str):
        node = self
        for part in name.strip('/').split('/'):
            ent = node._load()
            if part not in ent:
                raise KeyError(f"Unable to open object (object '{part}' doesn't exist)")
            node = node.file._open_object(ent[part])
        return node


class H5Lite(_Group):
    """Read-only HDF5 subset reader (see the module docstring).  ``H5Lite(path)['K1']['0'][()]``."""

    def __init__(self, path: str, mode: str = 'r'):
        if mode != 'r':
            raise H5LiteError('H5Lite is read-only')
        self.path = path
        self.buf = np.memmap(path, dtype=np.uint8, mode='r')
        b = self.buf
        start = 0
        while True:                                   # the superblock may sit at 0, 512, 1024, ...
            if start + 8 > len(b):
                raise H5LiteError(f'{path}: not an HDF5 file')
            if bytes(b[start:start + 8]) == SIGNATURE:
                break
            start = 512 if start == 0 else start * 2
        ver = int(b[start + 8])
        if ver in (0, 1):
            so, sl = int(b[start + 13]), int(b[start + 14])
            if (so, sl) != (8, 8):
                raise H5LiteError('only 8-byte offsets / lengths are supported')
            p = start + 24 + (4 if ver == 1 else 0)
            self.base = self._u64(p) + 0
            root = p + 32                              # base, free-space, end-of-file, driver-info addresses
            _, ohdr, cache, _ = struct.unpack_from('<QQII', b, root)
            if cache == 1:
                btree, heap = struct.unpack_from('<QQ', b, root + 24)
            else:
                btree, heap = self._symbol_table_message(ohdr)
        elif ver in (2, 3):
            raise H5LiteError("superblock version 2/3 (libver='latest': new-style groups) is not supported; write the file with "
                              "h5py's default libver")
        else:
            raise H5LiteError(f'unknown superblock version {ver}')
        self._gcol: Dict[int, Dict[int, bytes]] = {}
        super().__init__(self, btree, heap)

    # h5py-like conveniences
    def close(self):
        self.buf = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------ low level
    def _u64(self, off: int) -> int:
        return struct.unpack_from('<Q', self.buf, off)[0]

    def _messages(self, addr: int):
        """(type, flags, bytes) of every message of the version-1 object header at `addr`, continuation blocks included."""
        b, p = self.buf, self.base + addr
        if bytes(b[p:p + 4]) == b'OHDR':
            raise H5LiteError('version-2 object headers are not supported')
        ver, _, nmsg, _, size = struct.unpack_from('<BBHII', b, p)
        if ver != 1:
            raise H5LiteError(f'object header version {ver}')
        blocks = [(p + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            q, left = blocks.pop(0)
            while left >= 8 and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from('<HHB', b, q)
                data = bytes(b[q + 8:q + 8 + msize])
                q += 8 + msize
                left -= 8 + msize
                if mtype == 0x0010:                    # continuation
                    caddr, clen = struct.unpack_from('<QQ', data)
                    blocks.append((self.base + caddr, clen))
                out.append((mtype, flags, data))
        return out

    def _symbol_table_message(self, ohdr: int):
        for mtype, _, data in self._messages(ohdr):
            if mtype == 0x0011:
                return struct.unpack_from('<QQ', data)
        raise H5LiteError('object is not an old-style group (no symbol table message)')

    def _heap_data(self, heap: int) -> int:
        p = self.base + heap
        if bytes(self.buf[p:p + 4]) != b'HEAP':
            raise H5LiteError('bad local heap signature')
        return self.base + self._u64(p + 24)

    def _group_entries(self, btree: int, heap: int) -> Dict[str, int]:
        b = self.buf
        names = self._heap_data(heap)
        out: Dict[str, int] = {}
        stack = [btree]
        while stack:
            p = self.base + stack.pop()
            sig = bytes(b[p:p + 4])
            if sig == b'TREE':
                ntype, level, used = struct.unpack_from('<BBH', b, p + 4)
                if ntype != 0:
                    raise H5LiteError('expected a group B-tree node')
                q = p + 24                              # after the two sibling addresses: key0, child0, key1, ...
                for i in range(used):
                    stack.append(self._u64(q + 8 + 16 * i))
            elif sig == b'SNOD':
                n = struct.unpack_from('<H', b, p + 6)[0]
                for i in range(n):
                    noff, ohdr = struct.unpack_from('<QQ', b, p + 8 + 40 * i)
                    s = names + noff
                    e = s
                    while b[e] != 0:
                        e += 1
                    out[bytes(b[s:e]).decode()] = ohdr
            else:
                raise H5LiteError(f'unexpected node signature {sig!r} in a group B-tree')
        return out

    def _chunks(self, btree: int, ndims: int):
        b = self.buf
        stack = [btree]
        while stack:
            p = self.base + stack.pop()
            if bytes(b[p:p + 4]) != b'TREE':
                raise H5LiteError('bad chunk B-tree node')
            ntype, level, used = struct.unpack_from('<BBH', b, p + 4)
            if ntype != 1:
                raise H5LiteError('expected a chunk B-tree node')
            ksize = 8 + 8 * (ndims + 1)
            q = p + 24
            for i in range(used):
                nbytes, fmask = struct.unpack_from('<II', b, q)
                offs = struct.unpack_from(f'<{ndims}Q', b, q + 8)
                child = self._u64(q + ksize)
                if level > 0:
                    stack.append(child)
                else:
                    if fmask != 0:
                        pass
                    yield offs, child, nbytes
                q += ksize + 8

    def _vlen_bytes(self, elem: bytes) -> bytes:
        length, addr, idx = struct.unpack('<IQI', elem)
        if addr == 0 and idx == 0:
            return b''
        col = self._gcol.get(addr)
        if col is None:
            b, p = self.buf, self.base + addr
            if bytes(b[p:p + 4]) != b'GCOL':
                raise H5LiteError('bad global heap collection signature')
            size = self._u64(p + 8)
            col, q = {}, p + 16
            while q + 16 <= p + size:
                oi, _, _, osz = struct.unpack_from('<HHIQ', b, q)
                if oi == 0:
                    break
                col[oi] = bytes(b[q + 16:q + 16 + osz])
                q += 16 + ((osz + 7) & ~7)
            self._gcol[addr] = col
        return col[idx][:length]

    @staticmethod
    def _dtype(data: bytes):
        """datatype message -> (numpy dtype, is variable-length string)"""
        cls, ver = data[0] & 15, data[0] >> 4
        bits0 = data[1]
        size = struct.unpack_from('<I', data, 4)[0]
        order = '>' if bits0 & 1 else '<'
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 8 else 'u'}{size}"), False
        if cls == 1:
            return np.dtype(f'{order}f{size}'), False
        if cls == 3:
            return np.dtype(f'S{size}'), False
        if cls == 9:
            if (bits0 & 15) != 1:
                raise H5LiteError('variable-length sequences are not supported (only strings)')
            return np.dtype('V16'), True
        raise H5LiteError(f'unsupported datatype class {cls}')

    def _open_object(self, ohdr: int):
        msgs = self._messages(ohdr)
        kinds = {t for t, _, _ in msgs}
        if 0x0011 in kinds:
            bt, hp = next(struct.unpack_from('<QQ', d) for t, _, d in msgs if t == 0x0011)
            return _Group(self, bt, hp)
        if 0x0008 not in kinds:
            raise H5LiteError('object is neither an old-style group nor a dataset')
        if 0x000B in kinds:
            raise H5LiteError('filtered (compressed) datasets are not supported')
        shape, dtype, vlen, layout = (), None, False, None
        for t, _, d in msgs:
            if t == 0x0001:
                v, rank, flags = d[0], d[1], d[2]
                off = 8 if v == 1 else 4
                shape = struct.unpack_from(f'<{rank}Q', d, off) if rank else ()
            elif t == 0x0003:
                dtype, vlen = self._dtype(d)
            elif t == 0x0008:
                v = d[0]
                if v == 3:
                    c = d[1]
                    if c == 0:
                        n = struct.unpack_from('<H', d, 2)[0]
                        layout = ('compact', d[4:4 + n])
                    elif c == 1:
                        layout = ('contiguous', struct.unpack_from('<Q', d, 2)[0])
                    elif c == 2:
                        nd = d[2]
                        bt = struct.unpack_from('<Q', d, 3)[0]
                        dims = struct.unpack_from(f'<{nd}I', d, 11)
                        layout = ('chunked', bt, tuple(dims[:-1]))
                    else:
                        raise H5LiteError(f'layout class {c}')
                elif v in (1, 2):
                    nd, c = d[1], d[2]
                    p = 8
                    addr = UNDEF
                    if c != 0:
                        addr = struct.unpack_from('<Q', d, p)[0]
                        p += 8
                    dims = struct.unpack_from(f'<{nd}I', d, p)
                    p += 4 * nd
                    if c == 1:
                        layout = ('contiguous', addr)
                    elif c == 2:
                        layout = ('chunked', addr, tuple(dims[:-1]))
                    else:
                        n = struct.unpack_from('<I', d, p)[0]
                        layout = ('compact', d[p + 4:p + 4 + n])
                else:
                    raise H5LiteError(f'data layout message version {v}')
        if dtype is None or layout is None:
            raise H5LiteError('dataset without datatype / layout message')
        return _Dataset(self, shape, dtype, layout, vlen)


def open_h5(path: str):
    """h5py.File(path, 'r') when h5py is installed (and IMP_READER != 'h5lite'), else the built-in reader."""
    if os.environ.get('IMP_READER', '') != 'h5lite':
        try:
            import h5py  # type: ignore
            return h5py.File(path, 'r')
        except ImportError:
            pass
    return H5Lite(path)


class standard_reader:
    """components/readers.py:8-39 -- same constructor config keys, ``run(index)`` result, ``close`` and ``len``."""

    def __init__(self, config):
        self.raw_dir = config['rawdata_dir']
        self.dataset = open_h5(config['dataset_dir'])
        self.num_kpt = config['num_kpt']
        self.read_images = config.get('read_images', True)     # B200-side switch: the matcher itself never looks at pixels

    def run(self, index):
        ds = self.dataset
        K1, K2 = np.asarray(ds['K1'][str(index)]), np.asarray(ds['K2'][str(index)])
        R = np.asarray(ds['R'][str(index)])
        t = np.asarray(ds['T'][str(index)])
        t = t / np.sqrt((t ** 2).sum())
        desc1, desc2 = ds['desc1'][str(index)][()][:self.num_kpt], ds['desc2'][str(index)][()][:self.num_kpt]
        x1, x2 = ds['kpt1'][str(index)][()][:self.num_kpt], ds['kpt2'][str(index)][()][:self.num_kpt]
        e, f = ds['e'][str(index)][()], ds['f'][str(index)][()]
        img1_path, img2_path = ds['img_path1'][str(index)][()][0].decode(), ds['img_path2'][str(index)][()][0].decode()
        img1 = img2 = None
        if self.read_images:
            import cv2
            img1, img2 = cv2.imread(os.path.join(self.raw_dir, img1_path)), cv2.imread(os.path.join(self.raw_dir, img2_path))
        return {'index': index, 'K1': K1, 'K2': K2, 'R': R, 't': t, 'x1': x1, 'x2': x2, 'desc1': desc1, 'desc2': desc2,
                'img1': img1, 'img2': img2, 'e': e, 'f': f, 'r_gt': R, 't_gt': t, 'img1_path': img1_path, 'img2_path': img2_path}

    def close(self):
        self.dataset.close()

    def __len__(self):
        return len(self.dataset['K1'])


class reader_set(standard_reader):
    """components/readers.py:41-75 -- the same records behind the ``torch.utils.data.Dataset`` protocol (``ds[index]``)."""

    def __getitem__(self, index):
        return self.run(index)


try:        # make it a real Dataset subclass when torch is importable (it always is in this package; kept lazy for tools)
    from torch.utils.data import Dataset as _Dataset_t

    class reader_set(reader_set, _Dataset_t):  # noqa: F811
        pass
except ImportError:  # pragma: no cover
    pass


class PairBatcher:
    """Packs consecutive pairs of a ``standard_reader`` into fixed-capacity batch arrays in the matcher's input layout
    (what ``imp_release_b200.feeder.PairFeeder`` stages): keypoints [B, N, 2], scores [B, N], descriptors [B, N, 256] for both
    images, zero-padded to ``num_kpt`` with the true counts in ``n_keypoints0/1`` (the model's ragged-batch keys).  Rows are
    copied straight out of the memory-mapped file into the (optionally pinned) destination."""

    def __init__(self, reader: standard_reader, batch: int, pinned: bool = False):
        import torch
        self.reader, self.batch, self.N = reader, batch, reader.num_kpt
        mk = lambda *s, dt=torch.float32: (torch.zeros(*s, dtype=dt).pin_memory() if pinned else torch.zeros(*s, dtype=dt))
        self.buf = {}
        for i in (0, 1):
            self.buf[f'keypoints{i}'] = mk(batch, self.N, 2)
            self.buf[f'scores{i}'] = mk(batch, self.N)
            self.buf[f'descriptors{i}'] = mk(batch, self.N, 256)
            self.buf[f'n_keypoints{i}'] = mk(batch, dt=torch.int32)

    def __len__(self):
        return (len(self.reader) + self.batch - 1) // self.batch

    def load(self, b: int) -> Dict[str, object]:
        """Batch b -> dict of CPU tensors (views of the internal buffers; valid until the next load) + 'indices'."""
        lo, hi = b * self.batch, min((b + 1) * self.batch, len(self.reader))
        ds = self.reader.dataset
        views = {k: v.numpy() for k, v in self.buf.items()}
        for r, idx in enumerate(range(lo, hi)):
            for i, (kn, dn) in enumerate((('kpt1', 'desc1'), ('kpt2', 'desc2'))):
                k = np.asarray(ds[kn][str(idx)])[:self.N]
                d = np.asarray(ds[dn][str(idx)])[:self.N]
                n = len(k)
                views[f'keypoints{i}'][r, :n] = k[:, :2]
                views[f'scores{i}'][r, :n] = k[:, 2] if k.shape[1] > 2 else 1.0
                views[f'descriptors{i}'][r, :n] = d
                for name in (f'keypoints{i}', f'scores{i}', f'descriptors{i}'):
                    views[name][r, n:] = 0
                views[f'n_keypoints{i}'][r] = n
        out = {k: v[:hi - lo] for k, v in self.buf.items()}
        out['indices'] = list(range(lo, hi))
        return out
