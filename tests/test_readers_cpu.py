"""CPU: the pair-dataset reader (imp_release_b200/readers.py; reference components/readers.py:8-39) on files with the layout
of dump/dumper/base_dumper.py:78-111.  No HDF5 library exists in this image, so the files come from tests/h5_writer.py, an
independent writer of the same on-disk structures (parity with libhdf5-written files is unpinned, see the reader's docstring)."""
import os
import struct

import numpy as np
import pytest
import torch

from imp_release_b200 import readers
from tests import h5_writer


def _pairs(n, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        n1, n2 = int(rng.integers(900, 2000)), int(rng.integers(900, 2000))
        out.append({
            'K1': rng.normal(size=(3, 3)), 'K2': rng.normal(size=(3, 3)), 'R': rng.normal(size=(3, 3)), 'T': rng.normal(size=(3,)),
            'e': rng.normal(size=(3, 3)), 'f': rng.normal(size=(3, 3)),
            'desc1': rng.normal(size=(n1, 256)).astype(np.float32), 'desc2': rng.normal(size=(n2, 256)).astype(np.float32),
            'kpt1': rng.uniform(0, 1000, size=(n1, 3)).astype(np.float32), 'kpt2': rng.uniform(0, 1000, size=(n2, 3)).astype(np.float32),
            'img_path1': f'images/seq_{i // 7:03d}/frame_{i:06d}.jpg', 'img_path2': f'images/seq_{i // 7:03d}/frame_{i + 1:06d}.jpg'})
    return out


@pytest.fixture(scope='module')
def pair_file(tmp_path_factory):
    pairs = _pairs(41)
    path = str(tmp_path_factory.mktemp('h5') / 'yfcc_sp_2000.hdf5')
    h5_writer.write_pair_file(path, pairs)
    return path, pairs


def test_standard_reader_matches_the_written_pairs(pair_file, monkeypatch):
    path, pairs = pair_file
    monkeypatch.setenv('IMP_READER', 'h5lite')
    rd = readers.standard_reader({'rawdata_dir': '/nonexistent', 'dataset_dir': path, 'num_kpt': 1500, 'read_images': False})
    assert len(rd) == len(pairs)
    for idx in (0, 1, 9, 10, 17, 40):
        info, p = rd.run(idx), pairs[idx]
        for k in ('K1', 'K2', 'R', 'e', 'f'):
            assert info[k].dtype == np.float64 and np.array_equal(info[k], p[k])
        t = p['T'] / np.sqrt((p['T'] ** 2).sum())
        assert np.array_equal(info['t'], t) and np.array_equal(info['t_gt'], t) and np.array_equal(info['r_gt'], p['R'])
        assert np.array_equal(info['desc1'], p['desc1'][:1500]) and np.array_equal(info['desc2'], p['desc2'][:1500])
        assert np.array_equal(info['x1'], p['kpt1'][:1500]) and np.array_equal(info['x2'], p['kpt2'][:1500])
        assert info['img1_path'] == p['img_path1'] and info['img2_path'] == p['img_path2']
        assert info['index'] == idx
    with pytest.raises(KeyError):
        rd.dataset['K1']['41']
    rd.close()
    ds = readers.reader_set({'rawdata_dir': '', 'dataset_dir': path, 'num_kpt': 1000, 'read_images': False})
    from torch.utils.data import Dataset
    assert isinstance(ds, Dataset) and len(ds) == len(pairs) and np.array_equal(ds[3]['desc2'], pairs[3]['desc2'][:1000])


def test_big_group_uses_a_multi_level_btree(tmp_path):
    """600 datasets in one group: 75 symbol-table nodes under a two-level B-tree (leaf K = 4, internal K = 16)."""
    w = h5_writer.H5Writer()
    vals = {str(i): np.arange(i % 5 + 1, dtype=np.int64) * i for i in range(600)}
    g = w.group({k: w.dataset(v) for k, v in vals.items()})[0]
    path = str(tmp_path / 'big.h5')
    open(path, 'wb').write(w.finish({'g': g, 'scalar': w.dataset(np.float32(2.5)), 'empty': w.dataset(np.zeros((0, 3), np.float32))}))
    raw = open(path, 'rb').read()
    levels = {raw[i + 5] for i in range(0, len(raw) - 8, 8) if raw[i:i + 4] == b'TREE'}
    assert levels == {0, 1}
    f = readers.H5Lite(path)
    assert sorted(f.keys()) == ['empty', 'g', 'scalar'] and len(f['g']) == 600
    for k, v in vals.items():
        assert np.array_equal(f['g'][k][()], v)
    assert f['g/17'][()].dtype == np.int64 and float(f['scalar'][()]) == 2.5 and f['empty'][()].shape == (0, 3)
    assert np.array_equal(f['g']['599'][1:3], vals['599'][1:3])


def test_errors_are_loud(tmp_path):
    bad = tmp_path / 'not.h5'
    bad.write_bytes(b'\0' * 4096)
    with pytest.raises(readers.H5LiteError):
        readers.H5Lite(str(bad))
    v2 = tmp_path / 'v2.h5'
    v2.write_bytes(b'\x89HDF\r\n\x1a\n' + struct.pack('<BBBB', 2, 8, 8, 0) + b'\0' * 64)
    with pytest.raises(readers.H5LiteError, match='superblock version 2/3'):
        readers.H5Lite(str(v2))


def test_pair_batcher_layout(pair_file, monkeypatch):
    """Batches in the matcher's input layout (zero padded to num_kpt, true counts in n_keypoints0/1)."""
    path, pairs = pair_file
    monkeypatch.setenv('IMP_READER', 'h5lite')
    rd = readers.standard_reader({'rawdata_dir': '', 'dataset_dir': path, 'num_kpt': 1200, 'read_images': False})
    pb = readers.PairBatcher(rd, batch=16)
    assert len(pb) == 3
    b = pb.load(2)                       # the ragged last batch: pairs 32..40
    assert b['indices'] == list(range(32, 41)) and b['descriptors0'].shape == (9, 1200, 256)
    for r, idx in enumerate(b['indices']):
        p = pairs[idx]
        n0, n1 = min(1200, len(p['kpt1'])), min(1200, len(p['kpt2']))
        assert int(b['n_keypoints0'][r]) == n0 and int(b['n_keypoints1'][r]) == n1
        assert torch.equal(b['keypoints1'][r, :n1], torch.from_numpy(p['kpt2'][:n1, :2]))
        assert torch.equal(b['scores0'][r, :n0], torch.from_numpy(p['kpt1'][:n0, 2]))
        assert torch.equal(b['descriptors0'][r, :n0], torch.from_numpy(p['desc1'][:n0]))
        assert float(b['descriptors0'][r, n0:].abs().sum()) == 0.0 and float(b['keypoints1'][r, n1:].abs().sum()) == 0.0


def test_object_header_variants_layouts_and_superblock_1(tmp_path):
    """The structures a libhdf5-written file shows beyond the minimum: NIL / modification-time / fill-value messages around
    the essential ones, the last message in a continuation block, superblock version 1, compact and unfiltered chunked
    layouts (edge chunks), big-endian-free integer and float types."""
    rng = np.random.default_rng(5)
    w = h5_writer.H5Writer(rich=True, superblock=1)
    a = rng.normal(size=(37, 10)).astype(np.float32)
    b = rng.integers(-1000, 1000, size=(5, 7)).astype(np.int32)
    c = rng.normal(size=(3, 3))
    d = rng.normal(size=(19, 6, 4)).astype(np.float32)
    sub = w.group({'chunked2d': w.chunked_dataset(a, (16, 4)), 'compact': w.compact_dataset(b)})[0]
    root = {'sub': sub, 'contig': w.dataset(c), 'chunked3d': w.chunked_dataset(d, (8, 6, 3)),
            'names': w.string_dataset([b'alpha.jpg', b'', b'a/much/longer/path/with/directories/image_000123.png'])}
    path = str(tmp_path / 'rich.h5')
    open(path, 'wb').write(w.finish(root))
    f = readers.H5Lite(path)
    assert np.array_equal(f['sub']['chunked2d'][()], a) and f['sub/chunked2d'].shape == (37, 10)
    assert np.array_equal(f['sub']['compact'][()], b) and f['sub']['compact'][()].dtype == np.int32
    assert np.array_equal(np.asarray(f['contig']), c)
    assert np.array_equal(f['chunked3d'][()], d)
    assert [x.decode() for x in f['names'][()]] == ['alpha.jpg', '', 'a/much/longer/path/with/directories/image_000123.png']
    assert np.array_equal(f['sub']['chunked2d'][3:9], a[3:9])
