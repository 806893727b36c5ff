"""CPU: on-disk formats after the hot path against goldens recorded from
the unmodified reference (tests/golden/make_formats_golden.py): the pickled reads index and
the per-read statistics block."""
import base64
import json
import os
import pickle

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _reads(th):
    rs = np.random.RandomState(5)
    out = []
    for i in range(7):
        chrm, strand = ('chr1', '+') if i % 3 else ('chr2', '-')
        out.append((chrm, strand, th.readData(
            start=1000 * i, end=1000 * i + 450, filtered=bool(i == 4),
            read_start_rel_to_raw=int(rs.randint(0, 300)), strand=strand,
            fn='/data/run1/sub%d/read_%d.fast5' % (i % 2, i),
            corr_group='RawGenomeCorrected_000/BaseCalled_template', rna=False,
            sig_match_score=float(rs.uniform(0.5, 1.2)), mean_q_score=float(rs.uniform(7, 12)),
            read_id='id-%04d' % i)))
    return out


def test_reads_index_pickle_equals_reference_and_round_trips(tmp_path):
    from tombo_b200 import formats, tombo_helper as th
    g = json.load(open(os.path.join(HERE, 'golden', 'formats.json')))
    ix = formats.ReadsIndex('/data/run1')
    assert ix.index_fn == g['index_fn']
    for chrm, strand, rd in _reads(th):
        ix.add_read_data(chrm, strand, rd)
    ref = pickle.loads(base64.b64decode(g['index_pickle_b64']))
    assert ix.records() == ref
    assert ix.dumps() == base64.b64decode(g['index_pickle_b64'])     # byte for byte
    # round trip through a real file
    ix2 = formats.ReadsIndex(str(tmp_path / 'run1'))
    for chrm, strand, rd in _reads(th):
        ix2.add_read_data(chrm, strand, rd._replace(fn=rd.fn.replace('/data/run1', str(tmp_path / 'run1'))))
    fn = ix2.write_index_file()
    assert os.path.basename(fn) == '.run1.RawGenomeCorrected_000.tombo.index'
    back = formats.load_index(fn, str(tmp_path / 'run1') + '/')
    assert sorted(back) == [('chr1', '+'), ('chr2', '-')]
    flat = [rd for k in sorted(back) for rd in back[k]]
    assert len(flat) == 7 and all(rd.fn.startswith(str(tmp_path / 'run1')) for rd in flat)
    assert {rd.read_id for rd in flat} == {'id-%04d' % i for i in range(7)}


class _Node(dict):
    def __init__(self, data=None, kwargs=None):
        dict.__init__(self)
        self.attrs, self.data, self.kwargs = {}, data, kwargs or {}

    def create_group(self, name):
        self[name] = _Node()
        return self[name]

    def create_dataset(self, name, data=None, **kw):
        self[name] = _Node(data=np.array(data), kwargs=kw)
        return self[name]


def test_per_read_block_equals_reference():
    from tombo_b200 import formats
    g = json.load(open(os.path.join(HERE, 'golden', 'formats.json')))
    a = np.load(os.path.join(HERE, 'golden', 'formats_block.npz'))
    off = a['off']
    stats = [a['stats'][off[i]:off[i + 1]] for i in range(off.shape[0] - 1)]
    locs = [a['locs'][off[i]:off[i + 1]] for i in range(off.shape[0] - 1)]
    ids = [(str(rid).encode(), int(off[i + 1] - off[i])) for i, rid in enumerate(a['ids'])]
    block, lookup = formats.per_read_block(stats, locs, ids)
    ref = a['block']
    assert block.dtype == ref.dtype and block.shape == ref.shape
    assert np.array_equal(block['pos'], ref['pos']) and np.array_equal(block['stat'], ref['stat'])
    # the integer a read id maps to is arbitrary (the reference enumerates a set); the
    # mapping must be a bijection that names the same read on every row
    ref_lookup = dict((k, v) for k, v in g['lookup_sorted'])
    inv, ref_inv = {v: k for k, v in lookup.items()}, {v: k for k, v in ref_lookup.items()}
    assert sorted(lookup) == sorted(ref_lookup)
    assert [inv[i] for i in block['read_id']] == [ref_inv[i] for i in ref['read_id']]
    f = _Node()
    w = formats.PerReadStatsWriter(f, 'model_compare', 10000)
    w.write_block(block, lookup, 'chr3', '+', 7000)
    blk = f['Statistic_Blocks']['Block_0']
    assert dict(blk.attrs) == g['block_attrs']
    assert sorted(blk.keys()) == g['block_children']
    assert np.array_equal(blk['block_stats'].data['pos'], ref['pos'])
    assert sorted(zip(blk['read_ids'].data.tolist(), blk['read_id_vals'].data.tolist())) == \
        sorted(lookup.items())
    assert w.curr_block_num == 1 and f.attrs['stat_type'] == 'model_compare'
