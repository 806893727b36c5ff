"""Golden for tombo_b200/formats.py from the UNMODIFIED reference (oracle/_ref): the pickled
reads index of TomboReads.write_index_file and the per-read statistics block
collate_reg_stats / PerReadStats._write_per_read_block produce, written into a recording
HDF5 stand-in.    python tests/golden/make_formats_golden.py"""
import base64
import json
import os
import queue
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'oracle'))
sys.path.insert(0, HERE)

from make_fast5_layout_golden import Node  # noqa: E402


def reads(th):
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


def main():
    import ref_harness as rh
    m = rh.load_reference()
    th, ts = m['th'], m['ts']
    out = {}
    # ---- reads index ----
    tr = th.TomboReads.__new__(th.TomboReads)
    tr.corr_grp = 'RawGenomeCorrected_000'
    tr.fast5s_dirs = {}
    tr._prep_for_writing(['/data/run1'])
    basedir, index_fn = next(iter(tr.fast5s_dirs.items()))
    for chrm, strand, rd in reads(th):
        tr.add_read_data(chrm, strand, rd)
    written = {}
    import io as _io
    orig_open = _io.open

    class Cap(_io.BytesIO):
        def close(self):
            written['bytes'] = self.getvalue()
            _io.BytesIO.close(self)
    th.io.open = lambda fn, mode: Cap()
    th.status_message = lambda *a, **k: None
    try:
        tr.write_index_file()
    finally:
        th.io.open = orig_open
    out['index_fn'] = index_fn
    out['index_pickle_b64'] = base64.b64encode(written['bytes']).decode()
    # ---- per-read statistics block ----
    rs = np.random.RandomState(6)
    stats, locs, ids = [], [], []
    for i in range(5):
        n = int(rs.randint(5, 30))
        p = np.sort(rs.choice(np.arange(7000, 7100), n, replace=False)).astype(np.int64)
        v = rs.normal(0, 2, n)
        v[rs.randint(0, n)] = np.nan
        stats.append(v); locs.append(p); ids.append((('rid%d' % (i % 4)).encode(), n))

    class Reg(object):
        start, end, chrm, strand = 7000, 7100, 'chr3', '+'
    q = queue.Queue()
    with rh.ref_errstate(), np.errstate(invalid='ignore'):
        ts.collate_reg_stats([s.copy() for s in stats], [l.copy() for l in locs], ids, q, Reg(), 2.5,
                             -1.5, 'model_compare', '5mC', None)
    name, (block, lookup, chrm, strand, start) = q.get()
    prs = ts.PerReadStats.__new__(ts.PerReadStats)
    f = Node()
    f.flush = lambda: None
    prs._fp = f
    prs.per_read_blocks = f.create_group('Statistic_Blocks')
    prs.curr_block_num = 0
    import h5py
    h5py.special_dtype = lambda vlen=None: object

    class DS(Node):
        def __setitem__(self, k, v):
            if k is Ellipsis:
                self.data = np.array(v)
            else:
                dict.__setitem__(self, k, v)
    orig_cd = Node.create_dataset

    def cd(self, name, *a, **kw):
        data = kw.pop('data', None)
        if data is None and a:
            self[name] = DS(data=None, kwargs={k: v for k, v in kw.items() if k != 'dtype'})
            return self[name]
        return orig_cd(self, name, data=data, **kw)
    Node.create_dataset = cd
    try:
        prs._write_per_read_block(block, lookup, chrm, strand, start)
    finally:
        Node.create_dataset = orig_cd
    blk = f['Statistic_Blocks']['Block_0']
    out['block_attrs'] = dict(blk.attrs)
    out['block_children'] = sorted(blk.keys())
    out['lookup_sorted'] = sorted(lookup.items())
    out['name'] = name
    np.savez(os.path.join(HERE, 'formats_block.npz'), block=block,
             stats=np.concatenate(stats), locs=np.concatenate(locs),
             off=np.concatenate([[0], np.cumsum([s.shape[0] for s in stats])]),
             ids=np.array([i[0].decode() for i in ids]),
             read_id_vals=np.asarray(blk['read_id_vals'].data),
             read_ids=np.asarray(blk['read_ids'].data, dtype=str))
    json.dump(out, open(os.path.join(HERE, 'formats.json'), 'w'), sort_keys=True)
    print('written', sorted(out))


if __name__ == '__main__':
    main()
