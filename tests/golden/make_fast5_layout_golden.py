"""Generates tests/golden/fast5_layout.json + fast5_layout_events.npy by running the
UNMODIFIED reference's write_new_fast5_group (tombo_helper.py:2341-2460) against a
recording stand-in for the HDF5 file (h5py is not installed; only the layout matters:
group names, attribute names / values, the Events dataset).  Run where /root/reference
exists:   python tests/golden/make_fast5_layout_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'oracle'))


class Node(dict):
    """group / dataset stand-in recording what is written"""
    def __init__(self, data=None, kwargs=None):
        dict.__init__(self)
        self.attrs = {}
        self.data = data
        self.kwargs = kwargs or {}

    def __getitem__(self, key):
        node = self
        for part in [q for q in key.split('/') if q]:
            node = dict.__getitem__(node, part)
        return node

    def create_group(self, name):
        self[name] = Node()
        return self[name]

    def create_dataset(self, name, data=None, **kw):
        self[name] = Node(data=np.array(data), kwargs=kw)
        return self[name]


class FakeFile(Node):
    pass


def dump(node):
    out = {'attrs': {k: (v.item() if hasattr(v, 'item') else v) for k, v in node.attrs.items()}}
    if node.data is not None:
        out['dataset'] = {'dtype': [list(map(str, d)) for d in node.data.dtype.descr]
                          if node.data.dtype.names else str(node.data.dtype),
                          'shape': list(node.data.shape), 'kwargs': node.kwargs}
    out['children'] = {k: dump(v) for k, v in node.items()}
    return out


def main():
    import ref_harness as rh
    from tombo_b200 import synthetic as syn
    m = rh.load_reference()
    th = m['th']
    import h5py                                   # the harness' MagicMock
    h5py.File = FakeFile
    kmer_ref, cpos = syn.make_kmer_ref('DNA', 0)
    std_ref, _ = rh.make_models(kmer_ref, cpos)
    sst, p, sp = rh.make_params('DNA', (4.2, 4.2, 200, 1500, 20.0, 40, 750, 2500, 250))
    r = syn.make_read(kmer_ref, cpos, 120, 4242)
    res, err, info = rh.run_read(r.raw, r.genome_seq, std_ref, sst, p, sp, read_index=0)
    assert res is not None, err
    res = res._replace(
        align_info=th.alignInfo('read_id', 'BaseCalled_template', 3, 5, 2, 1, 110, 4),
        genome_loc=th.genomeLocation(1000, '+', 'chr7'))
    out = {}
    arrays = {}
    for tag, compute_sd in (('means_only', False), ('with_sd', True)):
        f = FakeFile()
        f.create_group('Analyses').create_group('RawGenomeCorrected_000')
        with rh.ref_errstate():
            th.write_new_fast5_group(f, 'RawGenomeCorrected_000', res, 'median', compute_sd, rna=False)
        out[tag] = dump(f)
        arrays[tag] = f['Analyses']['RawGenomeCorrected_000']['BaseCalled_template']['Events'].data
    out['input'] = {'segs': res.segs.tolist(), 'genome_seq': res.genome_seq,
                    'read_start_rel_to_raw': int(res.read_start_rel_to_raw),
                    'scale_values': [None if v is None else float(v) for v in res.scale_values],
                    'sig_match_score': float(res.sig_match_score)}
    json.dump(out, open(os.path.join(HERE, 'fast5_layout.json'), 'w'), sort_keys=True)
    np.savez(os.path.join(HERE, 'fast5_layout_events.npz'), raw_signal=res.raw_signal,
             **{k: v for k, v in arrays.items()})
    print('written', list(out))


if __name__ == '__main__':
    main()
