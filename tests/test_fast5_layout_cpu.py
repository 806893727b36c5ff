"""CPU: the FAST5 group the drop-in writes (tombo_b200.tombo_helper.write_new_fast5_group,
SURVEY.md 8(f)-3) equals what the unmodified reference writes -- golden recorded by
tests/golden/make_fast5_layout_golden.py against a recording HDF5 stand-in."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


class Node(dict):
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


def dump(node):
    out = {'attrs': {k: (v.item() if hasattr(v, 'item') else v) for k, v in node.attrs.items()}}
    if node.data is not None:
        out['dataset'] = {'dtype': [list(map(str, d)) for d in node.data.dtype.descr]
                          if node.data.dtype.names else str(node.data.dtype),
                          'shape': list(node.data.shape), 'kwargs': node.kwargs}
    out['children'] = {k: dump(v) for k, v in node.items()}
    return out


@pytest.mark.parametrize('tag,compute_sd', [('means_only', False), ('with_sd', True)])
def test_group_layout_and_events_match_reference(orc, tag, compute_sd):
    from tombo_b200 import tombo_helper as th
    g = json.load(open(os.path.join(HERE, 'golden', 'fast5_layout.json')))
    arrs = np.load(os.path.join(HERE, 'golden', 'fast5_layout_events.npz'))
    inp = g['input']
    segs = np.array(inp['segs'], dtype=np.int64)
    res = th.resquiggleResults(
        align_info=th.alignInfo('read_id', 'BaseCalled_template', 3, 5, 2, 1, 110, 4),
        genome_loc=th.genomeLocation(1000, '+', 'chr7'), genome_seq=inp['genome_seq'],
        mean_q_score=10.0, raw_signal=arrs['raw_signal'], segs=segs,
        read_start_rel_to_raw=inp['read_start_rel_to_raw'],
        scale_values=th.scaleValues(*inp['scale_values']),
        sig_match_score=inp['sig_match_score'])
    # per-base means: the test oracle here, the device in production
    if compute_sd:
        means, sds = orc.new_mean_stds(arrs['raw_signal'], segs)
    else:
        means, sds = orc.new_means(arrs['raw_signal'], segs), None
    f = Node()
    f.create_group('Analyses').create_group('RawGenomeCorrected_000')
    th.write_new_fast5_group(f, 'RawGenomeCorrected_000', res, 'median', compute_sd,
                             rna=False, norm_means=means, norm_stds=sds)
    assert dump(f) == g[tag]
    ev = f['Analyses/RawGenomeCorrected_000/BaseCalled_template/Events'].data
    ref = arrs[tag]
    assert ev.dtype == ref.dtype and ev.shape == ref.shape
    for name in ev.dtype.names:
        assert np.array_equal(ev[name], ref[name], equal_nan=(name == 'norm_stdev')), name
