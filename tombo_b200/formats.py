"""On-disk formats either side of the hot path, host only:

* the resquiggled-reads index the reference pickles next to a FAST5 directory
  (``TomboReads.write_index_file`` / ``_load_index_data``, tombo_helper.py:1068-1104,
  1235-1290): what ``tombo resquiggle`` leaves behind for every later command;
* the per-read statistics blocks of a ``PerReadStats`` file (tombo_stats.py:3335-3366) and
  the (pos, stat, read_id) table ``collate_reg_stats`` builds for them (:4136-4154).

The FAST5 ``Events`` group writer is in :mod:`tombo_b200.tombo_helper`.  HDF5 objects are
duck-typed (anything with ``create_group`` / ``create_dataset`` / ``attrs``), h5py itself is
not a requirement of this package."""
import io
import os
import pickle
import re
from collections import OrderedDict, defaultdict

import numpy as np

from . import tombo_helper as th

INDEX_SUFFIX = '.tombo.index'
PER_READ_BLOCKS_NAME = 'Statistic_Blocks'          # tombo_stats.py:114 STAT_BLOCKS_H5_NAME
PER_READ_DTYPE = [(str('pos'), 'u4'), (str('stat'), 'f8'), (str('read_id'), 'u4')]


def index_filename(fast5s_dir, corr_grp='RawGenomeCorrected_000'):
    """``.<fast5s_dir>.<corr_grp>.tombo.index`` next to the directory (tombo_helper.py:1099-1108)"""
    if fast5s_dir.endswith('/'):
        fast5s_dir = fast5s_dir[:-1]
    head, tail = os.path.split(fast5s_dir)
    return os.path.join(head, '.' + tail + '.' + corr_grp + INDEX_SUFFIX)


class ReadsIndex(object):
    """The write side of ``TomboReads`` (``for_writing=True``): collect
    :class:`tombo_helper.readData` per (chrm, strand) and pickle them in the reference's
    record layout, so the reference (or this class) can load the index back."""

    def __init__(self, fast5s_dir, corr_grp='RawGenomeCorrected_000'):
        self.basedir = fast5s_dir if fast5s_dir.endswith('/') else fast5s_dir + '/'
        self.corr_grp = corr_grp
        self.index_fn = index_filename(self.basedir, corr_grp)
        self.reads_index = defaultdict(list)

    def add_read_data(self, chrm, strand, read_data):
        self.reads_index[(chrm, strand)].append(read_data)

    def records(self):
        """{(chrm, strand): [11-tuples]} exactly as write_index_file builds them (:1086-1096)"""
        out = defaultdict(list)
        for chrm_strand, cs_reads in self.reads_index.items():
            for rd in cs_reads:
                grp = rd.corr_group.split('/')
                out[chrm_strand].append((
                    re.sub(self.basedir, '', rd.fn, 1), rd.start, rd.end, rd.read_start_rel_to_raw,
                    grp[0], grp[-1], rd.filtered, rd.rna, rd.sig_match_score, rd.mean_q_score,
                    rd.read_id))
        return dict(out)

    def dumps(self):
        return pickle.dumps(self.records(), protocol=2)      # protocol 2: py2 / py3 readable

    def write_index_file(self):
        with io.open(self.index_fn, 'wb') as fp:
            fp.write(self.dumps())
        return self.index_fn


def load_index(index_fn, fast5s_dir, corr_grp='RawGenomeCorrected_000'):
    """{(chrm, strand): [readData]} from an index file (``_load_index_data``: 8-, 10- and
    11-field records of the successive Tombo versions)"""
    with io.open(index_fn, 'rb') as fp:
        raw = pickle.load(fp)
    if not raw:
        raise th.TomboError('Tombo index file appears to be empty')
    n = len(next(iter(raw.values()))[0])
    if n not in (8, 10, 11):
        raise th.TomboError('Invalid Tombo index file.')
    out = {}
    for (chrm, strand), recs in raw.items():
        rds = []
        for rec in recs:
            fn, start, end, rsrtr, c_grp, s_grp, filtered, rna = rec[:8]
            extra = tuple(rec[8:])
            rds.append(th.readData(start, end, filtered, rsrtr, strand,
                                   os.path.join(fast5s_dir, fn), corr_grp + '/' + s_grp, rna,
                                   *extra))
        out[(chrm, strand)] = rds
    return out


# ---------------------------------------------------------------------------
# per-read statistics blocks
# ---------------------------------------------------------------------------
def per_read_block(stats, stat_locs, read_ids):
    """The table ``collate_reg_stats`` queues for the per-read statistics file
    (tombo_stats.py:4136-4154): ``stats`` / ``stat_locs`` are the per-read arrays of one
    region, ``read_ids`` the ``(read_id, n_stats)`` pairs.  Returns ``(block, lookup)``:
    a structured array (pos u4, stat f8, read_id u4) without the NaN statistics and the
    read_id -> integer lookup that keeps strings out of the table."""
    stats = np.concatenate(stats)
    stat_locs = np.concatenate(stat_locs)
    valid = ~np.isnan(stats)
    rep_ids = []
    for r_id, r_len in read_ids:
        rep_ids.extend([r_id.decode() if isinstance(r_id, bytes) else r_id] * int(r_len))
    if len(rep_ids) != stats.shape[0]:
        raise th.TomboError('read id counts do not match the statistics')
    valid_ids = [rid for rid, ok in zip(rep_ids, valid) if ok]
    lookup = OrderedDict()
    for rid in valid_ids:                      # first-seen order (the reference: set order)
        if rid not in lookup:
            lookup[rid] = len(lookup)
    block = np.empty(int(valid.sum()), dtype=PER_READ_DTYPE)
    block['pos'] = stat_locs[valid]
    block['stat'] = stats[valid]
    block['read_id'] = [lookup[rid] for rid in valid_ids]
    return block, lookup


class PerReadStatsWriter(object):
    """``PerReadStats._write_per_read_block`` (tombo_stats.py:3335-3366) over an h5py-like
    file object: one ``Block_<n>`` group per region with its chrm / strand / start attributes,
    the ``block_stats`` table and the read-id lookup as two parallel datasets."""

    def __init__(self, h5_file, stat_type, region_size):
        self._fp = h5_file
        h5_file.attrs['stat_type'] = stat_type
        h5_file.attrs['block_size'] = region_size
        self.per_read_blocks = h5_file.create_group(PER_READ_BLOCKS_NAME)
        self.curr_block_num = 0

    def write_block(self, block, read_id_lookup, chrm, strand, start):
        grp = self.per_read_blocks.create_group('Block_' + str(self.curr_block_num))
        self.curr_block_num += 1
        grp.attrs['chrm'] = chrm
        grp.attrs['strand'] = strand
        grp.attrs['start'] = start
        grp.create_dataset('block_stats', data=block, compression='gzip')
        ids = np.array(list(read_id_lookup.keys()), dtype=object)
        ds = grp.create_dataset('read_ids', data=ids, compression='gzip')
        grp.create_dataset('read_id_vals', data=np.array(list(read_id_lookup.values())),
                           compression='gzip')
        if hasattr(self._fp, 'flush'):
            self._fp.flush()
        return ds
