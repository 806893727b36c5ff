"""``tombo.tombo_stats`` surface of the resquiggle hot path, bound to the CUDA
library: k-mer models, signal normalisation, base means, sequence based rescaling,
parameter loading and per-read alternative-model log-likelihood ratios
(tombo_stats.py:203-573, 580-1123, 1505-1597, 2327-2370, 3888-4082).

Everything numeric runs in CUDA kernels through the C ABI (include/tombo_b200.h);
what the reference itself does in plain Python (parameter tuples, dict look-ups,
regex motif search, a two-line numpy score) stays plain Python here."""
import numpy as np

from . import _lib
from . import tombo_helper as th
from ._default_parameters import (
    ALGN_PARAMS_TABLE, SEG_PARAMS_TABLE, RNA_SAMP_TYPE, DNA_SAMP_TYPE,
    MIN_EVENT_TO_SEQ_RATIO, OCLLHR_SCALE, OCLLHR_HEIGHT, OCLLHR_POWER, STALL_PARAMS,
    FM_OFFSET_DEFAULT, SMALLEST_PVAL)

__all__ = [
    'TomboModel', 'AltModel', 'normalize_raw_signal', 'compute_base_means',
    'get_read_seg_score', 'calc_kmer_fitted_shift_scale', 'load_resquiggle_parameters',
    'compute_num_events', 'get_dynamic_prog_params', 'identify_stalls',
    'compute_alt_model_read_stats', 'trim_seq_and_means', 'apply_per_read_thresh',
    'collate_reg_stats', 'calc_damp_fraction', 'calc_window_fishers_method',
    'compute_de_novo_read_stats', 'compute_sample_compare_read_stats']

# E|N(0,1)| = sqrt(2 / pi); the reference evaluates scipy.stats.halfnorm.expect()
# (tombo_stats.py:84), which returns this value
HALF_NORM_EXPECTED_VAL = float(np.sqrt(2.0 / np.pi))
STANDARD_MODEL_NAME = 'standard'
CONST_SD_MODEL = True                     # tombo_stats.py:112
SAMP_COMP_TXT, DE_NOVO_TXT, ALT_MODEL_TXT = 'sample_compare', 'de_novo', 'model_compare'   # :89-91
NORM_TYPES = ('none', 'pA', 'pA_raw', 'median', 'robust_median', 'median_const_scale')
_CODE = {'A': 0, 'C': 1, 'G': 2, 'T': 3}


def _kmer_code(kmer):
    idx = 0
    for b in kmer:
        idx = idx * 4 + _CODE[b]
    return idx


class TomboModel(object):
    """Canonical k-mer model (tombo_stats.py:580-919).  Built from ``kmer_ref`` (a
    list of ``(kmer, mean, sd)``) and ``central_pos``; model files need h5py and
    are outside the hot path."""

    def __init__(self, ref_fn=None, is_text_model=False, kmer_ref=None, central_pos=None,
                 seq_samp_type=None, reads_index=None, fast5_fns=None, minimal_startup=True):
        if kmer_ref is None:
            raise th.TomboError(
                'tombo_b200.TomboModel is initialised from kmer_ref=/central_pos= '
                '(model files are read by the reference with h5py)')
        assert central_pos is not None, (
            'central_pos must be provided is TomboModel is loaded with a kmer_ref')
        self.means, self.sds = {}, {}
        for kmer, kmer_mean, kmer_std in kmer_ref:
            try:
                kmer = kmer.decode()
            except AttributeError:
                pass
            self.means[kmer] = kmer_mean
            self.sds[kmer] = kmer_std
        self.central_pos = central_pos
        self.name = STANDARD_MODEL_NAME
        self.seq_samp_type = seq_samp_type
        self.kmer_width = len(next(k for k in self.means))
        self.inv_var = None
        if not minimal_startup:
            self.inv_var = dict((k, 1 / (s * s)) for k, s in self.sds.items())
        self._tables = None

    def tables(self):
        """dense (means, sds) indexed by the base-4 k-mer code (device layout)"""
        if self._tables is None:
            n = 4 ** self.kmer_width
            m, s = np.full(n, np.nan), np.full(n, np.nan)
            for k, v in self.means.items():
                if all(b in _CODE for b in k):
                    m[_kmer_code(k)] = v
                    s[_kmer_code(k)] = self.sds[k]
            self._tables = (m, s)
        return self._tables

    def reverse_sequence_copy(self):
        rev = TomboModel(kmer_ref=[(k[::-1], m, self.sds[k]) for k, m in self.means.items()],
                         central_pos=self.kmer_width - self.central_pos - 1,
                         seq_samp_type=self.seq_samp_type,
                         minimal_startup=self.inv_var is None)
        return rev

    def get_exp_levels_from_seq(self, seq, rev_strand=False):
        """tombo_stats.py:834-862"""
        seq_kmers = th.get_seq_kmers(seq, self.kmer_width, rev_strand)
        return self.get_exp_levels_from_kmers(seq_kmers)

    def get_exp_levels_from_kmers(self, seq_kmers):
        """tombo_stats.py:864-884"""
        try:
            ref_means = np.array([self.means[kmer] for kmer in seq_kmers])
            ref_sds = np.array([self.sds[kmer] for kmer in seq_kmers])
        except KeyError:
            raise th.TomboError('Invalid sequence encountered from genome sequence.')
        return ref_means, ref_sds


class AltModel(object):
    """Alternative-base k-mer model (tombo_stats.py:922-1123), from ``kmer_ref`` rows
    ``(kmer, pos, mean, sd)``."""

    def __init__(self, ref_fn=None, kmer_ref=None, central_pos=None, alt_base=None, name=None,
                 motif=None, minimal_startup=True):
        if kmer_ref is None:
            raise th.TomboError('tombo_b200.AltModel is initialised from kmer_ref=')
        assert central_pos is not None and alt_base is not None, (
            'central_pos and alt_base must be provided if AltModel is loaded with a kmer_ref')
        self.means, self.sds = {}, {}
        for kmer, pos, kmer_mean, kmer_std in kmer_ref:
            try:
                kmer = kmer.decode()
            except AttributeError:
                pass
            self.means[(kmer, pos)] = kmer_mean
            self.sds[(kmer, pos)] = kmer_std
        self.central_pos = central_pos
        self.alt_base = alt_base
        self.name = name
        if motif is None:
            self.motif = th.TomboMotif(self.alt_base, 1)
        else:
            assert isinstance(motif, th.TomboMotif) and motif.mod_pos is not None
            self.motif = motif
        self.kmer_width = len(next(kmer for kmer, pos in self.means))
        self.inv_var = None
        self._table = None

    def table(self):
        """dense alt means [code, pos] (NaN where absent) -- the device layout"""
        if self._table is None:
            t = np.full((4 ** self.kmer_width, self.kmer_width), np.nan)
            for (k, pos), v in self.means.items():
                t[_kmer_code(k), pos] = v
            self._table = t
        return self._table

    def get_exp_level(self, kmer, pos):
        return self.means.get((kmer, pos), np.nan)

    def get_exp_sd(self, kmer, pos):
        return self.sds.get((kmer, pos), np.nan)

    def get_exp_levels_from_kmers(self, seq_kmers, rev_strand=False):
        """tombo_stats.py:1096-1123"""
        pos_range = (range(self.kmer_width) if rev_strand
                     else range(self.kmer_width - 1, -1, -1))
        ref_means = np.array([self.get_exp_level(k, p) for k, p in zip(seq_kmers, pos_range)])
        ref_sds = np.array([self.get_exp_sd(k, p) for k, p in zip(seq_kmers, pos_range)])
        return ref_means, ref_sds


# ---------------------------------------------------------------------------
# signal normalisation (tombo_stats.py:203-233, 482-573)
# ---------------------------------------------------------------------------
def compute_base_means(all_raw_signal, base_starts):
    """c_new_means over ``base_starts`` (tombo_stats.py:203-215)"""
    return _lib.get_context().new_means(
        np.asarray(all_raw_signal).astype(np.float64), base_starts)


def normalize_raw_signal(
        all_raw_signal, read_start_rel_to_raw=0, read_obs_len=None, norm_type='median',
        outlier_thresh=None, channel_info=None, scale_values=None, event_means=None,
        model_means=None, model_inv_vars=None, const_scale=None):
    """tombo_stats.py:482-573 for the normalisation types the resquiggle path uses:
    ``median``, ``median_const_scale`` and provided ``scale_values``."""
    if read_obs_len is None:
        read_obs_len = all_raw_signal.shape[0] - read_start_rel_to_raw
    if norm_type not in NORM_TYPES and scale_values is None:
        raise th.TomboError(
            'Normalization type ' + norm_type + ' is not a valid option and shift or scale '
            'parameters were not provided.')
    raw_signal = all_raw_signal[read_start_rel_to_raw:read_start_rel_to_raw + read_obs_len]
    if scale_values is None and norm_type not in ('median', 'median_const_scale'):
        raise NotImplementedError(
            'norm_type ' + norm_type + ' is not on the resquiggle path (SURVEY.md 8a)')
    if scale_values is None and norm_type == 'median_const_scale':
        assert const_scale is not None
    st, norm, sv = _lib.get_context().normalize_raw_signal(
        np.asarray(raw_signal, dtype=np.float64), outlier_thresh=outlier_thresh,
        scale_values=scale_values,
        const_scale=const_scale if (scale_values is None and
                                    norm_type == 'median_const_scale') else None)
    if st == 100:
        raise FloatingPointError('divide by zero encountered in signal normalization')
    th._raise_status(st)
    none = lambda v: None if np.isnan(v) else v   # noqa: E731
    return norm, th.scaleValues(sv[0], sv[1], none(sv[2]), none(sv[3]), outlier_thresh)


def identify_stalls(all_raw_signal, stall_params=None, return_metric=False):
    """Mean-window stall detection (tombo_stats.py:269-368 with MEAN_STALL_PARAMS)."""
    if return_metric:
        raise NotImplementedError('return_metric is a plotting debug aid of the reference')
    if stall_params is not None:
        d = dict(STALL_PARAMS)
        given = dict((k, getattr(stall_params, k)) for k in d)
        if given != d:
            raise NotImplementedError('only the default mean-window stall parameters')
    ints = _lib.get_context().identify_stalls(np.asarray(all_raw_signal, dtype=np.float64))
    return [list(map(int, iv)) for iv in ints]


def calc_kmer_fitted_shift_scale(
        prev_shift, prev_scale, r_event_means, r_model_means, r_model_inv_vars=None,
        method='theil_sen', subsample_key=0):
    """Theil-Sen sequence based rescaling (tombo_stats.py:370-450).  Reads with more
    than MAX_POINTS_FOR_THEIL_SEN bases are sub-sampled with a keyed bijection
    (``subsample_key``) instead of the reference's unseeded np.random.choice."""
    if method != 'theil_sen':
        raise NotImplementedError('only method="theil_sen" is on the resquiggle path')
    st, out = _lib.get_context().theil_sen(
        prev_shift, prev_scale, r_event_means, r_model_means, subsample_key)
    th._raise_status(st)
    return out


def get_read_seg_score(r_means, r_ref_means, r_ref_sds):
    """tombo_stats.py:2327-2338 (a numpy one-liner in the reference too)"""
    return np.mean(np.abs((r_means - r_ref_means) / r_ref_sds))


def get_dynamic_prog_params(match_evalue):
    """tombo_stats.py:2364-2370"""
    return HALF_NORM_EXPECTED_VAL + match_evalue, match_evalue


def load_resquiggle_parameters(seq_samp_type, sig_aln_params=None, seg_params=None,
                               use_save_bandwidth=False):
    """tombo_stats.py:1505-1556"""
    if sig_aln_params is None:
        (match_evalue, skip_pen, bandwidth, save_bandwidth, max_half_z_score, band_bound_thresh,
         start_bw, start_save_bw, start_n_bases) = ALGN_PARAMS_TABLE[seq_samp_type.name]
    else:
        (match_evalue, skip_pen, bandwidth, save_bandwidth, max_half_z_score, band_bound_thresh,
         start_bw, start_save_bw, start_n_bases) = sig_aln_params
        bandwidth, save_bandwidth = int(bandwidth), int(save_bandwidth)
        band_bound_thresh, start_bw = int(band_bound_thresh), int(start_bw)
        start_save_bw, start_n_bases = int(start_save_bw), int(start_n_bases)
    if use_save_bandwidth:
        bandwidth = save_bandwidth
    if seg_params is None:
        (running_stat_width, min_obs_per_base, raw_min_obs_per_base,
         mean_obs_per_event) = SEG_PARAMS_TABLE[seq_samp_type.name]
    else:
        (running_stat_width, min_obs_per_base, raw_min_obs_per_base,
         mean_obs_per_event) = seg_params
    z_shift, stay_pen = get_dynamic_prog_params(match_evalue)
    return th.resquiggleParams(
        match_evalue, skip_pen, bandwidth, max_half_z_score, running_stat_width,
        min_obs_per_base, raw_min_obs_per_base, mean_obs_per_event, z_shift, stay_pen,
        seq_samp_type.name == RNA_SAMP_TYPE, band_bound_thresh, start_bw, start_save_bw,
        start_n_bases)


def compute_num_events(signal_len, seq_len, mean_obs_per_event,
                       min_event_to_seq_ratio=MIN_EVENT_TO_SEQ_RATIO):
    """tombo_stats.py:1558-1574"""
    return max(signal_len // mean_obs_per_event, int(seq_len * min_event_to_seq_ratio))


# ---------------------------------------------------------------------------
# per-read alternative-model statistics (tombo_stats.py:3888-4082)
# ---------------------------------------------------------------------------
def _slice_padded(seq, lo, hi):
    """``seq[lo:hi]`` where positions outside the string read as 'N'"""
    return 'N' * max(0, -lo) + seq[max(0, lo):max(0, min(len(seq), hi))] + 'N' * max(0, hi - len(seq))


def trim_seq_and_means(seq, means, r_start, reg_start, reg_end, strand, kmer_width,
                       central_pos, max_motif_bb, max_motif_ab):
    """Clip a read's sequence / base levels to a region (tombo_stats.py:3888-3970).

    Returns ``(kmers, means, r_start, motif_search_seq)``: the k-mers and levels of the
    positions that have a full k-mer inside ``[reg_start - (K-1), reg_end + (K-1))``, the
    genome position of the first testable base, and the sequence window motif searches
    run over (padded with 'N' where a motif would reach outside the read).

    Everything is expressed through two overhangs -- how far the read sticks out of the
    region on the genome's low and high side -- mapped to the read's 5' / 3' end by strand.
    Two slicing quirks of the reference are kept on purpose (they decide which reads raise):
    a zero-width trailing trim empties the array (``x[:-0]``)."""
    flank = kmer_width - 1
    low_over = max(0, reg_start - (r_start + flank))
    high_over = max(0, (r_start + means.shape[0] - flank) - reg_end)
    clip5, clip3 = (low_over, high_over) if strand == '+' else (high_over, low_over)
    if low_over > 0:
        r_start = reg_start - flank
    trimmed_seq = seq[clip5:max(0, len(seq) - clip3)]
    tail = clip3 + kmer_width - central_pos - 1
    trimmed_means = means[clip5 + central_pos:]
    trimmed_means = trimmed_means[:max(0, trimmed_means.shape[0] - tail)] if tail else trimmed_means[:0]
    if trimmed_means.shape[0] < kmer_width:
        raise th.TomboError('Read sequence too short in this region.')
    kmers = th.get_seq_kmers(trimmed_seq, kmer_width)
    if len(kmers) != trimmed_means.shape[0]:
        raise th.TomboError('Mismatching k-mer and mean levels.')
    # motif window: from max_motif_bb bases before the first testable base to
    # max_motif_ab bases after the last one
    lead = clip5 + flank - max_motif_bb
    trail = clip3 + flank - max_motif_ab
    window = _slice_padded(seq, lead, len(seq) - trail) if trail else ''
    return kmers, trimmed_means, r_start + flank, window


def compute_alt_model_read_stats(r_data, std_ref, alt_refs, use_standard_llhr=False,
                                 reg_data=None):
    """tombo_stats.py:3972-4082.  Read data arrive through
    ``tombo_helper.get_multiple_slots_read_centric`` / ``get_raw_read_slot`` (the
    FAST5 seam, :4013-4016); every site's k-mer window is scored on the GPU."""
    reg_start = reg_data.start if reg_data is not None else r_data.start
    reg_end = reg_data.end if reg_data is not None else r_data.end
    max_motif_bb = max([alt_ref.motif.mod_pos - 1 for _, alt_ref in alt_refs])
    max_motif_ab = max([alt_ref.motif.motif_len - alt_ref.motif.mod_pos
                        for _, alt_ref in alt_refs])
    r_means, r_seq = th.get_multiple_slots_read_centric(
        r_data, ['norm_mean', 'base'], r_data.corr_group)
    try:
        read_id = th.get_raw_read_slot(r_data).attrs.get('read_id')
    except Exception:
        read_id = getattr(r_data, 'read_id', None)
    if r_means is None or r_seq is None:
        raise th.TomboError('Read does not contain valid re-squiggled data.')
    r_seq = b''.join(r_seq).decode() if not isinstance(r_seq, str) else r_seq
    r_kmers, r_means, r_start, motif_search_seq = trim_seq_and_means(
        r_seq, np.asarray(r_means, dtype=np.float64), r_data.start, reg_start, reg_end,
        r_data.strand, std_ref.kmer_width, std_ref.central_pos, max_motif_bb, max_motif_ab)
    K = std_ref.kmer_width
    testable_len = r_means.shape[0] - K + 1
    r_ref_means, r_ref_sds = std_ref.get_exp_levels_from_kmers(r_kmers)
    r_ref_vars = np.square(r_ref_sds)
    ctx = _lib.get_context()
    all_poss, all_llhrs = {}, {}
    win = np.arange(K)
    for alt_name, alt_ref in alt_refs:
        search = motif_search_seq[max_motif_bb - (alt_ref.motif.mod_pos - 1):]
        trim_end = max_motif_ab - (alt_ref.motif.motif_len - alt_ref.motif.mod_pos)
        if trim_end > 0:
            search = search[:-trim_end]
        alt_poss = np.array([m.start() for m in alt_ref.motif.motif_pat.finditer(search)],
                            dtype=np.int64)
        if r_data.strand == '+':
            gen_poss = r_start + alt_poss
        else:
            gen_poss = r_start + testable_len - alt_poss - 1
        if alt_poss.shape[0] == 0:
            all_llhrs[alt_name], all_poss[alt_name] = np.array([]), np.array([])
            continue
        idx = alt_poss[:, None] + win[None, :]
        means_w = r_means[idx]
        ref_w = r_ref_means[idx]
        alt_w = np.array([alt_ref.get_exp_levels_from_kmers(r_kmers[p:p + K])[0]
                          for p in alt_poss])
        if CONST_SD_MODEL:
            mode = 1 if use_standard_llhr else 0
            llhrs = ctx.calc_llh_ratio_windows(mode, means_w, ref_w, alt_w, r_ref_vars[alt_poss],
                                               None, OCLLHR_SCALE, OCLLHR_HEIGHT, OCLLHR_POWER)
        else:
            if not use_standard_llhr:
                raise th.TomboError('Variable SD scaled likelihood ratio not implemented.')
            alt_v = np.array([np.square(alt_ref.get_exp_levels_from_kmers(
                r_kmers[p:p + K])[1]) for p in alt_poss])
            llhrs = ctx.calc_llh_ratio_windows(2, means_w, ref_w, alt_w, r_ref_vars[idx], alt_v)
        all_llhrs[alt_name] = llhrs
        all_poss[alt_name] = gen_poss
    return all_llhrs, all_poss, read_id


# ---------------------------------------------------------------------------
# per-position aggregation (tombo_stats.py:4084-4178, 2537-2552)
# ---------------------------------------------------------------------------
def _region_counts(stats, stat_locs, single_read_thresh, lower_thresh, stat_type, device=0):
    """dense-counter aggregation of (position, statistic) pairs on the device"""
    ctx = _lib.get_context(device)
    stat_locs = np.asarray(stat_locs, dtype=np.int64)
    lo, hi = int(stat_locs.min()), int(stat_locs.max())
    ctx.region_stats_begin(lo, hi - lo + 1)
    ctx.region_stats_add(stats, stat_locs, single_read_thresh, lower_thresh,
                         0 if stat_type == ALT_MODEL_TXT else 1)
    return ctx.region_stats_finalize()


def apply_per_read_thresh(reg_base_stats, single_read_thresh, lower_thresh, stat_type,
                          stat_locs, ctrl_cov=None):
    """tombo_stats.py:4084-4122 -> (reg_frac_std_base, reg_cov, ctrl_cov, valid_cov).
    ``reg_base_stats`` is the per-position list of statistic arrays the reference builds;
    thresholds and counts run on the device."""
    n_pos = len(reg_base_stats)
    lens = np.array([b.shape[0] for b in reg_base_stats], dtype=np.int64)
    flat = np.concatenate(reg_base_stats) if n_pos else np.zeros(0)
    agg = _region_counts(flat, np.repeat(np.arange(n_pos, dtype=np.int64), lens),
                         single_read_thresh, lower_thresh, stat_type) if flat.shape[0] else None
    frac = np.full(n_pos, np.nan)
    reg_cov = lens.copy()
    valid_cov = np.zeros(n_pos, dtype=np.int64)
    if agg is not None:
        frac[agg['pos']] = agg['frac']
        valid_cov[agg['pos']] = agg['valid_cov']
    if stat_type == SAMP_COMP_TXT:
        ctrl_cov = [ctrl_cov[pos] if ctrl_cov is not None and pos in ctrl_cov else 0
                    for pos in stat_locs]
    else:
        ctrl_cov = [0] * int(np.asarray(stat_locs).shape[0])
    return frac, reg_cov, ctrl_cov, valid_cov


def collate_reg_stats(stats, stat_locs, read_ids, per_read_q, reg_data, single_read_thresh,
                      lower_thresh, stat_type, stat_name, ctrl_cov):
    """tombo_stats.py:4124-4178 -> :class:`tombo_helper.regionStats`.  The reference sorts
    the region's (position, statistic) pairs and splits them per position; here the pairs
    go to dense per-position counters on the device (a counting sort) and come back as the
    covered positions in ascending order.  ``per_read_q`` (the per-read statistics writer)
    is outside the hot path and must be None."""
    if per_read_q is not None:
        raise NotImplementedError('per-read statistics blocks are written by the reference')
    stats = np.concatenate(stats)
    stat_locs = np.concatenate(stat_locs).astype(np.int64)
    keep = ~np.isnan(stats)
    if not keep.any():
        raise th.TomboError('No valid positions in this region.')
    agg = _region_counts(stats, stat_locs, single_read_thresh, lower_thresh, stat_type)
    locs_sorted = np.sort(stat_locs[keep])
    if stat_type == SAMP_COMP_TXT:
        cc = [ctrl_cov[pos] if ctrl_cov is not None and pos in ctrl_cov else 0
              for pos in locs_sorted]
    else:
        cc = [0] * int(locs_sorted.shape[0])
    return th.regionStats(agg['frac'], agg['pos'], reg_data.chrm, reg_data.strand,
                          reg_data.start, agg['cov'], cc, agg['valid_cov'])


def calc_damp_fraction(cov_damp_counts, fracs, valid_cov):
    """tombo_stats.py:2537-2552 (elementwise; the device evaluates the same expression inside
    tb2_region_stats_finalize for the fused path)"""
    non_mod_counts = np.round(fracs * valid_cov)
    return (non_mod_counts + cov_damp_counts['unmod']) / (
        valid_cov + sum(list(cov_damp_counts.values())))


# ---------------------------------------------------------------------------
# de novo / sample-compare per-read tests (tombo_stats.py:2252-2271,
# 3675-3873)
# ---------------------------------------------------------------------------
def calc_window_fishers_method(pvals, lag):
    """tombo_stats.py:2252-2271 on the device (1-D input): Fisher's method over a moving
    window of 2 * lag + 1 p-values; NaN in the first / last ``lag`` positions."""
    assert lag > 0, 'Invalid p-value window provided.'
    pvals = np.asarray(pvals, dtype=np.float64)
    if pvals.ndim != 1:
        raise NotImplementedError('1-D p-value vectors only')
    if pvals.shape[-1] < (lag * 2) + 1:
        raise th.TomboError("P-values vector too short for Fisher's Method window compuation.")
    return _lib.get_context().window_fisher_pvals(
        pvals, None, None, np.array([0, pvals.shape[0]]), lag, False)


def _clip_to_region(r_means, r_seq, read_start, read_end, strand, reg_start, reg_end, lo_lag,
                    hi_lag):
    """keep the part of a read whose statistics fall inside [reg_start, reg_end): the read may
    stick out by its low / high lag (k-mer context + Fisher window)"""
    low_over = max(0, reg_start - (read_start + lo_lag))
    high_over = max(0, (read_end - hi_lag) - reg_end)
    clip5, clip3 = (low_over, high_over) if strand == '+' else (high_over, low_over)
    n = r_means.shape[0]
    r_means = r_means[clip5:max(clip5, n - clip3)] if clip3 or clip5 else r_means
    if r_seq is not None:
        r_seq = r_seq[clip5:max(clip5, len(r_seq) - clip3)] if clip3 or clip5 else r_seq
    if low_over:
        read_start = reg_start - lo_lag
    if high_over:
        read_end = reg_end + hi_lag
    return r_means, r_seq, read_start, read_end


def compute_de_novo_read_stats(r_data, std_ref, fm_offset=FM_OFFSET_DEFAULT, reg_data=None):
    """tombo_stats.py:3771-3873 -> ({'de_novo': p-values}, {'de_novo': positions}, read_id).
    Read data arrive through ``tombo_helper.get_multiple_slots_read_centric`` /
    ``get_raw_read_slot`` (the FAST5 seam); z-scores, p-values and the Fisher window run on
    the device."""
    reg_start = reg_data.start if reg_data is not None else r_data.start
    reg_end = reg_data.end if reg_data is not None else r_data.end
    dnstrm = std_ref.kmer_width - std_ref.central_pos - 1
    begin_lag, end_lag = (std_ref.central_pos, dnstrm) if r_data.strand == '+' else \
        (dnstrm, std_ref.central_pos)
    r_means, r_seq = th.get_multiple_slots_read_centric(r_data, ['norm_mean', 'base'],
                                                        r_data.corr_group)
    try:
        read_id = th.get_raw_read_slot(r_data).attrs.get('read_id')
    except Exception:
        read_id = getattr(r_data, 'read_id', None)
    if r_means is None or r_seq is None:
        raise th.TomboError('Read does not contain valid re-squiggled data.')
    r_seq = b''.join(r_seq).decode() if not isinstance(r_seq, str) else r_seq
    r_means, r_seq, read_start, read_end = _clip_to_region(
        np.asarray(r_means, dtype=np.float64), r_seq, r_data.start, r_data.end, r_data.strand,
        reg_start, reg_end, begin_lag + fm_offset, end_lag + fm_offset)
    if len(r_seq) < std_ref.kmer_width:
        raise th.TomboError('Read does not contain information in this region.')
    r_ref_means, r_ref_sds = std_ref.get_exp_levels_from_seq(r_seq, r_data.strand == '-')
    if r_data.strand == '-':
        r_means = r_means[::-1]
    r_means = r_means[begin_lag:r_means.shape[0] - end_lag] if end_lag else r_means[begin_lag:][:0]
    read_start += begin_lag
    read_end -= end_lag
    if fm_offset > 0 and r_means.shape[0] < 2 * fm_offset + 1:
        raise th.TomboError("P-values vector too short for Fisher's Method window compuation.")
    r_pvals = _lib.get_context().window_fisher_pvals(
        np.ascontiguousarray(r_means), r_ref_means, r_ref_sds, np.array([0, r_means.shape[0]]),
        fm_offset, True)
    return {DE_NOVO_TXT: r_pvals}, {DE_NOVO_TXT: np.arange(read_start, read_end)}, read_id


def compute_sample_compare_read_stats(r_data, ctrl_means, ctrl_sds, fm_offset=FM_OFFSET_DEFAULT,
                                      reg_data=None):
    """tombo_stats.py:3675-3769 -> ({'sample_compare': p-values}, {...: positions}, read_id);
    ``ctrl_means`` / ``ctrl_sds`` cover [reg_start - fm_offset, reg_end + fm_offset)."""
    reg_start = reg_data.start if reg_data is not None else r_data.start
    reg_end = reg_data.end if reg_data is not None else r_data.end
    got = th.get_multiple_slots_read_centric(r_data, ['norm_mean'], r_data.corr_group)
    r_means = got[0] if isinstance(got, (tuple, list)) else got
    try:
        read_id = th.get_raw_read_slot(r_data).attrs.get('read_id')
    except Exception:
        read_id = getattr(r_data, 'read_id', None)
    if r_means is None:
        raise th.TomboError('Read does not contain re-squiggled level means.')
    r_means, _, read_start, read_end = _clip_to_region(
        np.asarray(r_means, dtype=np.float64), None, r_data.start, r_data.end, r_data.strand,
        reg_start, reg_end, fm_offset, fm_offset)
    if r_data.strand == '-':
        r_means = r_means[::-1]
    a, b = read_start - reg_start + fm_offset, read_end - reg_start + fm_offset
    cm, cs = np.asarray(ctrl_means[a:b], dtype=np.float64), np.asarray(ctrl_sds[a:b], dtype=np.float64)
    with np.errstate(all='ignore'):
        if np.sum(~np.isnan(np.abs(r_means - cm) / cs)) == 0:
            raise th.TomboError('No valid z-scores in read.')
    if fm_offset > 0 and r_means.shape[0] < 2 * fm_offset + 1:
        raise th.TomboError("P-values vector too short for Fisher's Method window compuation.")
    r_pvals = _lib.get_context().window_fisher_pvals(
        np.ascontiguousarray(r_means), cm, cs, np.array([0, r_means.shape[0]]), fm_offset, False)
    r_poss = np.where(~np.isnan(r_pvals))[0]
    return {SAMP_COMP_TXT: r_pvals[r_poss]}, {SAMP_COMP_TXT: r_poss + read_start}, read_id
