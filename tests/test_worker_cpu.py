"""CPU tests of the batch assembler / failure bookkeeping (tombo_b200/worker.py,
SURVEY.md 8(f)-4): driven with an injected resquiggle function, no device needed."""
import collections

import numpy as np
import pytest

MapRes = collections.namedtuple('MapRes', ('raw_signal', 'genome_seq'))


def _reads(n, seed=0):
    rs = np.random.RandomState(seed)
    return [(MapRes(np.zeros(rs.randint(10, 50)), 'ACGT'), 'read_%d.fast5' % i) for i in range(n)]


def test_batches_flush_on_read_and_sample_limits_and_keep_order():
    from tombo_b200 import worker, tombo_helper as th
    calls = []

    def fake(map_results, std_ref, params, save_params, **kw):
        calls.append(len(map_results))
        out = []
        for m in map_results:
            n = len(m.raw_signal)
            out.append(th.TomboError('too short') if n % 7 == 0 else ('ok', n))
        return out

    reads = _reads(103, 3)
    got = list(worker.resquiggle_stream(iter(reads), None, None, None, max_reads=10,
                                        resquiggle_fn=fake))
    assert [fn for fn, _ in got] == [fn for _, fn in reads]
    assert calls == [10] * 10 + [3]
    for (m, fn), (gfn, msg) in zip(reads, got):
        if len(m.raw_signal) % 7 == 0:
            assert msg == [True, ['too short', fn, True]]
        else:
            assert msg == [False, ('ok', len(m.raw_signal))]
    # sample budget: a read that would overflow the budget opens the next batch
    calls[:] = []
    b = worker.ReadBatcher(None, None, None, max_reads=1000, max_samples=100, resquiggle_fn=fake)
    done = []
    for m, fn in reads[:20]:
        done += b.add(m, fn)
    done += b.flush()
    assert len(done) == 20 and sum(calls) == 20 and len(calls) > 2
    assert len(b) == 0 and b.flush() == []


def test_batch_level_exception_is_reported_per_read_not_raised():
    from tombo_b200 import worker, tombo_helper as th

    def boom(map_results, *a, **k):
        raise ValueError('device lost')

    def refuse(map_results, *a, **k):
        raise th.TomboError('model mismatch')

    out = list(worker.resquiggle_stream(iter(_reads(3)), None, None, resquiggle_fn=boom))
    assert all(msg[0] is True and msg[1][2] is False and 'device lost' in msg[1][0] for _, msg in out)
    out = list(worker.resquiggle_stream(iter(_reads(3)), None, None, resquiggle_fn=refuse))
    assert [msg for _, msg in out] == [[True, ['model mismatch', fn, True]] for fn, _ in out]


def test_failure_summary_matches_reference_layout(tmp_path):
    from tombo_b200 import worker
    fs = worker.FailureSummary()
    for i in range(6):
        fs.record([False, object()])
    for i in range(3):
        fs.record([True, ['Read event to sequence alignment extends beyond bandwidth', 'a%d' % i, True]])
    fs.record([True, ['Not enough raw signal around potential genomic deletion(s)', 'b0', True]])
    fs.record([True, ['Traceback ... ZeroDivisionError', 'c0', False]])
    assert fs.num_processed == 11
    assert sorted(fs.counts(), reverse=True)[0] == (3, 'Read event to sequence alignment extends beyond bandwidth')
    txt = fs.format('hdr', fs.counts(), 11, num_errs=4)
    lines = txt.split('\n')
    assert lines[0] == 'hdr' and len(lines) == 5
    assert lines[1].startswith('    27.3% (      3 reads) : Read event to sequence alignment extends beyond bandwidth')
    assert lines[-1] == '     -----'
    assert 'Unexpected error' in txt and fs.non_tombo_errors == ['c0\n:::\nTraceback ... ZeroDivisionError']
    final = fs.final_message(11)
    assert final.startswith('Final unsuccessful reads summary (45.5% reads unsuccessfully processed; 5 total reads):')
    fn = str(tmp_path / 'failed.txt')
    fs.write(fn)
    rows = open(fn).read().rstrip('\n').split('\n')
    assert rows[0] == 'Read event to sequence alignment extends beyond bandwidth\ta0, a1, a2'
    assert len(rows) == 3
    assert worker.FailureSummary().final_message(5) == 'All reads successfully re-squiggled!'
