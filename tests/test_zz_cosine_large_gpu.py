"""GPU parity of the cosine-ref matcher (src/sound.rs:23-38, 351-370) at BASELINE.json's config-3 size: 10 000 segments x
1 000 queries, C = 13 - EVERY index and distance bit-equal to the f64 oracle. At this size a length has two query groups
(the paired 2 x 4 register tile) plus stragglers, the launch has >= 8 CTAs per slice (the 32-CTAs-per-SM slicing of large
batches, exact.cu cosine_match_dev) and the dictionary walk crosses every length boundary; the same batch is matched as 4
shards (index_base) and merged on the host by (distance, index), which must give the same answer.

(The file sorts last on purpose: it covers a configuration of the scan that no smaller test reaches.)"""
import numpy as np
import pytest

from oracle import oracle as O
from soundsym_b200 import api, synth
from soundsym_b200._lib import SS_COSINE_REF

pytestmark = pytest.mark.gpu

C = 13


def test_config3_cosine_every_query_vs_oracle():
    ctx = api.Context(0)
    d, doff = synth.segments(10000, C, seed=1234)
    q, qoff = synth.segments(1000, C, seed=5678)
    oidx, odist = O.cosine_match(d, doff, q, qoff, C)
    dev = api.DeviceDictionary(ctx, d, doff)
    idx, dist = dev.match(q, qoff, SS_COSINE_REF, 1)
    assert np.array_equal(idx[:, 0], oidx), "indices differ from the oracle at %s" % (np.argwhere(idx[:, 0] != oidx)[:5].tolist(),)
    assert np.array_equal(dist[:, 0], odist)
    # per-query targets (SoundDictionary::at_distance, src/sound.rs:351-370)
    targets = np.random.default_rng(3).uniform(-1.0, 1.0, size=1000)
    ot, odt = O.cosine_match(d, doff, q, qoff, C, targets=targets)
    it, dt = dev.match(q, qoff, SS_COSINE_REF, 1, targets=targets)
    assert np.array_equal(it[:, 0], ot) and np.array_equal(dt[:, 0], odt)
    # four shards with a global index_base, merged by (distance, index): first minimum wins across shards too
    bounds = [0, 2500, 5000, 7500, 10000]
    best_i = np.zeros(1000, dtype=np.uint32)
    best_d = np.full(1000, 2.0)  # the fold's seed (0, 2.0), src/sound.rs:361
    found = np.zeros(1000, dtype=bool)
    for s0, s1 in zip(bounds[:-1], bounds[1:]):
        shard = api.DeviceDictionary(ctx, d, doff[s0:s1 + 1], index_base=s0)
        si, sd = shard.match(q, qoff, SS_COSINE_REF, 1)
        better = sd[:, 0] < best_d  # shards in index order + strict '<' = the lowest index among equal distances
        best_i[better], best_d[better], found[better] = si[better, 0], sd[better, 0], True
    assert np.array_equal(best_d, odist)
    assert np.array_equal(best_i[found], oidx[found])
