"""The device MT19937 (parallel chunks reached by polynomial jump-ahead) against numpy's own legacy generator:
words and final state must be bit-identical (pinned order D3: center_overdensity, helper.py:196-202, consumes
np.random's global stream)."""
import numpy as np
import pytest

from fof_b200 import _lib

pytestmark = pytest.mark.gpu
CHUNK_WORDS = 2048 * 624   # kMtChunkBlocks blocks of 624 words


def _numpy_words(state, n):
    rs = np.random.RandomState(0)
    rs.set_state(state)
    w = np.frombuffer(rs.bytes(4 * n), dtype="<u4")
    return w, rs.get_state()


@pytest.mark.parametrize("seed,burn,n_words", [
    (1, 0, 10),                       # inside the current block, no regeneration
    (2, 100, 524),                    # exactly to the end of the block
    (3, 7, 5000),                     # one chunk, no jump
    (4, 311, CHUNK_WORDS + 1234),     # 2 chunks: table 0
    (5, 623, 7 * CHUNK_WORDS + 99),   # 8 chunks: tables 0..2 in every combination
    (6, 0, 37 * CHUNK_WORDS),         # 37-38 chunks: tables 0..5
])
def test_words_and_state_match_numpy(seed, burn, n_words):
    rs = np.random.RandomState(seed)
    rs.bytes(4 * burn)
    state = rs.get_state()
    want, want_state = _numpy_words(state, n_words)
    got, got_state = _lib.get_handle(0).mt19937_words(n_words, state=state)
    assert np.array_equal(got, want)
    assert got_state[2] == want_state[2] and np.array_equal(got_state[1], want_state[1])


def test_fresh_seed_state_pos_624():
    np.random.seed(12345)            # a freshly seeded generator has pos == 624
    state = np.random.get_state()
    assert state[2] == 624
    want, want_state = _numpy_words(state, 3 * CHUNK_WORDS + 5)
    got, _ = _lib.get_handle(0).mt19937_words(3 * CHUNK_WORDS + 5)   # advances the global generator
    assert np.array_equal(got, want)
    now = np.random.get_state()
    assert now[2] == want_state[2] and np.array_equal(now[1], want_state[1])


def test_stream_continues_identically_after_device_draws():
    np.random.seed(99)
    _lib.get_handle(0).mt19937_words(2 * CHUNK_WORDS + 17)
    after_device = np.random.random_sample(5)
    np.random.seed(99)
    np.random.bytes(4 * (2 * CHUNK_WORDS + 17))
    assert np.array_equal(after_device, np.random.random_sample(5))


def test_speculation_shortfall_continues_the_stream():
    """The whole-job call starts the generator early for a guessed number of clusters; when the guess is too small the
    stream is continued from the state the first launch left behind.  Force that with a one-cluster guess."""
    from fof_b200.mock import make_catalog
    arr = make_catalog(6000, seed=12, as_frame=False)
    h = _lib.get_handle(0)
    p = _lib.default_params(richness=5, overdensity=-1e9)
    np.random.seed(3)
    Kf, total = h.pipeline_run_np_random(arr, p, 0.5, 2.52, 2.5)
    want = h.pipeline_fetch(Kf, total)
    want_state = np.random.get_state()
    assert Kf > 4                                              # several 1200-word blocks of the stream are needed
    np.random.seed(3)
    state = np.random.get_state()
    key = np.ascontiguousarray(state[1], dtype=np.uint32)
    h.pipeline_begin(arr, p, 0.5, 2.52, 2.5)
    h.check(h.lib.fofb_dist_prefetch_mt(h.h, key.ctypes.data, int(state[2]), 1))   # stream for ONE cluster only
    Kf2, total2, nkey, npos = h.pipeline_finish_mt(key, state[2])
    got = h.pipeline_fetch(Kf2, total2)
    assert (Kf2, total2) == (Kf, total)
    for a, b in zip(got[:5], want[:5]):
        assert np.array_equal(a, b)
    assert np.allclose(got[5], want[5], rtol=1e-12, atol=0)
    assert npos == want_state[2] and np.array_equal(nkey, want_state[1])
