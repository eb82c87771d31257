"""CPU: the oracle restatement reproduces the reference-generated golden fixtures bit for bit."""
import numpy as np
import pytest

from oracle import oracle
from tests.fixtures import GOLDEN_NAMES, load_golden_stream, sha


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_decoder_matches_reference_golden(name):
    g = load_golden_stream(name)
    nb = len(g["block_bytes"])
    dec, bits = oracle.Decoder(g["n_chan"], g["block_size"]).decode_stream(g["data"], nb)
    assert np.array_equal(bits, g["bits"])
    assert np.array_equal((bits + 7) // 8, g["block_bytes"].astype(np.int64))
    bs = g["block_size"]
    assert np.array_equal(dec[:4 * bs], g["dec_head"])
    assert np.array_equal(dec[-2 * bs:], g["dec_tail"])
    assert sha(dec) == g["dec_sha256"]          # bit-identical to the stock scalar reference build


def test_corpus_is_deterministic():
    from ulc_codec_b200 import corpus
    a = corpus.generate(7, 5000, "music")
    b = corpus.generate(7, 2000, "music", t0=3000)
    assert np.array_equal(a[3000:], b)          # random access in time: pure function of (id, t)
    assert sha(corpus.generate(0, 44100, "music")) == sha(corpus.generate(0, 44100, "music"))
    t = corpus.generate(3, 30000, "transient")
    assert abs(int(t.max())) > 29000 and np.abs(t).mean() < 600


def test_pcm16_rule():
    x = np.array([0.0, 1.0, -1.0, 0.5 / 32768, 1.5 / 32768, 2.5 / 32768, -0.5 / 32768, 0.99999], np.float32)
    assert oracle.float_to_pcm16(x).tolist() == [0, 32767, -32768, 0, 2, 2, 0, 32767]   # round-half-even, clamped


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_encoder_matches_reference_golden(name):
    """The encoder restatement reproduces the reference encoder's bytes, window decisions, complexity,
    MDCT coefficients and ranks (hashes) for every golden case - VBR, CBR and two-pass ABR."""
    from tests.fixtures import golden_pcm
    g = load_golden_stream(name)
    enc = oracle.Encoder(g["rate_hz"], g["n_chan"], g["block_size"]).encode_stream(
        golden_pcm(g), g["mode"], g["param"], g["avg_complexity"], dump=True)
    assert np.array_equal(enc["block_bytes"], g["block_bytes"])
    assert np.array_equal(enc["data"], g["data"])
    assert np.array_equal(enc["window_ctrl"], g["window_ctrl"])
    assert np.array_equal(enc["complexity"], g["complexity"])
    assert np.array_equal(enc["coefs"][:6], g["coefs_head"])
    assert sha(enc["coefs"]) == g["coefs_sha256"]
    assert sha(enc["ranks"]) == g["ranks_sha256"]


def test_oracle_on_grammar_vectors():
    """Hand-built streams with every token of the grammar (tests/grammar.py): the oracle's output must hash to what the compiled
    reference produced (tests/golden/grammar_vectors.json, written by tests/golden/make_grammar_golden.py).  This is what caught the
    oracle applying overlap scaling to decimation layout 0, which no encoder-made golden contains."""
    import hashlib
    import json
    import os
    from tests import grammar
    from tests.fixtures import GOLDEN
    g = json.load(open(os.path.join(GOLDEN, "grammar_vectors.json")))
    for case in g["cases"]:
        N, C = case["block_size"], case["n_chan"]
        st, sz = grammar.streams(g["n_streams"], g["n_blocks"], N, C, seed=case["seed"])
        for i, s in enumerate(st):
            pcm, bits = oracle.Decoder(C, N).decode_stream(np.concatenate([s, np.zeros(64, np.uint8)]), g["n_blocks"])
            assert bits.astype(int).tolist() == case["bits"][i], (N, C, i)
            assert np.array_equal((bits.astype(np.int64) + 7) // 8, sz[i].astype(np.int64))
            assert hashlib.sha256(np.ascontiguousarray(pcm, np.float32).tobytes()).hexdigest() == case["pcm_sha256"][i], (N, C, i)
