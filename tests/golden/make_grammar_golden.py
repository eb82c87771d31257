"""Writes tests/golden/grammar_vectors.json: SHA-256 of what the compiled UNMODIFIED reference decoder (oracle/_ref, scalar build)
produces for the hand-built grammar streams of tests/grammar.py, so the oracle stays pinned on them where /root/reference (and
oracle/_ref) are absent.  Run here, in the container that has the reference:  python tests/golden/make_grammar_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbin          # noqa: E402
from tests import grammar          # noqa: E402

CASES = [(2048, 2, 3), (256, 1, 4), (1024, 2, 5), (8192, 2, 6), (512, 1, 7), (4096, 1, 8)]
N_STREAMS, N_BLOCKS = 6, 24
out = {"n_streams": N_STREAMS, "n_blocks": N_BLOCKS, "cases": []}
for N, C, seed in CASES:
    st, sz = grammar.streams(N_STREAMS, N_BLOCKS, N, C, seed=seed)
    shas, bits_all = [], []
    for s in st:
        pcm, bits = refbin.decode(np.concatenate([s, np.zeros(64, np.uint8)]), N_BLOCKS, n_chan=C, block_size=N)
        shas.append(hashlib.sha256(np.ascontiguousarray(pcm, np.float32).tobytes()).hexdigest())
        bits_all.append(bits.astype(int).tolist())
    out["cases"].append({"block_size": N, "n_chan": C, "seed": seed, "pcm_sha256": shas, "bits": bits_all})
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "grammar_vectors.json"), "w"))
print("wrote", len(out["cases"]), "cases")
