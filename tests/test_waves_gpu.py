"""Launch-shape parity (-m gpu): several waves of streams through every persistent kernel, every stream compared with the
oracle.  The benchmark launches 4096 streams on a few hundred resident CTAs, so each CTA decodes / encodes many streams
back to back; these tests are what holds that shape to the reference (see tests/wavecheck.py)."""
import os
import subprocess
import sys

import pytest

from tests import wavecheck

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from ulc_codec_b200 import capi as m
    return m


@pytest.mark.parametrize("fmt", ["pcm16", "f32"])
def test_decode_three_waves_every_stream(capi, fmt):
    # 2300 streams x 5..7 blocks (odd and even counts, long and window-switched blocks mixed): > 3 waves of the 740 CTA slots
    assert wavecheck.check_decode_waves(capi, 2300, fmt=fmt) == 2300


@pytest.mark.parametrize("fmt", ["pcm16", "f32"])
def test_decode_three_waves_with_block_index(capi, fmt):
    # same streams with the caller-supplied block-size index (the block-parallel pipeline's direct entry)
    if not getattr(capi, "HAS_BLOCK_INDEX", False):
        pytest.skip("this build has no block-index entry")
    assert wavecheck.check_decode_waves(capi, 2300, fmt=fmt, with_index=True) == 2300


@pytest.mark.parametrize("chunks,overlap,with_index", [(4, 1, False), (5, 2, False), (3, 2, True), (1, 0, False), (0, 0, False)])
def test_decode_chunked_pipeline_every_stream(capi, chunks, overlap, with_index):
    """ULCB_SetDecodePipeline: the call's streams in contiguous chunks, upload + walk (overlap 1) and parse (overlap 2, two arena
    halves) of the next chunk on side streams beside the transform of the current one.  Ragged chunk sizes (2300 / 4, / 5, / 3),
    more chunks than arena halves, every stream against the oracle; (1, 0) = pipelining off, (0, 0) = the library's choice (none for these pageable
    buffers; test_decode_pinned_output_default_pipeline covers the page-locked case)."""
    assert wavecheck.check_decode_waves(capi, 2300, fmt="pcm16", with_index=with_index, pipeline=(chunks, overlap)) == 2300


@pytest.mark.parametrize("fmt,with_index", [("pcm16", False), ("f32", False), ("pcm16", True)])
def test_decode_pinned_output_default_pipeline(capi, fmt, with_index):
    """Page-locked output buffer, library defaults: the PCM is stored by the kernel into the caller's buffer and the call runs as the
    pipeline in chunks of 128 streams (chunked upload on the copy stream, walk + parse on the side stream, two arena halves)."""
    assert wavecheck.check_decode_waves(capi, 2300, fmt=fmt, with_index=with_index, pinned=True) == 2300


@pytest.mark.parametrize("env", [{"ULCB_DECODE_GENERIC": "1"}, {"ULCB_DECODE_WS": "1"}])
def test_decode_waves_other_kernels(env):
    """The kernels ulcb_decode_launch does not pick by default for this shape (selected through the environment)."""
    r = subprocess.run([sys.executable, "-m", "tests.wavecheck", "decode", "1600", "pcm16"], cwd=ROOT, env={**os.environ, **env},
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "1600 streams verified" in r.stdout


def test_encode_three_waves_every_stream_vbr(capi):
    # 450 streams x 5 blocks on <= 148 x k resident CTAs
    assert wavecheck.check_encode_waves(capi, 450, mode="vbr", rate=50.0) == 450


def test_encode_three_waves_every_stream_cbr(capi):
    assert wavecheck.check_encode_waves(capi, 450, mode="cbr", rate=96.0) == 450


def test_encode_detector_prepass_in_waves(capi):
    # more streams than one wave of the detector pre-pass takes (one CTA of 32 streams per SM = 4736): the second wave reuses
    # the scratch of the first and writes its own streams' window controls
    assert wavecheck.check_encode_waves(capi, 5000, mode="vbr", rate=50.0) == 5000

