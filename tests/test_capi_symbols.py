"""CPU: the C-ABI library loads and exports every symbol include/ulcb200.h declares (no compute)."""
import os
import re

from ulc_codec_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ulcb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ULCB?_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    L = capi.load()
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ulcb200.h but not exported"
    assert sorted(capi.EXPORTS) == names


def test_no_device_is_an_error_not_a_fallback():
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        return
    L = capi.load()
    h = C.c_void_p()
    assert L.ULCB_ContextCreate(C.byref(h), 0, 44100, 2, 2048, 16) == -5      # ULCB_E_NODEVICE
    st = capi.ULC_DecoderState_t(); st.nChan, st.BlockSize = 2, 2048
    assert L.ULC_DecoderState_Init(C.byref(st)) == -1
    st.nChan = 0
    assert L.ULC_DecoderState_Init(C.byref(st)) == -1                          # bad params (ulcDecoder.c:30-35)
    assert L.ULCB_ContextCreate(C.byref(h), 0, 44100, 2, 1000, 16) == -1       # not a power of two


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ulc_codec_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("oracle/", "").replace("the oracle", "") or f == "corpus.py", f
