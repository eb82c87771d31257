"""Integer-only synthetic PCM16 corpus ("Synthetic inputs").

Every sample is a pure function of (kind, stream id, sample index, channel) built from
32-bit integer arithmetic and two small int16 look-up tables, so the numpy generator here
and the CUDA generator in csrc/corpus.cu produce bit-identical int16 streams.  That is what
lets the CPU reference (oracle/_ref) and the GPU path be fed the *same* input without
shipping audio files.

kinds
  "music"      8 partials 55..7040 Hz (per-stream hashed detune and start phase, right channel
               detuned x1.003), 0.5 Hz triangle AM, white noise at about -26 dBFS, peak ~ -5 dBFS.
  "transient"  -60 dBFS noise floor, 4-sample +-0.9 FS impulses every 6300 samples, 20 ms
               exponentially decaying noise bursts every 14700 samples.
"""
from __future__ import annotations

import numpy as np

RATE_HZ = 44100
KIND_MUSIC = 0
KIND_TRANSIENT = 1
_KINDS = {"music": KIND_MUSIC, "transient": KIND_TRANSIENT}

SINE_LUT_BITS = 10
BURST_LEN = 882  # 20 ms at 44.1 kHz
PARTIAL_AMP = (6000, 4200, 3000, 2100, 1500, 1000, 700, 500)


def sine_lut() -> np.ndarray:
    """int16[1024] sine table.  Built on the host in float64 and handed to the GPU as data."""
    i = np.arange(1 << SINE_LUT_BITS, dtype=np.float64)
    return np.round(32767.0 * np.sin(2.0 * np.pi * i / (1 << SINE_LUT_BITS))).astype(np.int16)


def burst_lut() -> np.ndarray:
    """int16[882] exponential burst envelope (falls to e^-5 over 20 ms)."""
    k = np.arange(BURST_LEN, dtype=np.float64)
    return np.round(32767.0 * np.exp(-k / 176.4)).astype(np.int16)


def _mix32(x: np.ndarray) -> np.ndarray:
    """murmur3 fmix32 on uint32 arrays (wraps modulo 2^32)."""
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x85EBCA6B)
    x ^= x >> np.uint32(13)
    x *= np.uint32(0xC2B2AE35)
    x ^= x >> np.uint32(16)
    return x


def stream_seed(stream_id: int, kind: int) -> int:
    return ((0x9E3779B9 * (stream_id + 1)) ^ (0x51ED270B * (kind + 1))) & 0xFFFFFFFF


def partial_params(stream_id: int, kind: int = KIND_MUSIC):
    """Per-stream phase increments/start phases: (inc[2][8], ph0[2][8]) as python ints (uint32)."""
    seed = stream_seed(stream_id, kind)
    inc = [[0] * 8 for _ in range(2)]
    ph0 = [[0] * 8 for _ in range(2)]
    for p in range(8):
        h = int(_mix32(np.array([seed ^ (0x1000193 * (p + 1))], dtype=np.uint32))[0])
        base = ((55 << p) << 32) // RATE_HZ            # exact integer phase increment for 55*2^p Hz
        detune = (h >> 8) % 2001 - 1000                # +-1 % in units of 1e-5
        left = base + (base * detune) // 100000
        right = left + (left * 3) // 1000              # x1.003
        inc[0][p] = left & 0xFFFFFFFF
        inc[1][p] = right & 0xFFFFFFFF
        ph0[0][p] = (h * 0x9E3779B1) & 0xFFFFFFFF
        ph0[1][p] = (h * 0x85EBCA77) & 0xFFFFFFFF
    return inc, ph0


def _noise16(seed: int, t: np.ndarray, ch: int) -> np.ndarray:
    """Counter-hash white noise, uniform in [-32768, 32767], int32."""
    ctr = (t.astype(np.uint32) * np.uint32(2) + np.uint32(ch)) * np.uint32(0x9E3779B1)
    h = _mix32(ctr ^ np.uint32(seed))
    return (h & np.uint32(0xFFFF)).astype(np.int32) - 32768


def generate(stream_id: int, n_samples: int, kind: str = "music", t0: int = 0) -> np.ndarray:
    """int16[n_samples, 2] for samples t0 .. t0+n_samples-1 of stream `stream_id`."""
    k = _KINDS[kind]
    seed = stream_seed(stream_id, k)
    t = np.arange(t0, t0 + n_samples, dtype=np.int64)
    out = np.empty((n_samples, 2), dtype=np.int16)
    if k == KIND_MUSIC:
        lut = sine_lut().astype(np.int32)
        inc, ph0 = partial_params(stream_id, k)
        am_phase = int(_mix32(np.array([seed ^ 0xA5A5A5A5], dtype=np.uint32))[0]) % 88200
        tri = np.abs(((t + am_phase) % 88200) - 44100)              # 0..44100
        gain = 16384 + (16384 * tri) // 44100                       # 16384..32768
        tu = t.astype(np.uint64)
        for ch in range(2):
            acc = np.zeros(n_samples, dtype=np.int64)
            for p in range(8):
                phase = (np.uint64(ph0[ch][p]) + np.uint64(inc[ch][p]) * tu) & np.uint64(0xFFFFFFFF)
                idx = (phase >> np.uint64(32 - SINE_LUT_BITS)).astype(np.int64)
                acc += PARTIAL_AMP[p] * lut[idx]
            tone = acc >> 15
            tone = (tone * gain) >> 15
            noise = (_noise16(seed, t, ch).astype(np.int64) * 2844) >> 15
            out[:, ch] = np.clip(tone + noise, -32768, 32767).astype(np.int16)
    else:
        env = burst_lut().astype(np.int64)
        imp_off = int(_mix32(np.array([seed ^ 0x0BADC0DE], dtype=np.uint32))[0]) % 6300
        bur_off = int(_mix32(np.array([seed ^ 0x00C0FFEE], dtype=np.uint32))[0]) % 14700
        ti = (t + imp_off) % 6300
        tb = (t + bur_off) % 14700
        for ch in range(2):
            n16 = _noise16(seed, t, ch).astype(np.int64)
            s = (n16 * 57) >> 15                                    # -60 dBFS floor
            imp = np.where(ti < 4, np.where((ti & 1) == 0, 29491, -29491), 0)
            e = np.where(tb < BURST_LEN, env[np.minimum(tb, BURST_LEN - 1)], 0)
            burst = (((n16 * e) >> 15) * 16384) >> 15
            out[:, ch] = np.clip(s + imp + burst, -32768, 32767).astype(np.int16)
    return out
