#!/usr/bin/env python
"""Conformance run for BASELINE.json's three criteria on a larger sample than the unit tests (GPU box).

For every case (corpus kind x rate mode x block size) encodes S distinct synthetic streams with the CUDA path and with
the oracle (the C restatement that tests/test_oracle_vs_ref.py pins bit-identical to the reference's scalar build),
and reports
  (2) blocks whose bitstream differs from the oracle's ("flips", allowed: 0.1 % of blocks), window-control and
      complexity mismatches;
  (1) the CUDA decoder on the oracle's bytes vs the oracle decoder: PCM16 max LSB error and share of samples that differ,
      float error relative to the peak;
  (3) SNR of [CUDA-encoded, oracle-decoded] minus SNR of [oracle-encoded, oracle-decoded] against the source, in dB.
usage: python scripts/conformance.py [--streams 24] [--seconds 12] > profiles/rNN_conformance.json
The oracle is the checker here, never the thing measured."""
import argparse, json, os, sys, time
from concurrent.futures import ProcessPoolExecutor
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [("music", "vbr", 50.0, 0.0, 2048), ("music", "vbr", 75.0, 0.0, 2048), ("music", "cbr", 128.0, 0.0, 2048),
         ("transient", "abr", 64.0, 0.7, 2048), ("transient", "vbr", 50.0, 0.0, 256), ("music", "cbr", 96.0, 0.0, 1024),
         ("transient", "abr", 64.0, 0.7, 8192), ("music", "vbr", 60.0, 0.0, 4096)]


def _oracle_one(job):
    from oracle import oracle
    from ulc_codec_b200 import corpus
    sid, ns, kind, mode, rate, avgc, bs = job
    pcm = corpus.generate(sid, ns, kind)
    e = oracle.Encoder(44100, 2, bs).encode_stream(pcm, mode, rate, avgc)
    nb = len(e["block_bytes"])
    dec, _ = oracle.Decoder(2, bs).decode_stream(e["data"], nb)
    return e["data"], e["block_bytes"], e["window_ctrl"], e["complexity"], dec


def snr_db(ref, x):
    ref = ref.astype(np.float64); x = x.astype(np.float64)
    n = ((ref - x) ** 2).sum()
    return 10 * np.log10((ref ** 2).sum() / max(n, 1e-300))


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--streams", type=int, default=24); ap.add_argument("--seconds", type=float, default=12.0)
    a = ap.parse_args()
    from oracle import oracle
    from ulc_codec_b200 import capi, corpus
    ns = int(a.seconds * 44100)
    out = {"streams_per_case": a.streams, "seconds_per_stream": a.seconds, "cases": []}
    tot_blocks = tot_flips = 0
    with ProcessPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        for ci, (kind, mode, rate, avgc, bs) in enumerate(CASES):
            sids = [5000 + 100 * ci + i for i in range(a.streams)]
            t0 = time.time()
            refs = list(ex.map(_oracle_one, [(s, ns, kind, mode, rate, avgc, bs) for s in sids]))
            pcms = [corpus.generate(s, ns, kind) for s in sids]
            nb = len(refs[0][1])
            with capi.Context(44100, 2, bs, max_blocks=nb) as ctx:
                g = ctx.encode_streams(pcms, mode, rate, avgc)
                dpcm16, st16, _ = ctx.decode_streams([r[0] for r in refs], nb, fmt="pcm16")
                dflt, _, _ = ctx.decode_streams([r[0] for r in refs], nb, fmt="f32")
            flips = wc_bad = cx_bad = 0
            lsb_max = 0; lsb_diff = 0.0; ferr = 0.0; dsnr = []
            for i, (rdata, rbb, rwc, rcx, rdec) in enumerate(refs):
                gbb = g["block_bytes"][i].astype(np.int64); go = np.concatenate([[0], np.cumsum(gbb)]); ro = np.concatenate([[0], np.cumsum(rbb.astype(np.int64))])
                gs = g["streams"][i]
                for b in range(nb):
                    if gbb[b] != rbb[b] or not np.array_equal(gs[go[b]:go[b + 1]], rdata[ro[b]:ro[b + 1]]): flips += 1
                wc_bad += int((g["window_ctrl"][i] != rwc).sum()); cx_bad += int((g["complexity"][i] != rcx).sum())
                r16 = oracle.float_to_pcm16(rdec)
                d = np.abs(dpcm16[i, :nb * bs].astype(np.int32) - r16.astype(np.int32))
                lsb_max = max(lsb_max, int(d.max())); lsb_diff += float((d > 0).mean())
                ferr = max(ferr, float(np.abs(dflt[i, :nb * bs] - rdec).max() / max(1e-30, np.abs(rdec).max())))
                # criterion 3: decode the CUDA-encoded stream with the ORACLE decoder, compare SNRs against the delayed source
                gdec, _ = oracle.Decoder(2, bs).decode_stream(gs, nb)
                src = np.zeros((nb * bs, 2), np.float32); src[2 * bs:2 * bs + ns] = pcms[i].astype(np.float32)[: nb * bs - 2 * bs] / 32768.0
                dsnr.append(snr_db(src, gdec) - snr_db(src, rdec))
            blocks = nb * a.streams
            tot_blocks += blocks; tot_flips += flips
            out["cases"].append({"corpus": kind, "mode": mode, "rate": rate, "block_size": bs, "blocks": blocks,
                                 "bitstream_flips": flips, "flip_rate": flips / blocks, "window_ctrl_mismatches": wc_bad,
                                 "complexity_mismatches": cx_bad, "decode_pcm16_max_lsb": lsb_max,
                                 "decode_pcm16_share_differing": lsb_diff / a.streams, "decode_float_err_rel_peak": ferr,
                                 "decode_status_bad": int((st16 != -1).sum()), "snr_delta_db_max_abs": float(np.abs(dsnr).max()),
                                 "seconds": round(time.time() - t0, 1)})
            print(json.dumps(out["cases"][-1]), file=sys.stderr)
    out["total_blocks"] = tot_blocks; out["total_flips"] = tot_flips; out["flip_rate"] = tot_flips / tot_blocks
    out["allowed_flip_rate"] = 1e-3
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
