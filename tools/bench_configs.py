#!/usr/bin/env python3
"""Device-timed encode (and decode) of the other BASELINE.json configurations (bench.py covers config 4):
  C1  data/lenna.gif alone          C2  all 50 data/*.gif as one batch
  C3  one 7680x4320 image           C5  one 32768x32768 image at q in {90,80,50,20,10,5}
  S1  16384^2 uniform noise (q 90/50/10)    S2  16384^2 flat image      [optional argv: config name prefixes]
Pixels resident in HBM, CUDA events around tic_encode_batch, best and median of N runs; the C port of the
reference path (oracle/, test infrastructure) is timed on the host next to it where that takes seconds.
Each line also carries the decode side: the streams just produced, still in HBM, decoded by tic_decode_batch
(median ms, synchronisation rounds, pixel identity with the CPU restatement of the reference decoder).
One JSON line per configuration."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(enc, d_imgs, q, runs):
    import torch
    stream = torch.cuda.current_stream()
    out = None
    ms = []
    for i in range(runs + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = enc.encode_batch_device(d_imgs, q, out=out, stream=stream)
        e1.record(stream)
        res.finish()
        out = res.out
        if i >= 2:
            ms.append(e0.elapsed_time(e1))
    st = enc.stats()
    return res, float(np.min(ms)), float(np.median(ms)), st


def main():
    import torch
    import tinyimgcodec_b200 as tic
    from oracle import oracle_lib as O
    from tests.cases import big_synthetic, synthetic_image
    from tests.golden_io import Golden
    enc = tic.get_encoder(0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    gifs = Golden().all_gifs()
    configs = [
        ("C1 lenna 512x512 q50", [gifs["lenna"]], [50]),
        ("C2 50 x 512x512 gifs, one batch, q50", [gifs[k] for k in sorted(gifs)], [50]),
        ("C3 7680x4320 synthetic q50", [synthetic_image(4320, 7680, seed=0)], [50]),
        ("C5 32768x32768 synthetic", [big_synthetic(32768, 32768, seed=5)], [90, 80, 50, 20, 10, 5]),
        # stress distributions of SURVEY.md 8(d): the scan / pack worst case and the 6-bits-per-block floor
        ("S1 16384x16384 uniform noise", [np.random.default_rng(7).integers(0, 256, (16384, 16384), dtype=np.uint8)],
         [90, 50, 10]),
        ("S2 16384x16384 flat (value 77)", [np.full((16384, 16384), 77, np.uint8)], [50]),
    ]
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    if only:
        configs = [c for c in configs if any(c[0].startswith(o) for o in only)]
    for name, imgs, qualities in configs:
        d_imgs = [torch.from_numpy(im).cuda() for im in imgs]
        px = sum(im.size for im in imgs)
        for q in qualities:
            res, best, med, st = timed(enc, d_imgs, q, 10 if px < (1 << 28) else 4)
            nbytes = int(res.sizes.sum().item())
            line = {"config": name, "quality": q, "pixels": px, "stream_bytes": nbytes, "bpp": 8.0 * nbytes / px,
                    "ms_best": best, "ms_median": med, "mpixel_per_s": px / (med * 1e-3) / 1e6,
                    "roofline_frac_whole_call": (px + nbytes) / (med * 1e-3) / 1e9 / peak,
                    "encode_kernel_ms": st["encode_kernel_ms_sum"] / max(1, st["timed_batches"]),
                    "tiles": st["tiles"]}
            offs, sizes = res.offsets.cpu().numpy(), res.sizes.cpu().numpy()
            hs, ws = [im.shape[0] for im in imgs], [im.shape[1] for im in imgs]
            stream = torch.cuda.current_stream()
            dms = []
            for i in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                outs, _ = enc.decode_batch_device((res.out, offs), sizes, hs, ws, stream=stream)
                e1.record(stream)
                torch.cuda.synchronize()
                if i >= 1:
                    dms.append(e0.elapsed_time(e1))
            dst = enc.decode_stats()
            line["decode"] = {"ms_median": float(np.median(dms)), "mpixel_per_s": px / (float(np.median(dms)) * 1e-3) / 1e6,
                              "sync_rounds": int(dst["sync_rounds"]), "subsequences": int(dst["subsequences"]),
                              "phases_ms": {k: dst[k] for k in ("sync_ms", "scan_ms", "scatter_ms", "idct_ms")}}
            if px <= (1 << 26):
                t0 = time.perf_counter()
                ref = [O.compress(im, q) for im in imgs]
                line["cpu_port_1_thread_ms"] = 1e3 * (time.perf_counter() - t0)
                line["identical_to_oracle"] = ref == res.to_bytes()
                t0 = time.perf_counter()
                want = [O.decompress(s) for s in ref]
                line["decode"]["cpu_port_1_thread_ms"] = 1e3 * (time.perf_counter() - t0)
                line["decode"]["identical_to_oracle"] = all(
                    np.array_equal(outs[i].cpu().numpy(), want[i]) for i in range(len(imgs)))
            del outs
            print(json.dumps(line), flush=True)
        del d_imgs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
