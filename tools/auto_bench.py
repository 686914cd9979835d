"""Driver for profiling the per-image-table path and encode(): N synthetic 1024x1024 images, K encodes with
auto_generate_huffman_table=True (symbol_stats_kernel, build_tables_kernel, encode_tiles_kernel<1>), K with the
C-variant stream (encode_tiles_kernel<2>), and one tic_encode_coeffs (coeffs_kernel).  One JSON line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import tinyimgcodec_b200 as tic


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    enc = tic.get_encoder(0)
    d_images = bench.synth_images_device(0, a.images, dev)
    stream = torch.cuda.current_stream(dev)
    out = {}
    for name, kw, q in (("auto", {"auto_generate_huffman_table": True}, 50), ("c_variant", {"c_variant": True}, "med")):
        ms = []
        for _ in range(a.steps + 1):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(stream)
            res = enc.encode_batch_device(d_images, q, stream=stream, **kw)
            t1.record(stream)
            res.finish()
            ms.append(t0.elapsed_time(t1))
        out[name] = {"ms": ms[1:], "stats": enc.stats(), "bytes": int(res.sizes.sum().item())}
    e = enc.encode(d_images[0].cpu().numpy(), 50)
    out["coeffs"] = {"nonzero_ac": int((e["ac"] != 0).sum()), "stats": enc.stats()}
    print(json.dumps({"images": a.images, **out}))


if __name__ == "__main__":
    main()
