"""Decode-side driver for profiling: N synthetic 1024x1024 images encoded on the device (q50), then decoded K
times from the encoder's output buffer.  Prints one JSON line with the per-phase device times."""
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
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--quality", type=int, default=50)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    enc = tic.get_encoder(0)
    d_images = bench.synth_images_device(0, a.images, dev)
    res = enc.encode_batch_device(d_images, a.quality).finish()
    offs, sizes = res.offsets.cpu().numpy(), res.sizes.cpu().numpy()
    n, h, w = d_images.shape
    d_px = torch.empty(n * h * w + 16, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    times = []
    for _ in range(a.steps):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        outs, st = enc.decode_batch_device((res.out, offs), sizes, [h] * n, [w] * n, pixels=d_px, stream=stream)
        t1.record(stream)
        torch.cuda.synchronize()
        times.append(t0.elapsed_time(t1))
    mae = float((outs.float() - d_images.float()).abs().mean().item())
    print(json.dumps({"images": n, "quality": a.quality, "stream_bytes": int(sizes.sum()), "ms": times,
                      "stats": enc.decode_stats(), "mean_abs_error": mae, "status_or": int(st.max())}))


if __name__ == "__main__":
    main()
