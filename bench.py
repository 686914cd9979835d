#!/usr/bin/env python3
"""bench.py — encoded Mpixel/s of the tinyimgcodec encode path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--images M]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[3]): 4096 synthetic 1024x1024 grayscale images, quality 50,
sharded by image across the N GPUs (no collective on the data path; strong scaling: the batch
is fixed, each rank encodes 4096/N images).  A step is one encode of the whole batch.

  value        whole-job Mpixel/s with pixels already resident in HBM and the streams left in
               HBM (CUDA events, max over ranks)
  e2e          the same through the public host API: pinned host pixels -> H2D -> encode -> D2H
  roofline     algorithmic bytes (pixels in + stream bytes out) / average device duration of one
               encode, against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline the reference's own C encoder (oracle/_ref/encode, built from /root/reference/c)
               one process per host core, plus the C port of the Python path (oracle/), both on a
               bounded sample; rank 0, N=1 only

`--impl reference` times only the CPU reference arm (rank 0), printing the same JSON shape.
Only the cpu_baseline / reference legs and the post-run parity spot check touch oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "encoded Mpixel/s"
UNIT = "Mpixel/s"
IMG_H = IMG_W = 1024
QUALITY = 50
TOTAL_IMAGES = 4096


# ---------------------------------------------------------------------------------------------------
# synthetic data: the BASELINE.md §4 generator, seed = global image index, bit for bit what
# tests/cases.py::synthetic_image(1024, 1024, seed) returns.  The two `rng.integers` draws of an image (coarse
# grid, +-10 noise) are made with numpy on the host — worker processes forked before CUDA is touched, results in
# shared memory — and the integer bilinear x32 upsample, the sum and the clip run on the GPU
# (tests/test_host_bench.py::test_bench_generator_is_the_numpy_generator compares the two on the CPU).
# `--generator torch` keeps round 1's variant: same construction, torch's RNG on the device, no host work.
# ---------------------------------------------------------------------------------------------------
def synth_draws(seed, h=IMG_H, w=IMG_W):
    """The random draws of BASELINE.md §4 for one image: (grid uint8 [h/32+2, w/32+2], noise int8 [h, w])."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 256, (h // 32 + 2, w // 32 + 2))
    noise = rng.integers(-10, 11, (h, w))
    return g.astype(np.uint8), noise.astype(np.int8)


def synth_assemble(g, noise, device):
    """Images [n, h, w] uint8 from the draws of n images (torch tensors or arrays: g [n, gh, gw], noise [n, h, w])."""
    import torch
    g = torch.as_tensor(g).to(device=device, dtype=torch.int32)
    noise = torch.as_tensor(noise).to(device=device, dtype=torch.int32)
    n, h, w = noise.shape
    ys = torch.arange(h, device=device)
    xs = torch.arange(w, device=device)
    gy, fy = ys // 32, (ys % 32).view(1, h, 1).to(torch.int32)
    gx, fx = xs // 32, (xs % 32).view(1, 1, w).to(torch.int32)
    a = g[:, gy][:, :, gx]
    b = g[:, gy][:, :, gx + 1]
    c = g[:, gy + 1][:, :, gx]
    d = g[:, gy + 1][:, :, gx + 1]
    base = ((a * (32 - fx) + b * fx) * (32 - fy) + (c * (32 - fx) + d * fx) * fy) // 1024
    return (base + noise).clamp_(0, 255).to(torch.uint8)


def _draw_worker(G, N, first, lo, hi):
    for i in range(lo, hi):
        G[i], N[i] = synth_draws(first + i, N.shape[1], N.shape[2])


class HostDraws:
    """The draws of images first .. first+count-1 in an anonymous shared mapping (not /dev/shm: a container's mount
    may be 64 MB), filled by forked workers.  Create it BEFORE the process initialises CUDA; .arrays(), .close()."""

    def __init__(self, first, count, h=IMG_H, w=IMG_W, workers=None):
        import mmap
        import multiprocessing as mp
        self.count = count
        gh, gw = h // 32 + 2, w // 32 + 2
        self.shape_g, self.shape_n = (count, gh, gw), (count, h, w)
        self.mg = mmap.mmap(-1, max(1, count * gh * gw))
        self.mn = mmap.mmap(-1, max(1, count * h * w))
        G, N = self.arrays()
        workers = max(1, min(workers or len(os.sched_getaffinity(0)), 64, count))
        t0 = time.perf_counter()
        try:
            ctx = mp.get_context("fork")
            procs = []
            for k in range(workers):
                p = ctx.Process(target=_draw_worker, args=(G, N, first, k * count // workers, (k + 1) * count // workers))
                p.start()
                procs.append(p)
            for p in procs:
                p.join()
            if any(p.exitcode != 0 for p in procs):
                raise RuntimeError("a generator worker failed")
            self.how = f"{workers} forked workers"
        except Exception as ex:   # no fork here: the same draws in this process (about 7 ms per image)
            _draw_worker(G, N, first, 0, count)
            self.how = f"in-process ({type(ex).__name__})"
        del G, N
        self.seconds = time.perf_counter() - t0

    def arrays(self):
        return (np.ndarray(self.shape_g, dtype=np.uint8, buffer=self.mg),
                np.ndarray(self.shape_n, dtype=np.int8, buffer=self.mn))

    def close(self):
        for m in (self.mg, self.mn):
            try:
                m.close()
            except Exception:
                pass


def synth_images_numpy(draws, device, chunk=128):
    """draws: HostDraws -> [count, h, w] uint8 on `device`, identical to tests/cases.py::synthetic_image per seed."""
    import torch
    G, N = draws.arrays()
    out = torch.empty(draws.shape_n, dtype=torch.uint8, device=device)
    for lo in range(0, draws.count, chunk):
        hi = min(draws.count, lo + chunk)
        out[lo:hi] = synth_assemble(torch.from_numpy(G[lo:hi]), torch.from_numpy(N[lo:hi]), device)
    del G, N
    return out


def synth_images_device(first, count, device, h=IMG_H, w=IMG_W, chunk=128):
    """`--generator torch`: the same construction with torch's generator on the device (seed per chunk of 128)."""
    import torch
    out = torch.empty((count, h, w), dtype=torch.uint8, device=device)
    for lo in range(0, count, chunk):
        n = min(chunk, count - lo)
        gen = torch.Generator(device=device)
        gen.manual_seed(1_000_003 * (first + lo) + 17)
        g = torch.randint(0, 256, (n, h // 32 + 2, w // 32 + 2), generator=gen, device=device, dtype=torch.int32)
        noise = torch.randint(-10, 11, (n, h, w), generator=gen, device=device, dtype=torch.int32)
        out[lo:lo + n] = synth_assemble(g, noise, device)
        del g, noise
    return out


# ---------------------------------------------------------------------------------------------------
# host side of a rank: run on the CPUs next to the rank's GPU, so that the pinned buffers of the end-to-end
# leg are allocated on that NUMA node and each GPU's H2D / D2H traffic stays on its own socket
# ---------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(index):
    """Returns a short description for the bench line (or why it was not done)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        ideal = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = ideal & allowed
        if not use:
            return f"gpu {phys}: none of its {len(ideal)} local CPUs is in this process's cpuset ({len(allowed)} CPUs)"
        os.sched_setaffinity(0, use)
        return f"gpu {phys}: bound to {len(use)} of {len(ideal)} local CPUs"
    except Exception as ex:
        return f"not bound ({type(ex).__name__}: {ex})"


# ---------------------------------------------------------------------------------------------------
# clocks during the timed region (NVML, sampled from a thread)
# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# CPU arms (test infrastructure: oracle/_ref = the reference's C encoder, oracle/ = C port)
# ---------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _numpy_sample_images(count):
    from tests.cases import synthetic_image
    return [synthetic_image(IMG_H, IMG_W, seed=i) for i in range(count)]


def time_reference_python(cores, images_per_core):
    """The reference's PYTHON encoder — tinyimgcodec.codec.compress, the path BASELINE.json's north_star names — one
    process per host core on distinct synthetic 1024x1024 images (precedent: tests/benchmark.py:12-28).  Unmodified
    reference code (oracle/ref_py_worker.py: sources in the build container, their bytecode on the GPU box), with the
    pure-Python bidict / bitarray stand-ins.  None if the reference is not available."""
    from oracle.ref_harness import reference_python_available
    if not reference_python_available():
        return None
    worker = os.path.join(ROOT, "oracle", "ref_py_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(images_per_core), str(IMG_H), str(IMG_W), str(1000 * i)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE) for i in range(cores)]
    outs, errs = [], []
    for p in procs:
        o, e = p.communicate()
        if p.returncode == 0 and o.strip():
            outs.append(json.loads(o.decode().strip().splitlines()[-1]))
        else:
            errs.append(e.decode(errors="replace").strip().splitlines()[-1:] or [f"rc={p.returncode}"])
    if len(outs) != cores:
        return {"value": None, "error": f"{len(errs)} of {cores} workers failed: {errs[0][0][:200]}"}
    px, slowest = sum(o["pixels"] for o in outs), max(o["seconds"] for o in outs)
    return {"value": px / slowest / 1e6, "unit": UNIT, "cores": cores, "kind": "reference",
            "per_core": px / sum(o["seconds"] for o in outs) / 1e6,
            "sample": f"{cores} procs x {images_per_core} distinct synthetic {IMG_H}x{IMG_W} images, q{QUALITY}, unmodified "
                      f"tinyimgcodec.codec.compress (Python; pure-Python bitarray/bidict stand-ins), slowest process {slowest:.1f} s"}


def time_reference_c_encoder(cores, images_per_core, distinct=64):
    """The reference's own C encoder (c/encode.c), one process per core, raw rows on stdin,
    `med` quality, exactly how tests/cbenchmark.py:24-26 drives it — but with the input in a file
    so that the pipe is not the test.  Emits the flag-bit-30 stream variant (integer AAN DCT)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "encode")
    if not os.path.isfile(exe):
        return None
    imgs = _numpy_sample_images(distinct)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "in.raw")
        with open(path, "wb") as f:
            for i in range(images_per_core):
                f.write(imgs[i % distinct].tobytes())
        # one tall image: encode.c reads 8-row stripes until EOF (c/encode.c:47-60)
        cmd = [exe, str(IMG_W), str(IMG_H * images_per_core), "med"]
        subprocess.run(cmd, stdin=open(path, "rb"), stdout=subprocess.DEVNULL, check=True)   # warm the page cache
        t0 = time.perf_counter()
        procs = [subprocess.Popen(cmd, stdin=open(path, "rb"), stdout=subprocess.DEVNULL) for _ in range(cores)]
        for p in procs:
            p.wait()
        dt = time.perf_counter() - t0
    return cores * images_per_core * IMG_H * IMG_W / dt / 1e6, dt


def time_oracle_port(cores, images_per_core, distinct=16):
    """The C restatement of the Python path (oracle/tic_oracle.c), one thread per core."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle_lib as O
    imgs = _numpy_sample_images(distinct)
    O.compress(imgs[0], QUALITY)

    def work(_):
        for i in range(images_per_core):
            O.compress(imgs[i % distinct], QUALITY)   # ctypes releases the GIL

    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(work, range(cores)))
    dt = time.perf_counter() - t0
    return cores * images_per_core * IMG_H * IMG_W / dt / 1e6, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    per_core = 48
    kind = "reference"
    vals = []
    if args.warmup > 0:
        time_reference_c_encoder(cores, 8)   # one short untimed pass (page cache, CPU clocks)
    steps = max(1, min(args.steps, 5))
    for _ in range(steps):
        r = time_reference_c_encoder(cores, per_core)
        if r is None:
            kind = "port"
            r = time_oracle_port(cores, per_core)
        vals.append(r)
    total_px = steps * cores * per_core * IMG_H * IMG_W
    total_t = sum(dt for _, dt in vals)
    value = total_px / total_t / 1e6
    sample = (f"{cores} procs x {per_core} synthetic 1024x1024 images per step ({min(per_core, 64)} distinct per process) "
              + ("(reference C encoder c/encode.c, quality 'med', flag-bit-30 stream variant: compare with the GPU arm's "
                 "c_variant figures)" if kind == "reference" else "(C port of the Python path, q50)"))
    py = time_reference_python(cores, 4)   # the path north_star names, on the same host cores
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "4096 synthetic 1024x1024 grayscale images, q50 (bounded sample per step)",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "python_reference": py, "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload_key):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get(workload_key)
    except Exception:
        return None


def run_gpu_arm(args):
    # the numpy draws of this rank's images first: the workers are forked, so before anything initialises CUDA
    draws = None
    if args.generator == "numpy":
        _w, _r = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
        _lo, _hi = _r * args.images // _w, (_r + 1) * args.images // _w
        draws = HostDraws(_lo, _hi - _lo, workers=max(1, len(os.sched_getaffinity(0)) // _w))
    import torch
    import torch.distributed as dist
    import tinyimgcodec_b200 as tic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout must stay ONE JSON line: NCCL prints its version banner (and INFO traces, if the caller asked
        # for them) to fd 1 when the communicator comes up, so fd 1 points at stderr until the first barrier
        # has created it
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            torch.cuda.set_device(local_rank)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    n_gpus = world
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "single rank: not bound"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    total = args.images
    lo, hi = rank * total // n_gpus, (rank + 1) * total // n_gpus
    n_local = hi - lo

    enc = tic.get_encoder(local_rank)
    if draws is not None:
        d_images = synth_images_numpy(draws, dev)
        generator = (f"BASELINE.md §4 numpy generator, seed = image index (draws by numpy on the host, {draws.how}, "
                     f"{draws.seconds:.1f} s; integer upsample + clip on the device)")
        draws.close()
    else:
        d_images = synth_images_device(lo, n_local, dev)
        generator = "BASELINE.md §4 synthetic generator evaluated on-device (torch RNG)"
    torch.cuda.synchronize()
    out_cap = int(n_local * IMG_H * IMG_W * 0.5) + (1 << 20)
    d_out = torch.empty(out_cap, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step():
        return enc.encode_batch_device(d_images, QUALITY, out=d_out, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        res = step()
    res.finish()
    stream_bytes_local = int(res.sizes.sum().item())
    stats = enc.stats()

    sampler = ClockSampler(local_rank)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    t_begin.record(stream)
    for i in range(args.steps):
        ev[i][0].record(stream)
        res = step()
        ev[i][1].record(stream)
    t_end.record(stream)
    barrier()
    clocks = sampler.stop()
    res.finish()
    elapsed_ms = t_begin.elapsed_time(t_end)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    if world > 1:
        tt = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tt.item())
        sb = torch.tensor([stream_bytes_local], device=dev, dtype=torch.int64)
        dist.all_reduce(sb)
        stream_bytes_total = int(sb.item())
    else:
        stream_bytes_total = stream_bytes_local
    ms_per_step = elapsed_ms / args.steps
    total_px = total * IMG_H * IMG_W
    value = total_px / (ms_per_step * 1e-3) / 1e6

    # roofline of the dominant kernel, encode_tiles_kernel (one launch per step per rank): its own device
    # time, from CUDA events the library records on the launching stream around that launch in every
    # timed step (include/tinyimgcodec_cuda.h, tic_last_stats [5..7])
    peak, peak_src = measured_peak()
    kst = enc.stats()
    alg_bytes = n_local * IMG_H * IMG_W + stream_bytes_local
    nb = max(1, int(kst["timed_batches"]))
    kernel_ms = kst["encode_kernel_ms_sum"] / nb
    compact_ms = kst["compact_kernel_ms_sum"] / nb
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    wl_key = f"{n_local}x{IMG_H}x{IMG_W}_q{QUALITY}"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(wl_key), "peak_source": peak_src, "kernel": "encode_tiles_kernel",
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": kernel_ms, "launches_timed": nb,
                "step_ms": float(np.mean(step_ms)), "step_ms_min": float(np.min(step_ms)),
                "other_kernels_ms": {"compact_kernel": compact_ms,
                                     "prep + scan_chunks (+ spine) + scan_apply (+ sizes, summary) (by difference, with launch gaps)":
                                         float(np.mean(step_ms)) - kernel_ms - compact_ms},
                "whole_step": {"achieved": alg_bytes / (float(np.mean(step_ms)) * 1e-3) / 1e9,
                               "frac": alg_bytes / (float(np.mean(step_ms)) * 1e-3) / 1e9 / peak},
                "note": "launch_ms = CUDA events around encode_tiles_kernel alone, on its stream, averaged over the "
                        "timed steps; step_ms = events around one whole tic_encode_batch (5 kernels: prep, encode, two scan "
                        "kernels, compact; no copy, no memset); algorithmic bytes = pixels read once + stream bytes written once"}

    # end to end through the public host API: pinned host pixels in, streams back on the host.  Next to it the plain
    # H2D ceiling of the same pinned buffer at the same N (all ranks copy at the same time): the pipeline moves 14 x
    # more bytes in than out, so the end-to-end figure is a host-link number and is reported as a fraction of that ceiling.
    e2e = None
    e2e_parity = None
    h_images = None
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        h_images = torch.empty((n_local, IMG_H, IMG_W), dtype=torch.uint8).pin_memory()   # after bind_to_gpu_numa_node
        h_images.copy_(d_images)
        torch.cuda.synchronize()
        chunk, nbuf = 64, 4
        enc.compress_batch_pinned(h_images, QUALITY, chunk=chunk, nbuf=nbuf)   # warm-up (allocates the pipeline buffers)
        e2e_steps = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(e2e_steps):
            h_out, index = enc.compress_batch_pinned(h_images, QUALITY, chunk=chunk, nbuf=nbuf)
            d2h = sum(s for _, s in index)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:   # parity of what the pipeline returned (host buffer + index), outside the timed region
            from oracle import oracle_lib as O
            host = h_out.numpy()
            picks = sorted({0, 1, chunk - 1, chunk, n_local // 2, n_local - 1} & set(range(n_local)))
            ok = sum(int(host[index[i][0]: index[i][0] + index[i][1]].tobytes() == O.compress(h_images[i].numpy(), QUALITY)) for i in picks)
            e2e_parity = {"checked": len(picks), "identical": ok}
        # plain H2D of the same bytes, same chunking, one stream
        d_tmp = torch.empty((chunk, IMG_H, IMG_W), dtype=torch.uint8, device=dev)
        barrier()
        c0 = time.perf_counter()
        for lo_ in range(0, n_local, chunk):
            hi_ = min(n_local, lo_ + chunk)
            d_tmp[: hi_ - lo_].copy_(h_images[lo_:hi_], non_blocking=True)
        torch.cuda.synchronize()
        h2d_dt = time.perf_counter() - c0
        del d_tmp
        if world > 1:
            tt = torch.tensor([dt, h2d_dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt, h2d_dt = float(tt[0].item()), float(tt[1].item())
        h2d_bytes = n_local * IMG_H * IMG_W * n_gpus
        e2e_gbs = h2d_bytes * e2e_steps / dt / 1e9
        ceil_gbs = h2d_bytes / h2d_dt / 1e9
        e2e = {"value": total_px * e2e_steps / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h * n_gpus,
               "steps": e2e_steps, "api": f"Encoder.compress_batch_pinned(chunk={chunk}, nbuf={nbuf}): H2D / encode / D2H on three streams, "
                                          "no host synchronisation in front of the GPU",
               "timing": "host wall clock around the API calls, max over ranks", "host_affinity_rank0": numa,
               "h2d_gbs_all_gpus": e2e_gbs, "h2d_ceiling_gbs_all_gpus": ceil_gbs, "frac_of_h2d_ceiling": e2e_gbs / ceil_gbs,
               "saturating_link": "host -> device (PCIe Gen5 x16 per GPU; the ceiling is a plain chunked cudaMemcpyAsync of the same "
                                  "pinned buffer, all ranks at once)",
               "parity": e2e_parity}
    except Exception as ex:  # keep the device-timed line even if the host leg cannot run
        e2e = {"value": None, "unit": UNIT, "error": f"{type(ex).__name__}: {ex}"}

    # parity spot check of the benchmarked data against the oracle (outside every timed region)
    parity = {"checked": 0, "identical": 0}
    if rank == 0:
        try:
            from oracle import oracle_lib as O
            # every 16th image of the rank's batch (all of them with --parity-all) against the C restatement of the
            # reference path, on all host threads (the library call releases the GIL)
            import torch
            from concurrent.futures import ThreadPoolExecutor
            streams = res.to_bytes()
            stride = 1 if args.parity_all else max(1, n_local // 256)
            picks = sorted(set(range(0, n_local, stride)) | {0, 1, n_local // 2, n_local - 1})
            with ThreadPoolExecutor(max(1, len(os.sched_getaffinity(0)))) as ex:
                for lo in range(0, len(picks), 256):   # 256 MB of pixels at a time
                    part = picks[lo:lo + 256]
                    px = d_images[torch.tensor(part, device=d_images.device)].cpu().numpy()
                    same = list(ex.map(lambda j: streams[part[j]] == O.compress(px[j], QUALITY), range(len(part))))
                    parity["checked"] += len(part)
                    parity["identical"] += int(sum(same))
            del streams
        except Exception as ex:
            parity["error"] = f"{type(ex).__name__}: {ex}"

    # the same batch as the C-variant stream (TIC_FLAG_C_VARIANT, quality 'med'): the format the reference's C
    # encoder — the CPU arm of this bench — emits, so that arm and this figure code the same thing
    c_variant = None
    try:
        cv_steps = max(1, min(args.steps, 10))
        for _ in range(2):
            cres = enc.encode_batch_device(d_images, "med", out=d_out, stream=stream, c_variant=True)
        cres.finish()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(cv_steps):
            cres = enc.encode_batch_device(d_images, "med", out=d_out, stream=stream, c_variant=True)
        c1.record(stream)
        barrier()
        cres.finish()
        cms = c0.elapsed_time(c1) / cv_steps
        if world > 1:
            tt = torch.tensor([cms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            cms = float(tt.item())
        cst = enc.stats()
        c_variant = {"value": total_px / (cms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": cms, "steps": cv_steps,
                     "encode_kernel_ms": cst["encode_kernel_ms_sum"] / max(1, int(cst["timed_batches"])),
                     "stream_bytes_rank0": int(cres.sizes.sum().item()),
                     "what": "integer-FDCT stream of c/encode.c, quality 'med', device-resident, CUDA events"}
        if h_images is not None:   # the same through the host pipeline: what the reference arm's ratio should be read against
            enc.compress_batch_pinned(h_images, "med", c_variant=True)
            barrier()
            t0 = time.perf_counter()
            enc.compress_batch_pinned(h_images, "med", c_variant=True)
            torch.cuda.synchronize()
            cdt = time.perf_counter() - t0
            if world > 1:
                tt = torch.tensor([cdt], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                cdt = float(tt.item())
            c_variant["e2e"] = {"value": total_px / cdt / 1e6, "unit": UNIT, "steps": 1,
                                "api": "Encoder.compress_batch_pinned(c_variant=True), host wall clock, max over ranks"}
        if rank == 0:
            from oracle import oracle_lib as O
            if O.ref_c_available():   # the reference binary itself, 2 images (outside every timed region)
                outs = cres.to_bytes()
                ok = 0
                for i in (0, n_local - 1):
                    want = O.ref_c_compress(d_images[i].cpu().numpy(), "med")
                    ok += int(outs[i][:-1] == want[: len(outs[i]) - 1])
                c_variant["parity_vs_reference_binary"] = {"checked": 2, "identical_up_to_flush_byte": ok}
        del cres
    except Exception as ex:
        c_variant = {"value": None, "error": f"{type(ex).__name__}: {ex}"}
    h_images = None

    # the same batch with per-image Huffman tables (auto_generate_huffman_table=True, codec.py:146-148): symbol statistics
    # + table construction + encode, all on the device
    auto = None
    try:
        a_steps = max(1, min(args.steps, 5))
        for _ in range(2):
            ares = enc.encode_batch_device(d_images, QUALITY, out=d_out, stream=stream, auto_generate_huffman_table=True)
        ares.finish()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(a_steps):
            ares = enc.encode_batch_device(d_images, QUALITY, out=d_out, stream=stream, auto_generate_huffman_table=True)
        c1.record(stream)
        barrier()
        ares.finish()
        ams = c0.elapsed_time(c1) / a_steps
        if world > 1:
            tt = torch.tensor([ams], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ams = float(tt.item())
        ast_ = enc.stats()
        auto = {"value": total_px / (ams * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ams, "steps": a_steps,
                "encode_kernel_ms": ast_["encode_kernel_ms_sum"] / max(1, int(ast_["timed_batches"])),
                "launches": int(ast_["launches"]), "stream_bytes_rank0": int(ares.sizes.sum().item()),
                "what": "per-image Huffman tables: symbol_stats_kernel + build_tables_kernel + encode_tiles_kernel<1>, "
                        "device-resident, CUDA events around the whole call"}
        if rank == 0:
            from oracle import oracle_lib as O
            outs = ares.to_bytes()
            picks = (0, n_local - 1)
            auto["parity_vs_oracle"] = {"checked": len(picks), "identical": sum(
                int(outs[i] == O.compress(d_images[i].cpu().numpy(), QUALITY, True)) for i in picks)}
            del outs
        del ares
    except Exception as ex:
        auto = {"value": None, "error": f"{type(ex).__name__}: {ex}"}

    # the decode side: the benchmarked batch's own streams, still in HBM, decoded back to
    # pixels by tic_decode_batch (self-synchronising Huffman decode + float64 IDCT); CUDA events around the call
    decode = None
    try:
        if args.no_decode:
            raise RuntimeError("skipped (--no-decode)")
        res = step().finish()
        offs, sizes = res.offsets.cpu().numpy(), res.sizes.cpu().numpy()
        d_px = torch.empty(n_local * IMG_H * IMG_W + 16, dtype=torch.uint8, device=dev)
        hs, ws = [IMG_H] * n_local, [IMG_W] * n_local
        for _ in range(2):
            outs, _st = enc.decode_batch_device((res.out, offs), sizes, hs, ws, pixels=d_px, stream=stream)
        dec_steps = max(1, min(args.steps, 5))
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(dec_steps):
            outs, _st = enc.decode_batch_device((res.out, offs), sizes, hs, ws, pixels=d_px, stream=stream)
        c1.record(stream)
        barrier()
        dms = c0.elapsed_time(c1) / dec_steps
        if world > 1:
            tt = torch.tensor([dms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dms = float(tt.item())
        dst = enc.decode_stats()
        mae = float((outs[0].float() - d_images[0].float()).abs().mean().item())
        decode = {"value": total_px / (dms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": dms, "steps": dec_steps,
                  "phases_ms": {k: dst[k] for k in ("sync_ms", "scan_ms", "scatter_ms", "idct_ms")},
                  "sync_rounds": int(dst["sync_rounds"]), "subsequences": int(dst["subsequences"]),
                  "launches": int(dst["launches"]), "mean_abs_error_image0": mae,
                  "what": "tic_decode_batch + tic_decode_finish on the batch's own q50 streams, device-resident, CUDA "
                          "events around the two calls (no host synchronisation inside tic_decode_batch)"}
        if rank == 0:
            from oracle import oracle_lib as O
            host = res.to_bytes()
            ok = 0
            for i in (0, n_local - 1):
                ok += int(np.array_equal(outs[i].cpu().numpy(), O.decompress(host[i])))
            decode["parity_vs_oracle"] = {"checked": 2, "identical": ok}
            del host
        del outs, d_px
    except Exception as ex:
        decode = {"value": None, "error": f"{type(ex).__name__}: {ex}"}

    line = None
    if rank == 0:
        cpu_baseline = None
        extra = {}
        if n_gpus == 1 and not args.no_cpu:
            cores = host_cores()
            r = time_reference_c_encoder(cores, 64)
            if r is not None:
                cpu_baseline = {"value": r[0], "unit": UNIT, "cores": cores, "kind": "reference",
                                "sample": f"{cores} procs x 64 synthetic 1024x1024 images, reference C encoder "
                                          f"(c/encode.c, 'med', flag-bit-30 variant), {r[1]:.1f} s"}
            p = time_oracle_port(cores, 32)
            extra["cpu_baseline_port"] = {"value": p[0], "unit": UNIT, "cores": cores, "kind": "port",
                                          "sample": f"{cores} threads x 32 synthetic 1024x1024 images, C port of the "
                                                    f"Python path (oracle/tic_oracle.c, q50, byte-identical streams), {p[1]:.1f} s"}
            if cpu_baseline is None:
                cpu_baseline = extra.pop("cpu_baseline_port")
            py = time_reference_python(cores, 4)
            if py is not None:
                extra["cpu_baseline_python_reference"] = py
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "f16 x f16 -> f32 tensor-core fdct (hi/lo split), f64 exact ties, int32 entropy",
            "data": "synthetic",
            "config": {"workload": f"{total} synthetic {IMG_H}x{IMG_W} grayscale images, quality {QUALITY}, "
                                   f"default Huffman tables, sharded by image over {n_gpus} GPU(s)",
                       "images_per_gpu": n_local, "l2": "inputs larger than L2 (per-GPU pixel bytes >> 126 MB)",
                       "generator": generator,
                       "stream_bytes": stream_bytes_total, "bits_per_pixel": 8.0 * stream_bytes_total / total_px,
                       "exact_path": {k: stats[k] for k in ("exact_items", "exact_changed", "blocks", "tiles")}},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(stats["launches"]) * args.steps * n_gpus, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "parity": parity, "c_variant": c_variant, "auto_tables": auto, "decode": decode,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images", type=int, default=TOTAL_IMAGES, help="total images in the batch (default 4096)")
    ap.add_argument("--parity-all", action="store_true", help="compare EVERY stream of the batch with the oracle (default: every 16th)")
    ap.add_argument("--generator", default="numpy", choices=["numpy", "torch"],
                    help="numpy: BASELINE.md §4 generator with seed = image index (default); torch: the same construction with the device RNG")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host end-to-end leg (profiling runs)")
    ap.add_argument("--no-decode", action="store_true", help="skip the decode-side leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
               "--images", str(args.images)] + (["--no-cpu"] if args.no_cpu else []) + (["--no-e2e"] if args.no_e2e else []) + \
              (["--no-decode"] if args.no_decode else []) + ["--generator", args.generator] + \
              (["--parity-all"] if args.parity_all else [])
        return subprocess.call(cmd)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
