#!/usr/bin/env python3
"""BASELINE configs[2]: a directory of N 1920x1080 PNGs through the reference-shaped CLI pipeline on G GPUs
(`-g 0,..,G-1 -j 4:8:4`): out-MP/s (i) of the proc stage alone (pre-decoded inputs, no encode; image-granular dynamic pull from one
queue like main.cpp:320, the images of the last partial round tile-sharded over the idle GPUs) at 1 and at G GPUs -> strong scaling
of the batch, and (ii) end to end with PNG decode / encode through both command lines.
Usage:  bench_dir.py [--gpus G] [--images N] [--tile 400] [--content noise|natural] [--work DIR] [--skip-e2e]
Not part of bench.py's contract line (that is configs[1]); run it on a multi-GPU box and copy the JSON to profiles/."""
import argparse, json, os, shutil, sys, tempfile, threading, time, queue
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realsr_ncnn_vulkan_b200 as rb
from realsr_ncnn_vulkan_b200 import cli


def natural(h, w, seed):
    """box-blurred noise (SURVEY 8d 'natural' variant): smooth content, full u8 range -- compressible like a photograph"""
    r = np.random.default_rng(seed).integers(0, 256, (h + 8, w + 8, 3)).astype(np.float32)
    c = np.cumsum(np.cumsum(np.pad(r, ((1, 0), (1, 0), (0, 0))), 0, dtype=np.float64), 1)
    k = 7
    b = (c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]) / (k * k)
    b = b[:h, :w]
    b = (b - b.min()) / (b.max() - b.min()) * 255.0
    return np.clip(np.rint(b), 0, 255).astype(np.uint8)


def proc_stage(imgs, gpus, tile, proc_threads, shard_tail):
    """wall seconds of the proc stage for all images on `gpus` GPUs: proc_threads threads per GPU pull images from one queue; with
    shard_tail the n % G images of the last partial round are cut into tile ranges over G / (n % G) GPUs each (rsr_process_tiles)"""
    mdir = os.path.join(ROOT, "models", "models-DF2K")
    engines = []
    for g in range(gpus):
        e = rb.RealSR(g)  # one object per GPU shared by its proc threads, like the reference (main.cpp:778-791,:811-828)
        assert e.load(os.path.join(mdir, "x4.param"), os.path.join(mdir, "x4.bin")) == 0, e.last_error()
        e.tilesize = tile
        engines.append(e)
    n = len(imgs)
    rem = n % gpus if (shard_tail and gpus > 1) else 0
    bulk, tail = imgs[:n - rem], imgs[n - rem:]
    q = queue.Queue()
    outs = [[np.empty((4320, 7680, 3), np.uint8) for _ in range(proc_threads)] for _ in engines]

    def proc(e, buf):
        while True:
            im = q.get()
            if im is None:
                return
            assert e.process(im, buf) == 0, e.last_error()

    for e, bufs in zip(engines, outs):  # warm-up: one image per thread (buffers, clocks, tile descriptors)
        for b in bufs:
            assert e.process(imgs[0], b) == 0
    threads = [threading.Thread(target=proc, args=(e, b)) for e, bufs in zip(engines, outs) for b in bufs]
    for im in bulk:
        q.put(im)
    for _ in threads:
        q.put(None)
    t0 = time.time()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if rem:
        per = gpus // rem  # GPUs per leftover image
        def shard(k):
            sub = engines[k * per:(k + 1) * per] if k < rem - 1 else engines[k * per:]
            assert rb.process_sharded(sub, tail[k], outs[k * per][0]) == 0
        st = [threading.Thread(target=shard, args=(k,)) for k in range(rem)]
        for t in st:
            t.start()
        for t in st:
            t.join()
    return time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--images", type=int, default=100)
    ap.add_argument("--tile", type=int, default=400)
    ap.add_argument("--proc-threads", type=int, default=2, help="proc threads per GPU (the reference default)")
    ap.add_argument("--content", default="natural", choices=["noise", "natural"], help="noise PNGs do not compress (100 MB per result)")
    ap.add_argument("--work", default=None, help="work directory (default: a fresh one under /dev/shm if present, else the temp dir)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--e2e-images", type=int, default=0, help="images of the end-to-end legs (default: all)")
    a = ap.parse_args()
    from PIL import Image
    need = (a.e2e_images or a.images) * 110e6  # worst case: incompressible results, ~100 MB each
    roots = [d for d in ("/dev/shm", tempfile.gettempdir(), os.path.join(ROOT, "gpurun_out")) if os.path.isdir(d) and shutil.disk_usage(d).free > need]
    base = a.work or tempfile.mkdtemp(prefix="rsr_dir_", dir=roots[0] if roots else None)
    din, dout = os.path.join(base, "in"), os.path.join(base, "out")
    os.makedirs(din, exist_ok=True); os.makedirs(dout, exist_ok=True)
    imgs = []
    for i in range(a.images):  # seeds 1000..1099
        imgs.append(natural(1080, 1920, 1000 + i) if a.content == "natural" else np.random.default_rng(1000 + i).integers(0, 256, (1080, 1920, 3), dtype=np.uint8))
    mdir = os.path.join(ROOT, "models", "models-DF2K")
    out_mp = a.images * 1920 * 1080 * 16 / 1e6
    res = {"config": "configs[2]: %d x 1920x1080 %s RGB images, models-DF2K, tile %d, %d GPU(s), %d proc threads per GPU"
                     % (a.images, a.content, a.tile, a.gpus, a.proc_threads)}
    # (i) proc stage only, at 1 GPU and at G GPUs on the same box: strong scaling of the batch
    t1 = proc_stage(imgs, 1, a.tile, a.proc_threads, False)
    res["proc_only_1gpu"] = {"seconds": t1, "out_mp_s": out_mp / t1}
    if a.gpus > 1:
        tg = proc_stage(imgs, a.gpus, a.tile, a.proc_threads, False)
        ts = proc_stage(imgs, a.gpus, a.tile, a.proc_threads, True)
        res["proc_only"] = {"seconds": tg, "out_mp_s": out_mp / tg, "speedup_vs_1gpu": t1 / tg, "schedule": "image-granular dynamic queue (main.cpp:320)"}
        res["proc_only_tail_sharded"] = {"seconds": ts, "out_mp_s": out_mp / ts, "speedup_vs_1gpu": t1 / ts,
                                         "schedule": "same + the %d images of the last partial round tile-sharded over the idle GPUs" % (a.images % a.gpus)}
    print(json.dumps(res), flush=True)
    if a.skip_e2e:
        return
    # (ii) end to end (decode -> proc -> encode), reference thread counts, through both command lines
    ne = a.e2e_images or a.images
    for i in range(ne):
        p = os.path.join(din, "%04d.png" % i)
        if not os.path.exists(p):
            Image.fromarray(imgs[i]).save(p, "PNG", compress_level=1)
    e2e_mp = ne * 1920 * 1080 * 16 / 1e6
    argv = ["-i", din, "-o", dout, "-m", mdir, "-t", ",".join([str(a.tile)] * a.gpus), "-g", ",".join(str(g) for g in range(a.gpus)),
            "-j", "4:%d:4" % 8 if a.gpus > 1 else "4:%d:4" % a.proc_threads]
    exe = os.path.join(ROOT, "realsr-ncnn-vulkan_b200", "realsr-b200")
    if os.path.exists(exe):  # the C++ command line (own PNG codec: Paeth + Z_RLE in parallel row bands)
        import subprocess
        t0 = time.time()
        r = subprocess.run([exe] + argv, capture_output=True, text=True)
        t_cpp = time.time() - t0
        sz = sum(os.path.getsize(os.path.join(dout, f)) for f in os.listdir(dout))
        res["end_to_end_cpp"] = {"seconds": t_cpp, "out_mp_s": e2e_mp / t_cpp, "rc": r.returncode, "files_written": len(os.listdir(dout)), "jobs": argv[-1],
                                 "codec": "host/png_codec.h", "images": ne, "output_bytes": sz, "stderr_tail": r.stderr[-300:]}
        for f in os.listdir(dout):
            os.remove(os.path.join(dout, f))
    t0 = time.time()
    rc = cli.main(argv)
    t_e2e = time.time() - t0
    res["end_to_end_python"] = {"seconds": t_e2e, "out_mp_s": e2e_mp / t_e2e, "rc": rc, "files_written": len(os.listdir(dout)), "jobs": argv[-1],
                                "codec": "Pillow PNG", "images": ne}
    print(json.dumps(res), flush=True)
    if not a.work:
        shutil.rmtree(base, ignore_errors=True)


if __name__ == "__main__":
    main()
