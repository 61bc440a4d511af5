#!/usr/bin/env python3
"""BASELINE configs[2]: a directory of N 1920x1080 PNGs through the reference-shaped CLI pipeline on G GPUs
(`-g 0,..,G-1 -j 4:8:4`): out-MP/s (i) of the proc stage alone (pre-decoded inputs, no encode) and (ii) end to end with
PNG decode / encode (Pillow).  Usage:  bench_dir.py [--gpus G] [--images N] [--tile 400] [--keep DIR]
Not part of bench.py's contract line (that is configs[1]); run it on a multi-GPU box and copy the JSON to profiles/."""
import argparse, json, os, shutil, sys, tempfile, threading, time, queue
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realsr_ncnn_vulkan_b200 as rb
from realsr_ncnn_vulkan_b200 import cli


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--images", type=int, default=100)
    ap.add_argument("--tile", type=int, default=400)
    ap.add_argument("--proc-threads", type=int, default=2, help="proc threads per GPU (the reference default)")
    ap.add_argument("--keep", default=None)
    a = ap.parse_args()
    from PIL import Image
    work = a.keep or tempfile.mkdtemp(prefix="rsr_dir_")
    din, dout = os.path.join(work, "in"), os.path.join(work, "out")
    os.makedirs(din, exist_ok=True); os.makedirs(dout, exist_ok=True)
    imgs = []
    for i in range(a.images):  # seeds 1000..1099
        im = np.random.default_rng(1000 + i).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
        imgs.append(im)
        p = os.path.join(din, "%04d.png" % i)
        if not os.path.exists(p):
            Image.fromarray(im).save(p, "PNG", compress_level=1)
    mdir = os.path.join(ROOT, "models", "models-DF2K")
    out_mp = a.images * 1920 * 1080 * 16 / 1e6

    # (i) proc stage only: G engines, proc threads pulling pre-decoded images from one queue (main.cpp:311-331, :811-828)
    engines = []
    for g in range(a.gpus):
        e = rb.RealSR(g)
        assert e.load(os.path.join(mdir, "x4.param"), os.path.join(mdir, "x4.bin")) == 0, e.last_error()
        e.tilesize = a.tile
        engines.append(e)
    q = queue.Queue()
    outs = [[np.empty((4320, 7680, 3), np.uint8) for _ in range(a.proc_threads)] for _ in engines]

    def proc(e, buf):
        while True:
            im = q.get()
            if im is None:
                return
            assert e.process(im, buf) == 0, e.last_error()

    for e, bufs in zip(engines, outs):  # warm-up: one image per thread
        for b in bufs:
            assert e.process(imgs[0], b) == 0
    threads = [threading.Thread(target=proc, args=(e, b)) for e, bufs in zip(engines, outs) for b in bufs]
    for im in imgs:
        q.put(im)
    for _ in threads:
        q.put(None)
    t0 = time.time()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    t_proc = time.time() - t0
    del engines, outs

    # (ii) end to end through the CLI (decode -> proc -> encode), reference thread counts
    argv = ["-i", din, "-o", dout, "-m", mdir, "-t", ",".join([str(a.tile)] * a.gpus), "-g", ",".join(str(g) for g in range(a.gpus)),
            "-j", "4:%d:4" % 8 if a.gpus > 1 else "4:%d:4" % a.proc_threads]
    t0 = time.time()
    rc = cli.main(argv)
    t_e2e = time.time() - t0
    n_out = len(os.listdir(dout))
    # (iii) the same through the C++ command line (own PNG codec: Paeth + Z_RLE in parallel row bands)
    cpp = None
    exe = os.path.join(ROOT, "realsr-ncnn-vulkan_b200", "realsr-b200")
    if os.path.exists(exe):
        import subprocess
        dout2 = os.path.join(work, "out_cpp")
        os.makedirs(dout2, exist_ok=True)
        t0 = time.time()
        r = subprocess.run([exe] + argv[:2] + ["-o", dout2] + argv[4:], capture_output=True, text=True)
        t_cpp = time.time() - t0
        cpp = {"seconds": t_cpp, "out_mp_s": out_mp / t_cpp, "rc": r.returncode, "files_written": len(os.listdir(dout2)), "jobs": argv[-1],
               "codec": "host/png_codec.h"}
    print(json.dumps({"config": "configs[2]: %d x 1920x1080 PNG, models-DF2K, tile %d, %d GPU(s)" % (a.images, a.tile, a.gpus),
                      "proc_only": {"seconds": t_proc, "out_mp_s": out_mp / t_proc, "proc_threads_per_gpu": a.proc_threads},
                      "end_to_end": {"seconds": t_e2e, "out_mp_s": out_mp / t_e2e, "rc": rc, "files_written": n_out, "jobs": argv[-1],
                                     "codec": "Pillow PNG"},
                      "end_to_end_cpp": cpp}))
    if not a.keep:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
