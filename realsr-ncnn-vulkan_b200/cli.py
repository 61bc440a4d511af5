#!/usr/bin/env python3
"""Command line of the reference (reference: /root/reference/src/main.cpp:101-115,:422-879) on top of the B200 engine.

    python -m realsr_ncnn_vulkan_b200.cli -i infile -o outfile [options]...

      -h                   show this help
      -v                   verbose output
      -i input-path        input image path (jpg/png/webp) or directory
      -o output-path       output image path (jpg/png/webp) or directory
      -s scale             upscale ratio (4, default=4)
      -t tile-size         tile size (>=32/0=auto, default=0) can be 0,0,0 for multi-gpu
      -m model-path        realsr model path (default=models-DF2K_JPEG)
      -g gpu-id            gpu device to use (default=0) can be 0,1,2 for multi-gpu
      -j load:proc:save    thread count for load/proc/save (default=1:2:2) can be 1:2,2,2:2 for multi-gpu
      -x                   enable tta mode
      -f format            output image format (jpg/png/webp, default=ext/png)

Same flag surface, validation messages and file naming as the reference; same 3-stage pipeline (load threads -> bounded
queue of 8 -> proc threads, `jobs_proc[i]` per GPU sharing one RealSR object -> bounded queue of 8 -> save threads,
main.cpp:130-177,:311-416,:802-866).  Differences: `-g` takes CUDA ordinals (`-g -1`, the ncnn CPU mode, is refused);
a single proc count is broadcast to all GPUs (the reference rejects `-j 4:8:4` with 8 GPUs, main.cpp:560-564; BASELINE
config 2 uses exactly that).  Auto tile size (-t 0) is 200: what the reference's heuristic picks on any GPU with more than
1.9 GB (main.cpp:753-774) -- tile geometry is part of the numerical contract, so default flags give the reference's image;
pass -t 400 for the larger tiles BASELINE config 1 names.  Codecs: Pillow (the reference vendors
stb_image / libwebp, main.cpp:243-288,:383-412); RGB and RGBA, grayscale is expanded to RGB like stbi_load(..., 3).
This file is host plumbing only: every pixel goes through RealSR.process (the C ABI)."""
import getopt
import os
import queue
import sys
import threading

import numpy as np

from .realsr import RealSR, RealSRError

# zlib level 1 with the run-length strategy (Pillow passes compress_type straight to deflateInit2: 3 = Z_RLE): a 7680x4320
# result is written in ~2 s instead of ~4 s at the same file size -- the save stage has to absorb ~9 such images per
# second and GPU (the C++ command line goes further: Paeth + Z_RLE in parallel row bands, host/png_codec.h)
PNG_FAST = dict(compress_level=1, compress_type=3)

USAGE = __doc__.split("\n\n")[1] + "\n\n" + __doc__.split("\n\n")[2]


def _print_usage():
    sys.stderr.write("Usage: realsr-ncnn-vulkan -i infile -o outfile [options]...\n\n" + __doc__.split("\n\n")[2] + "\n")


def parse_int_list(s):
    return [int(x) for x in s.split(",") if x != ""]


def parse_jobs(s):
    """'%d:%*[^:]:%d' plus the comma list in the middle (main.cpp:507-508)"""
    parts = s.split(":")
    if len(parts) != 3:
        raise ValueError("invalid thread count argument")
    return int(parts[0]), parse_int_list(parts[1]), int(parts[2])


def plan_files(inputpath, outputpath, fmt):
    """input/output path lists (main.cpp:606-659): dir->dir with sorted names and the collision rename, or file->file"""
    if os.path.isdir(inputpath) and os.path.isdir(outputpath):
        names = sorted(n for n in os.listdir(inputpath) if os.path.isfile(os.path.join(inputpath, n)))
        ins, outs = [], []
        last, last_noext = None, None
        for name in names:
            noext = os.path.splitext(name)[0]
            out = noext + "." + fmt
            if noext == last_noext:
                out2 = name + "." + fmt
                sys.stderr.write("both %s and %s output %s ! %s will output %s\n" % (name, last, out, name, out2))
                out = out2
            else:
                last, last_noext = name, noext
            ins.append(os.path.join(inputpath, name))
            outs.append(os.path.join(outputpath, out))
        return ins, outs
    if not os.path.isdir(inputpath) and not os.path.isdir(outputpath):
        return [inputpath], [outputpath]
    raise ValueError("inputpath and outputpath must be either file or directory at the same time")


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    inputpath = outputpath = None
    scale, tilesize, model, gpuid = 4, [], "models-DF2K_JPEG", []
    jobs_load, jobs_proc, jobs_save = 1, [], 2
    verbose = tta = False
    fmt = "png"
    try:
        opts, _ = getopt.getopt(argv, "i:o:s:t:m:g:j:f:vxh")
    except getopt.GetoptError:
        _print_usage()
        return -1
    try:
        for k, v in opts:
            if k == "-i": inputpath = v
            elif k == "-o": outputpath = v
            elif k == "-s": scale = int(v)
            elif k == "-t": tilesize = parse_int_list(v)
            elif k == "-m": model = v
            elif k == "-g": gpuid = parse_int_list(v)
            elif k == "-j": jobs_load, jobs_proc, jobs_save = parse_jobs(v)
            elif k == "-f": fmt = v
            elif k == "-v": verbose = True
            elif k == "-x": tta = True
            elif k == "-h":
                _print_usage()
                return -1
    except ValueError:
        sys.stderr.write("invalid argument\n")
        return -1
    if not inputpath or not outputpath:
        _print_usage()
        return -1
    if scale != 4:
        sys.stderr.write("invalid scale argument\n")
        return -1
    if tilesize and len(tilesize) != (len(gpuid) if gpuid else 1):
        sys.stderr.write("invalid tilesize argument\n")
        return -1
    if any(t != 0 and t < 32 for t in tilesize):
        sys.stderr.write("invalid tilesize argument\n")
        return -1
    if jobs_load < 1 or jobs_save < 1:
        sys.stderr.write("invalid thread count argument\n")
        return -1
    if not gpuid:
        gpuid = [0]
    if len(jobs_proc) == 1 and len(gpuid) > 1:
        jobs_proc = jobs_proc * len(gpuid)  # superset of the reference: broadcast one proc count to every GPU
    if jobs_proc and len(jobs_proc) != len(gpuid) or any(j < 1 for j in jobs_proc):
        sys.stderr.write("invalid jobs_proc thread count argument\n")
        return -1
    if not jobs_proc:
        jobs_proc = [2] * len(gpuid)
    if not tilesize:
        tilesize = [0] * len(gpuid)
    if not os.path.isdir(outputpath):  # guess format from outputpath no matter what format argument specified (main.cpp:574-597)
        ext = os.path.splitext(outputpath)[1].lstrip(".")
        if ext in ("png", "PNG"):
            fmt = "png"
        elif ext in ("webp", "WEBP"):
            fmt = "webp"
        elif ext in ("jpg", "JPG", "jpeg", "JPEG"):
            fmt = "jpg"
        else:
            sys.stderr.write("invalid outputpath extension type\n")
            return -1
    if fmt not in ("png", "webp", "jpg"):
        sys.stderr.write("invalid format argument\n")
        return -1
    try:
        input_files, output_files = plan_files(inputpath, outputpath, fmt)
    except ValueError as e:
        sys.stderr.write(str(e) + "\n")
        return -1
    if "models-DF2K" not in model:
        sys.stderr.write("unknown model dir type\n")
        return -1
    prepadding = 10
    param, binf = os.path.join(model, "x4.param"), os.path.join(model, "x4.bin")
    if any(g < 0 for g in gpuid):
        sys.stderr.write("invalid gpu device (the B200 engine has no CPU mode)\n")
        return -1

    engines = []
    for g, t in zip(gpuid, tilesize):
        e = RealSR(g, tta, 1)
        if e.load(param, binf):
            sys.stderr.write("load failed: %s\n" % e.last_error())
            return -1
        e.scale, e.tilesize, e.prepadding = scale, (t if t else 200), prepadding
        engines.append(e)

    from PIL import Image
    toproc, tosave = queue.Queue(maxsize=8), queue.Queue(maxsize=8)
    END = object()
    failed = []

    def load_worker(indices):
        for i in indices:
            try:
                im = Image.open(input_files[i])
                im = im.convert("RGBA" if (im.mode in ("RGBA", "LA") or "transparency" in im.info) else "RGB")
                toproc.put((i, np.ascontiguousarray(np.asarray(im, dtype=np.uint8))))
            except Exception:
                sys.stderr.write("decode image %s failed\n" % input_files[i])
                failed.append(i)

    def proc_worker(engine):
        while True:
            item = toproc.get()
            if item is END:
                return
            i, img = item
            try:  # any failure marks the image failed and keeps the worker alive (a dead worker would wedge the queues)
                out = np.empty((img.shape[0] * scale, img.shape[1] * scale, img.shape[2]), dtype=np.uint8)
                rc = engine.process(img, out)
            except Exception as ex:
                rc = -1
                sys.stderr.write(str(ex) + "\n")
            if rc:
                sys.stderr.write("process %s failed: %s\n" % (input_files[i], engine.last_error()))
                failed.append(i)
                continue
            tosave.put((i, out))

    def save_worker():
        while True:
            item = tosave.get()
            if item is END:
                return
            i, out = item
            path = output_files[i]
            try:
                im = Image.fromarray(out)
                if fmt == "jpg":
                    if out.shape[2] == 4:  # the reference appends ".png" to the jpg name of alpha outputs (main.cpp:279-288)
                        path = path + ".png"
                        sys.stderr.write("image %s has alpha channel ! %s will output %s\n" % (input_files[i], input_files[i], path))
                        im.save(path, "PNG", **PNG_FAST)
                    else:
                        im.save(path, "JPEG", quality=100)
                elif fmt == "webp":
                    im.save(path, "WEBP", lossless=True)
                else:
                    im.save(path, "PNG", **PNG_FAST)
                if verbose:
                    sys.stderr.write("%s -> %s done\n" % (input_files[i], path))
            except Exception:
                sys.stderr.write("encode image %s failed\n" % path)
                failed.append(i)

    n = len(input_files)
    loaders = [threading.Thread(target=load_worker, args=(range(k, n, jobs_load),)) for k in range(jobs_load)]
    procs = [threading.Thread(target=proc_worker, args=(e,)) for e, j in zip(engines, jobs_proc) for _ in range(j)]
    savers = [threading.Thread(target=save_worker) for _ in range(jobs_save)]
    for t in loaders + procs + savers:
        t.start()
    for t in loaders:
        t.join()
    for _ in procs:
        toproc.put(END)
    for t in procs:
        t.join()
    for _ in savers:
        tosave.put(END)
    for t in savers:
        t.join()
    return 0  # like the reference (main.cpp:866-879): failures are reported per image on stderr, the exit status stays 0


if __name__ == "__main__":
    sys.exit(main())
