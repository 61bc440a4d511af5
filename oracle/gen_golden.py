#!/usr/bin/env python3
"""Generate tests/golden/* with the REAL reference CPU path (oracle/_ref/oracle_ref_avx2 = the reference's own
RealSR::process_cpu + vendored ncnn b4ba207, AVX2 kernels, see oracle/build_ref.py).  Run in the build container
(needs /root/reference for build_ref.py and images/0.png); the fixtures are committed because the GPU box has
no /root/reference.  TEST INFRASTRUCTURE ONLY.

Each case = <name>.in.npy (u8 HWC), <name>.out.npy (u8 HWC, 4x) and an entry in manifest.json."""
import json, os, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref", "oracle_ref_avx2")
REF512 = os.path.join(ROOT, "oracle", "_ref", "oracle_ref_avx512")  # what bench.py's CPU arm runs on an AVX-512 host


def noise(h, w, seed):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def natural(h, w, seed):
    """box-blurred noise (SURVEY 8d 'natural' variant): smooth content, full u8 range"""
    r = np.random.default_rng(seed).integers(0, 256, (h + 8, w + 8, 3)).astype(np.float64)
    c = np.cumsum(np.cumsum(np.pad(r, ((1, 0), (1, 0), (0, 0))), 0), 1)
    k = 7
    b = (c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]) / (k * k)
    b = b[:h, :w]
    b = (b - b.min()) / (b.max() - b.min()) * 255.0
    return np.clip(np.rint(b), 0, 255).astype(np.uint8)


def run_ref(img, model, tilesize, tta, threads=8, exe=None):
    exe = exe or REF
    h, w, c = img.shape
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, "i.raw"), os.path.join(td, "o.raw")
        img.tofile(fi)
        mdir = os.path.join(ROOT, "models", model)
        r = subprocess.run([exe, os.path.join(mdir, "x4.param"), os.path.join(mdir, "x4.bin"), fi, str(w), str(h),
                            str(c), fo, str(tilesize), str(int(tta)), str(threads)], capture_output=True, text=True)
        if r.returncode:
            raise SystemExit(r.stderr)
        info = json.loads(r.stdout.strip().splitlines()[-1])
        return np.fromfile(fo, dtype=np.uint8).reshape(h * 4, w * 4, c), info


def main():
    os.makedirs(GOLD, exist_ok=True)
    cases = [
        # name, image, model, tilesize, tta
        ("noise_40x36_jpeg_t200", noise(36, 40, 1001), "models-DF2K_JPEG", 200, 0),
        ("natural_48x40_df2k_t32", natural(40, 48, 1002), "models-DF2K", 32, 0),
        ("noise_50x33_df2k_t32", noise(33, 50, 1003), "models-DF2K", 32, 0),
        ("natural_28x20_df2k_tta", natural(20, 28, 1004), "models-DF2K", 200, 1),
        ("noise_37x23_jpeg_tta_t32", noise(23, 37, 1005), "models-DF2K_JPEG", 32, 1),
        ("cfg1_noise_256_jpeg_t200", noise(256, 256, 1000), "models-DF2K_JPEG", 200, 0),
        ("cfg1_noise_256_jpeg_t256", noise(256, 256, 1000), "models-DF2K_JPEG", 256, 0),
    ]
    try:
        from PIL import Image
        im = np.asarray(Image.open("/root/reference/images/0.png").convert("RGB"))
        cases.append(("readme0_crop64_df2k_t200", np.ascontiguousarray(im[70:134, 60:124]), "models-DF2K", 200, 0))
        cases.append(("readme0_crop96x80_df2k_t64", np.ascontiguousarray(im[40:120, 30:126]), "models-DF2K", 64, 0))
    except Exception as e:  # PIL or the image absent
        print("skipping README image cases:", e)
    # RGBA (single-tile only: the reference's CPU branch is not self-consistent for multi-tile RGBA, see realsr_oracle.c)
    def rgba(img, seed):
        h, w, _ = img.shape
        r = np.random.default_rng(seed)
        yy, xx = np.mgrid[0:h, 0:w]
        a = np.clip((xx * 6 + yy * 3) % 256 + r.integers(-20, 20, (h, w)), 0, 255).astype(np.uint8)
        a[: h // 4, : w // 3] = 255
        a[h // 2:, w // 2:] = 0
        return np.ascontiguousarray(np.dstack([img, a]))
    cases.append(("rgba_noise_41x30_df2k_t200", rgba(noise(30, 41, 1006), 6), "models-DF2K", 200, 0))
    cases.append(("rgba_natural_36x28_jpeg_t200", rgba(natural(28, 36, 1007), 7), "models-DF2K_JPEG", 200, 0))
    # smallest sizes for which the reference is defined: with w or h <= prepadding (10) ncnn's reflect border
    # (copy_make_border, called at realsr.cpp:613,:764) reads outside the image -- the oracle port and the reference then
    # disagree arbitrarily (probed: 10x10, 1x1), the engine clamps; 11 px is the first well-defined size
    cases.append(("noise_12x11_df2k_tta_t32", noise(11, 12, 1008), "models-DF2K", 32, 1))
    cases.append(("noise_45x11_jpeg_t32", noise(11, 45, 1009), "models-DF2K_JPEG", 32, 0))
    # shapes the benchmark times: one 420x300 padded tile of BASELINE configs[1] (tile 400), and a TTA
    # 100x140 edge tile of configs[4] (tile 200; variants 4..7 run it transposed)
    cases.append(("cfg2_tile_420x300_natural_df2k_t400", natural(280, 400, 1010), "models-DF2K", 400, 0))
    cases.append(("cfg5_tta_edge_100x140_natural_df2k_t200", natural(120, 80, 1011), "models-DF2K", 200, 1))
    only_new = "--only-new" in sys.argv
    manifest = {"generator": "oracle/gen_golden.py", "reference_binary": "oracle/_ref/oracle_ref_avx2 (ncnn b4ba207, AVX2+FMA, OpenMP)",
                "prepadding": 10, "cases": []}
    flags = [l for l in open("/proc/cpuinfo") if l.startswith("model name")][:1]
    manifest["host_cpu"] = flags[0].split(":", 1)[1].strip() if flags else "?"
    old = {}
    if only_new and os.path.exists(os.path.join(GOLD, "manifest.json")):
        old = {c["name"]: c for c in json.load(open(os.path.join(GOLD, "manifest.json")))["cases"]}
    for name, img, model, ts, tta in cases:
        if name in old:
            manifest["cases"].append(old[name])
            continue
        out, info = run_ref(img, model, ts, tta)
        np.save(os.path.join(GOLD, name + ".in.npy"), img)
        if out.size > (4 << 20):  # large outputs are stored compressed (tests/common.py golden() reads either form)
            np.savez_compressed(os.path.join(GOLD, name + ".out.npz"), out=out)
        else:
            np.save(os.path.join(GOLD, name + ".out.npy"), out)
        manifest["cases"].append({"name": name, "model": model, "tilesize": ts, "tta": tta, "h": img.shape[0],
                                  "w": img.shape[1], "c": int(img.shape[2]), "ref_seconds": info["seconds"]})
        print(name, info["seconds"], "s")
    # ISA cross-check (ncnn's conv kernels, hence the bit pattern, depend on the ISA): the goldens are the AVX2
    # build's outputs; the AVX-512 build (pack16 Winograd) is run on every case and its byte differences are recorded, so the
    # contract "vs the reference CPU path" is pinned for both builds
    if "--cross-isa" in sys.argv and os.path.exists(REF512) and " avx512f" in open("/proc/cpuinfo").read():
        byname = {c[0]: c for c in cases}
        for c in manifest["cases"]:
            if "avx512_vs_avx2" in c and only_new:
                continue
            name, img, model, ts, tta = byname[c["name"]]
            f = os.path.join(GOLD, name + ".out.npy")
            gold = np.load(f) if os.path.exists(f) else np.load(os.path.join(GOLD, name + ".out.npz"))["out"]
            out512, _ = run_ref(img, model, ts, tta, exe=REF512)
            d = np.abs(out512.astype(np.int16) - gold.astype(np.int16))
            c["avx512_vs_avx2"] = {"max": int(d.max()), "bytes_differ": int((d > 0).sum()), "bytes": int(d.size)}
            print(name, "avx512 vs avx2:", c["avx512_vs_avx2"])
        manifest["isa_note"] = ("goldens = AVX2 build; avx512_vs_avx2 = byte differences of the AVX-512 build (ncnn pack16 Winograd) on the same "
                                "input, generated on " + manifest["host_cpu"])
    json.dump(manifest, open(os.path.join(GOLD, "manifest.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
