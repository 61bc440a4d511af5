#!/usr/bin/env python3
"""Build oracle/_ref/: the reference's own CPU path (RealSR::process_cpu + vendored ncnn), compiled from the
sources where they lie under /root/reference.  TEST INFRASTRUCTURE ONLY (see oracle/README.md).

This does NOT run the reference's cmake build system.  It is a short hand-written recipe:
  * the three headers ncnn's cmake would generate from *.h.in templates (platform.h, layer_declaration.h,
    layer_registry.h, layer_type_enum.h, ncnn_export.h) are written here, with the same 15 layer types the
    reference build keeps (/root/reference/src/CMakeLists.txt:110-203) and Vulkan off;
  * the needed ncnn translation units are compiled with g++ directly (-O3 -fopenmp, one ISA, no runtime dispatch);
  * RealSR::process_cpu is taken verbatim at build time from /root/reference/src/realsr.cpp:525-838 into
    oracle/_ref/process_cpu.inc (git-ignored; never committed) and linked with oracle/ref_main.cpp.
Outputs go only to oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).

Usage: python oracle/build_ref.py [--isa avx2|avx512] [--jobs N]
"""
import argparse, os, re, subprocess, sys, shutil
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REALSR_REFERENCE", "/root/reference")
NCNN = os.path.join(REF, "src", "ncnn", "src")
OUT = os.path.join(HERE, "_ref")

KEEP = ["Concat", "Convolution", "Crop", "Eltwise", "Flatten", "InnerProduct", "Input", "ReLU", "Split",
        "BinaryOp", "Padding", "Interp", "Packing", "Cast", "PixelShuffle"]

CORE = ["allocator", "benchmark", "blob", "cpu", "datareader", "layer", "mat", "mat_pixel", "mat_pixel_resize",
        "modelbin", "net", "option", "paramdict", "command", "gpu", "pipeline", "pipelinecache"]

ISA_FLAGS = {
    "avx2": ["-mavx2", "-mfma", "-mf16c"],
    "avx512": ["-mavx512f", "-mavx512cd", "-mavx512bw", "-mavx512dq", "-mavx512vl", "-mavx2", "-mfma", "-mf16c"],
}


def layer_list():
    txt = open(os.path.join(NCNN, "CMakeLists.txt")).read()
    return re.findall(r"^ncnn_add_layer\((\w+)", txt, flags=re.M)


def gen_headers(gen, isa):
    os.makedirs(gen, exist_ok=True)
    defs = dict(NCNN_STDIO=1, NCNN_STRING=1, NCNN_THREADS=1, NCNN_PLATFORM_API=1, NCNN_PIXEL=1, NCNN_AVX=1,
                NCNN_FMA=1, NCNN_F16C=1, NCNN_AVX2=1, NCNN_AVX512=1 if isa == "avx512" else 0, NCNN_FORCE_INLINE=1)
    out = []
    for line in open(os.path.join(NCNN, "platform.h.in")):
        m = re.match(r"#cmakedefine01 (\w+)", line)
        if m:
            out.append("#define %s %d\n" % (m.group(1), defs.get(m.group(1), 0)))
        elif line.startswith("#cmakedefine NCNN_VERSION_STRING"):
            out.append('#define NCNN_VERSION_STRING "b4ba207-oracle"\n')
        else:
            out.append(line)
    open(os.path.join(gen, "platform.h"), "w").writelines(out)
    open(os.path.join(gen, "ncnn_export.h"), "w").write(
        "#ifndef NCNN_EXPORT_H\n#define NCNN_EXPORT_H\n#define NCNN_EXPORT\n#define NCNN_NO_EXPORT\n"
        "#define NCNN_DEPRECATED\n#endif\n")
    decl, reg, enum = [], [], []
    for i, cls in enumerate(layer_list()):
        name = cls.lower()
        enum.append("%s = %d," % (cls, i))
        if cls not in KEEP:
            reg.append('{"%s", 0},' % cls)
            continue
        bases = [cls]
        decl.append('#include "layer/%s.h"' % name)
        if os.path.exists(os.path.join(NCNN, "layer", "x86", name + "_x86.cpp")):
            decl.append('#include "layer/x86/%s_x86.h"' % name)
            bases.append(cls + "_x86")
        decl.append("namespace ncnn {\nclass %s_final : %s\n{\npublic:" %
                    (cls, ", ".join("virtual public " + b for b in bases)))
        decl.append("    virtual int create_pipeline(const Option& opt) {")
        for b in bases:
            decl.append("        { int ret = %s::create_pipeline(opt); if (ret) return ret; }" % b)
        decl.append("        return 0;\n    }\n    virtual int destroy_pipeline(const Option& opt) {")
        for b in reversed(bases):
            decl.append("        { int ret = %s::destroy_pipeline(opt); if (ret) return ret; }" % b)
        decl.append("        return 0;\n    }\n};\nDEFINE_LAYER_CREATOR(%s_final)\n} // namespace ncnn\n" % cls)
        reg.append('{"%s", %s_final_layer_creator},' % (cls, cls))
    open(os.path.join(gen, "layer_declaration.h"), "w").write("\n".join(decl) + "\n")
    open(os.path.join(gen, "layer_registry.h"), "w").write(
        "static const layer_registry_entry layer_registry[] = {\n" + "\n".join(reg) + "\n};\n")
    open(os.path.join(gen, "layer_type_enum.h"), "w").write("\n".join(enum) + "\n")
    open(os.path.join(gen, "layer_shader_type_enum.h"), "w").write("\n")  # Vulkan off: no shaders


def sources():
    srcs = [os.path.join(NCNN, c + ".cpp") for c in CORE]
    for cls in KEEP:
        n = cls.lower()
        srcs.append(os.path.join(NCNN, "layer", n + ".cpp"))
        x = os.path.join(NCNN, "layer", "x86", n + "_x86.cpp")
        if os.path.exists(x):
            srcs.append(x)
    return srcs


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-4000:] + r.stderr[-8000:] + "\n")
        raise SystemExit("build_ref: compile failed")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--isa", default="avx2", choices=sorted(ISA_FLAGS))
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    if not os.path.isdir(NCNN):
        raise SystemExit("build_ref: %s not found (the reference is only present in the build container)" % NCNN)
    exe = os.path.join(OUT, "oracle_ref_" + a.isa)
    if os.path.exists(exe) and not a.force:
        print("build_ref: %s already built" % exe)
    else:
        bld = os.path.join(OUT, "build_" + a.isa)
        gen = os.path.join(bld, "gen")
        gen_headers(gen, a.isa)
        cxx = ["g++", "-O3", "-fopenmp", "-fno-rtti", "-fno-exceptions", "-fvisibility=hidden", "-w",
               "-std=c++11", "-DNDEBUG"] + ISA_FLAGS[a.isa] + ["-I", gen, "-I", NCNN, "-I", os.path.join(NCNN, "layer"),
                                     "-I", os.path.join(NCNN, "layer", "x86")]
        objs = []
        jobs = []
        for s in sources():
            o = os.path.join(bld, os.path.relpath(s, NCNN).replace("/", "_")[:-4] + ".o")
            objs.append(o)
            jobs.append(cxx + ["-c", s, "-o", o])
        with ThreadPoolExecutor(a.jobs) as ex:
            list(ex.map(run, jobs))
        # RealSR::process_cpu verbatim from the reference, extracted at build time only
        lines = open(os.path.join(REF, "src", "realsr.cpp")).read().split("\n")
        body = "\n".join(lines[524:838]) + "\n"
        assert body.startswith("int RealSR::process_cpu"), body[:80]
        open(os.path.join(bld, "process_cpu.inc"), "w").write(body)
        run(cxx + ["-I", bld, "-I", os.path.join(REF, "src"), os.path.join(HERE, "ref_main.cpp")] + objs + ["-o", exe])
        print("build_ref: built", exe)
    # model files (data, not code) travel with the repo snapshot but stay out of git history
    mdst = os.path.join(os.path.dirname(HERE), "models")
    for m in ("models-DF2K", "models-DF2K_JPEG"):
        os.makedirs(os.path.join(mdst, m), exist_ok=True)
        for f in ("x4.param", "x4.bin"):
            d = os.path.join(mdst, m, f)
            if not os.path.exists(d):
                shutil.copyfile(os.path.join(REF, "models", m, f), d)


if __name__ == "__main__":
    main()
