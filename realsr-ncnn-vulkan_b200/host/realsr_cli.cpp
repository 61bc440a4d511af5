// realsr_cli.cpp -- the reference's command line (reference: /root/reference/src/main.cpp:101-115 usage, :422-879 main,
// :179-335 load / proc stages, :339-416 save stage) in C++ on top of `class RealSR` (host/realsr.h -> the C ABI).
//
//   realsr-b200 -i infile -o outfile [options]...      (same flags as realsr-ncnn-vulkan: -h -v -i -o -s -t -m -g -j -x -f)
//
// Same flag surface, validation messages, file planning and 3-stage thread pipeline (load threads -> queue of 8 -> proc
// threads, jobs_proc[i] per GPU sharing that GPU's RealSR object -> queue of 8 -> save threads; sentinel shutdown).
// Differences: `-g` takes CUDA ordinals and `-g -1` (ncnn CPU mode) is refused; one proc count is broadcast to all
// GPUs (BASELINE configs[2] uses `-j 4:8:4` with 8 GPUs, which the reference itself rejects, main.cpp:560-564); auto
// tile size (-t 0) is 200, the reference's pick on any GPU with more than 1.9 GB (main.cpp:753-774: tile geometry is part of the
// numerical contract); codecs: PNG (host/png_codec.h on zlib), JPEG (host/jpeg_codec.h: sequential + progressive reader,
// baseline writer) and WebP (host/webp_codec.h: lossless reader + writer; host/vp8_codec.h: lossy reader incl. alpha planes), all
// written for this repo -- the reference's stb / libwebp files are not copied.  Host plumbing only: every pixel goes through
// RealSR::process.
#include <dirent.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <new>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "jpeg_codec.h"
#include "png_codec.h"
#include "realsr.h"
#include "webp_codec.h"

static void print_usage()
{
    fprintf(stderr, "Usage: realsr-b200 -i infile -o outfile [options]...\n\n");
    fprintf(stderr, "  -h                   show this help\n");
    fprintf(stderr, "  -v                   verbose output\n");
    fprintf(stderr, "  -i input-path        input image path (jpg/png/webp) or directory\n");
    fprintf(stderr, "  -o output-path       output image path (jpg/png/webp) or directory\n");
    fprintf(stderr, "  -s scale             upscale ratio (4, default=4)\n");
    fprintf(stderr, "  -t tile-size         tile size (>=32/0=auto, default=0) can be 0,0,0 for multi-gpu\n");
    fprintf(stderr, "  -m model-path        realsr model path (default=models-DF2K_JPEG)\n");
    fprintf(stderr, "  -g gpu-id            gpu device to use (default=0) can be 0,1,2 for multi-gpu\n");
    fprintf(stderr, "  -j load:proc:save    thread count for load/proc/save (default=1:2:2) can be 1:2,2,2:2 for multi-gpu\n");
    fprintf(stderr, "  -x                   enable tta mode\n");
    fprintf(stderr, "  -f format            output image format (jpg/png/webp, default=ext/png)\n");
}

static bool is_dir(const std::string& p)
{
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
static std::string file_ext(const std::string& p)
{
    const size_t d = p.rfind('.'), s = p.rfind('/');
    return (d == std::string::npos || (s != std::string::npos && d < s)) ? "" : p.substr(d + 1);
}
static std::string name_noext(const std::string& n)
{
    const size_t d = n.rfind('.');
    return d == std::string::npos ? n : n.substr(0, d);
}
static std::vector<int> parse_int_list(const char* s, bool* ok)
{
    std::vector<int> v;
    const char* p = s;
    while (*p) {
        char* end;
        const long x = strtol(p, &end, 10);
        if (end == p) { *ok = false; return v; }
        v.push_back((int)x);
        p = *end == ',' ? end + 1 : end;
        if (*end && *end != ',') { *ok = false; return v; }
    }
    return v;
}

struct Task {
    int id = 0;
    std::string inpath, outpath;
    int w = 0, h = 0, c = 0;
    std::vector<unsigned char> in, out;
};

// bounded blocking queue (main.cpp:130-177: capacity 8)
class TaskQueue
{
public:
    void put(Task&& t)
    {
        std::unique_lock<std::mutex> lk(mu);
        not_full.wait(lk, [&] { return q.size() < 8; });
        q.push(std::move(t));
        not_empty.notify_one();
    }
    void get(Task& t)
    {
        std::unique_lock<std::mutex> lk(mu);
        not_empty.wait(lk, [&] { return !q.empty(); });
        t = std::move(q.front());
        q.pop();
        not_full.notify_one();
    }

private:
    std::mutex mu;
    std::condition_variable not_full, not_empty;
    std::queue<Task> q;
};

static TaskQueue toproc, tosave;
static const int END_ID = -233;  // the reference's sentinel id (main.cpp:320,:349)

int main(int argc, char** argv)
{
    std::string inputpath, outputpath, model = "models-DF2K_JPEG", format = "png";
    int scale = 4;
    std::vector<int> tilesize, gpuid, jobs_proc;
    int jobs_load = 1, jobs_save = 2, verbose = 0, tta_mode = 0;
    int opt;
    while ((opt = getopt(argc, argv, "i:o:s:t:m:g:j:f:vxh")) != -1) {
        bool ok = true;
        switch (opt) {
        case 'i': inputpath = optarg; break;
        case 'o': outputpath = optarg; break;
        case 's': scale = atoi(optarg); break;
        case 't': tilesize = parse_int_list(optarg, &ok); break;
        case 'm': model = optarg; break;
        case 'g': gpuid = parse_int_list(optarg, &ok); break;
        case 'j': {
            // "%d:%*[^:]:%d" plus the comma list in the middle (main.cpp:507-508)
            const char* c1 = strchr(optarg, ':');
            const char* c2 = c1 ? strchr(c1 + 1, ':') : nullptr;
            if (!c1 || !c2) { ok = false; break; }
            jobs_load = atoi(optarg);
            jobs_proc = parse_int_list(std::string(c1 + 1, c2).c_str(), &ok);
            jobs_save = atoi(c2 + 1);
            break;
        }
        case 'f': format = optarg; break;
        case 'v': verbose = 1; break;
        case 'x': tta_mode = 1; break;
        case 'h':
        default: print_usage(); return -1;
        }
        if (!ok) {
            fprintf(stderr, "invalid argument\n");
            return -1;
        }
    }
    if (inputpath.empty() || outputpath.empty()) {
        print_usage();
        return -1;
    }
    if (scale != 4) {
        fprintf(stderr, "invalid scale argument\n");
        return -1;
    }
    if (!tilesize.empty() && tilesize.size() != (gpuid.empty() ? 1 : gpuid.size())) {
        fprintf(stderr, "invalid tilesize argument\n");
        return -1;
    }
    for (int t : tilesize)
        if (t != 0 && t < 32) {
            fprintf(stderr, "invalid tilesize argument\n");
            return -1;
        }
    if (jobs_load < 1 || jobs_save < 1) {
        fprintf(stderr, "invalid thread count argument\n");
        return -1;
    }
    if (gpuid.empty()) gpuid.push_back(0);
    if (jobs_proc.size() == 1 && gpuid.size() > 1) jobs_proc.assign(gpuid.size(), jobs_proc[0]);  // broadcast (superset of the reference)
    if ((!jobs_proc.empty() && jobs_proc.size() != gpuid.size()) || std::any_of(jobs_proc.begin(), jobs_proc.end(), [](int j) { return j < 1; })) {
        fprintf(stderr, "invalid jobs_proc thread count argument\n");
        return -1;
    }
    if (jobs_proc.empty()) jobs_proc.assign(gpuid.size(), 2);
    if (tilesize.empty()) tilesize.assign(gpuid.size(), 0);
    if (!is_dir(outputpath)) {  // guess format from outputpath no matter what format argument specified (main.cpp:574-597)
        const std::string ext = file_ext(outputpath);
        if (ext == "png" || ext == "PNG") format = "png";
        else if (ext == "webp" || ext == "WEBP") format = "webp";
        else if (ext == "jpg" || ext == "JPG" || ext == "jpeg" || ext == "JPEG") format = "jpg";
        else {
            fprintf(stderr, "invalid outputpath extension type\n");
            return -1;
        }
    }
    if (format != "png" && format != "webp" && format != "jpg") {
        fprintf(stderr, "invalid format argument\n");
        return -1;
    }
    // collect input and output file paths (main.cpp:606-659)
    std::vector<std::string> input_files, output_files;
    if (is_dir(inputpath) && is_dir(outputpath)) {
        std::vector<std::string> names;
        if (DIR* d = opendir(inputpath.c_str())) {
            while (struct dirent* e = readdir(d)) {
                const std::string n = e->d_name;
                struct stat st;
                if (stat((inputpath + "/" + n).c_str(), &st) == 0 && S_ISREG(st.st_mode)) names.push_back(n);
            }
            closedir(d);
        }
        std::sort(names.begin(), names.end());
        std::string last, last_noext;
        bool have_last = false;
        for (const std::string& name : names) {
            const std::string noext = name_noext(name);
            std::string out = noext + "." + format;
            if (have_last && noext == last_noext) {  // filename collision: a.jpg and a.png both -> a.png
                const std::string out2 = name + "." + format;
                fprintf(stderr, "both %s and %s output %s ! %s will output %s\n", name.c_str(), last.c_str(), out.c_str(), name.c_str(), out2.c_str());
                out = out2;
            } else {
                last = name;
                last_noext = noext;
                have_last = true;
            }
            input_files.push_back(inputpath + "/" + name);
            output_files.push_back(outputpath + "/" + out);
        }
    } else if (!is_dir(inputpath) && !is_dir(outputpath)) {
        input_files.push_back(inputpath);
        output_files.push_back(outputpath);
    } else {
        fprintf(stderr, "inputpath and outputpath must be either file or directory at the same time\n");
        return -1;
    }
    if (model.find("models-DF2K") == std::string::npos) {
        fprintf(stderr, "unknown model dir type\n");
        return -1;
    }
    const int prepadding = 10;
    const std::string parampath = model + "/x4.param", modelpath = model + "/x4.bin";
    for (int g : gpuid)
        if (g < 0) {
            fprintf(stderr, "invalid gpu device (the B200 engine has no CPU mode)\n");
            return -1;
        }

    std::vector<RealSR*> realsr(gpuid.size(), nullptr);
    for (size_t i = 0; i < gpuid.size(); i++) {
        realsr[i] = new RealSR(gpuid[i], tta_mode != 0, 1);
        if (realsr[i]->load(parampath, modelpath)) {
            fprintf(stderr, "load failed: %s\n", realsr[i]->last_error());
            return -1;
        }
        realsr[i]->scale = scale;
        realsr[i]->tilesize = tilesize[i] ? tilesize[i] : 200;  // main.cpp:766-767
        realsr[i]->prepadding = prepadding;
    }

    // ---- load stage (main.cpp:179-306) ----
    const int n = (int)input_files.size();
    std::vector<std::thread> loaders, procs, savers;
    for (int k = 0; k < jobs_load; k++)
        loaders.emplace_back([&, k] {
            for (int i = k; i < n; i += jobs_load) {
                Task v;
                v.id = i;
                v.inpath = input_files[i];
                v.outpath = output_files[i];
                // by content, like the reference's decoders: PNG signature, else JPEG
                std::string why;
                unsigned char magic[4] = {0, 0, 0, 0};
                if (FILE* fp = fopen(v.inpath.c_str(), "rb")) {
                    if (fread(magic, 1, 4, fp) != 4) magic[0] = magic[1] = 0;
                    fclose(fp);
                }
                bool ok = false;
                try {
                    ok = (magic[0] == 0xFF && magic[1] == 0xD8) ? rsrhost::jpeg_decode_file(v.inpath, v.w, v.h, v.c, v.in, &why)
                         : !memcmp(magic, "RIFF", 4)            ? rsrhost::webp_decode_file(v.inpath, v.w, v.h, v.c, v.in, &why)
                                                                : rsrhost::png_decode_file(v.inpath, v.w, v.h, v.c, v.in, &why);
                    if (ok) v.out.resize((size_t)v.w * scale * v.h * scale * v.c);
                } catch (const std::bad_alloc&) {  // a header announcing more pixels than the host can hold
                    ok = false;
                    why = "out of memory";
                }
                if (!ok) {
                    fprintf(stderr, "decode image %s failed\n", v.inpath.c_str());
                    if (verbose && !why.empty()) fprintf(stderr, "  (%s)\n", why.c_str());
                    continue;
                }
                const std::string oext = file_ext(v.outpath);
                if (v.c == 4 && (oext == "jpg" || oext == "JPG" || oext == "jpeg" || oext == "JPEG")) {  // main.cpp:279-288
                    const std::string out2 = v.outpath + ".png";
                    fprintf(stderr, "image %s has alpha channel ! %s will output %s\n", v.inpath.c_str(), v.inpath.c_str(), out2.c_str());
                    v.outpath = out2;
                }
                toproc.put(std::move(v));
            }
        });
    // ---- proc stage (main.cpp:311-331): jobs_proc[i] threads share GPU i's RealSR object ----
    for (size_t g = 0; g < gpuid.size(); g++)
        for (int j = 0; j < jobs_proc[g]; j++)
            procs.emplace_back([&, g] {
                for (;;) {
                    Task v;
                    toproc.get(v);
                    if (v.id == END_ID) break;
                    ncnn::Mat inimage(v.w, v.h, (void*)v.in.data(), (size_t)v.c, v.c);
                    ncnn::Mat outimage(v.w * scale, v.h * scale, (void*)v.out.data(), (size_t)v.c, v.c);
                    const int rc = realsr[g]->process(inimage, outimage);
                    if (rc) {
                        fprintf(stderr, "process %s failed: %s\n", v.inpath.c_str(), realsr[g]->last_error());
                        continue;
                    }
                    tosave.put(std::move(v));
                }
            });
    // ---- save stage (main.cpp:339-416) ----
    // a 33-MP PNG costs ~1.4 core-seconds even in the fast mode: every save thread compresses its image in parallel row bands
    const int hw = (int)std::thread::hardware_concurrency();
    const int png_threads = std::max(1, std::min(8, (hw > 0 ? hw : 8) / std::max(1, jobs_save)));
    for (int k = 0; k < jobs_save; k++)
        savers.emplace_back([&] {
            for (;;) {
                Task v;
                tosave.get(v);
                if (v.id == END_ID) break;
                const std::string ext = file_ext(v.outpath);
                const bool ok = (ext == "png" || ext == "PNG")     ? rsrhost::png_encode_file(v.outpath, v.out.data(), v.w * scale, v.h * scale, v.c, 1, png_threads)
                                : (ext == "webp" || ext == "WEBP") ? rsrhost::webp_encode_file(v.outpath, v.out.data(), v.w * scale, v.h * scale, v.c)
                                                                   : rsrhost::jpeg_encode_file(v.outpath, v.out.data(), v.w * scale, v.h * scale, v.c, 100);
                if (ok) {
                    if (verbose) fprintf(stderr, "%s -> %s done\n", v.inpath.c_str(), v.outpath.c_str());
                } else {
                    fprintf(stderr, "encode image %s failed\n", v.outpath.c_str());
                }
            }
        });
    // sentinel shutdown, stage by stage (main.cpp:843-866)
    for (auto& t : loaders) t.join();
    for (size_t i = 0; i < procs.size(); i++) {
        Task end;
        end.id = END_ID;
        toproc.put(std::move(end));
    }
    for (auto& t : procs) t.join();
    for (size_t i = 0; i < savers.size(); i++) {
        Task end;
        end.id = END_ID;
        tosave.put(std::move(end));
    }
    for (auto& t : savers) t.join();
    for (RealSR* r : realsr) delete r;
    return 0;
}
