// xmipp_phantom_project_b200 — the Fourier mode of xmipp_phantom_project (libraries/reconstruction/project.cpp:35-86,
// 89-140, 992: FourierProjector over the requested orientations) on the GPU projector of librecfourier_b200.so.
//
//   xmipp_phantom_project_b200 -i volume.vol -o image.xmp --angles <rot> <tilt> <psi> [--method fourier <pad=2> <maxfreq=0.25> <interp=bspline>]
//   xmipp_phantom_project_b200 -i volume.vol -o stack.stk --angles_md angles.xmd [...]      (extension: one projection per row of
//                              angleRot / angleTilt / anglePsi; writes the stack and <stack root>.xmd)
//
// Only --method fourier is implemented (the reference defaults to real_space ray tracing; asking for it, or for --params /
// shifted single projections, is refused with a message).  Like the reconstruction program the CUDA library is bound at
// run time; there is no CPU fallback.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "image_io.h"
#include "metadata.h"
#include "recfourier_b200.h"

using namespace rfhost;

namespace {

struct Api {
    int (*create)(const float*, int32_t, double, double, int32_t, int32_t, rfb200_projector*) = nullptr;
    int (*project)(rfb200_projector, const double*, const float*, int32_t, float*) = nullptr;
    const char* (*last_error)(rfb200_projector) = nullptr;
    void (*destroy)(rfb200_projector) = nullptr;
};

std::string selfDir() {
    Dl_info info;
    if (dladdr((void*)&selfDir, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t s = p.rfind('/');
        if (s != std::string::npos) return p.substr(0, s);
    }
    return ".";
}

Api loadApi() {
    Api a;
    const char* env = getenv("RFB200_LIB");
    std::vector<std::string> names;
    if (env && *env) names.push_back(env);
    names.push_back(selfDir() + "/librecfourier_b200.so");
    names.push_back("librecfourier_b200.so");
    void* lib = nullptr;
    std::string tried;
    for (auto& n : names) {
        lib = dlopen(n.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (lib) break;
        const char* e = dlerror();
        tried += "\n  " + n + ": " + (e ? e : "?");
    }
    if (!lib) throw std::runtime_error("cannot load the CUDA library librecfourier_b200.so (there is no CPU fallback):" + tried);
    a.create = (decltype(a.create))dlsym(lib, "rfb200_projector_create");
    a.project = (decltype(a.project))dlsym(lib, "rfb200_projector_project");
    a.last_error = (decltype(a.last_error))dlsym(lib, "rfb200_projector_last_error");
    a.destroy = (decltype(a.destroy))dlsym(lib, "rfb200_projector_destroy");
    if (!a.create || !a.project || !a.last_error || !a.destroy) throw std::runtime_error("librecfourier_b200.so lacks the projector symbols");
    return a;
}

const char* kUsage =
    "Generate projections from a volume (Fourier central-slice method on the GPU).\n"
    "   -i <volume_file>                           : Voxel volume (Spider / MRC), cubic\n"
    "   -o <image_file>                            : Output image or stack\n"
    "  [--method fourier <pad=2> <maxfreq=0.25> <interp=bspline>] : interp: nearest, linear, bspline (only method implemented)\n"
    "  [--angles <rot> <tilt> <psi>]               : Angles for a single projection\n"
    "  [--angles_md <md_file>]                     : one projection per row (angleRot, angleTilt, anglePsi); writes a stack\n"
    "                                                and <output root>.xmd (extension of this implementation)\n"
    "  [--device <dev=0>]                          : GPU device\n";

bool isOpt(const std::string& t) { return t.size() > 1 && t[0] == '-' && !(isdigit((unsigned char)t[1]) || t[1] == '.'); }

}  // namespace

int main(int argc, char** argv) {
    try {
        std::vector<std::string> v(argv + 1, argv + argc);
        auto values = [&](const std::string& name, bool& present) {
            std::vector<std::string> out;
            present = false;
            for (size_t i = 0; i < v.size(); ++i)
                if (v[i] == name) {
                    present = true;
                    for (size_t k = i + 1; k < v.size() && !isOpt(v[k]); ++k) out.push_back(v[k]);
                    break;
                }
            return out;
        };
        static const char* known[] = {"-i", "-o", "--method", "--angles", "--angles_md", "--device", "--params", "--sym", "--xdim",
                                      "--sampling_rate", "--high_sampling_rate", "--only_create_angles", "-h", "--help", "-v"};
        for (auto& t : v) {
            if (!isOpt(t)) continue;
            bool ok = false;
            for (auto k : known) ok = ok || t == k;
            if (!ok) throw std::invalid_argument("unknown option " + t + "\n" + kUsage);
        }
        bool has;
        values("-h", has);
        bool help = has;
        values("--help", has);
        if (help || has || v.empty()) {
            std::cerr << kUsage;
            return 2;
        }
        auto fnIn = values("-i", has);
        if (!has || fnIn.empty()) throw std::invalid_argument(std::string("-i <volume_file> is required\n") + kUsage);
        auto fnOut = values("-o", has);
        if (!has || fnOut.empty()) throw std::invalid_argument(std::string("-o <image_file> is required\n") + kUsage);
        values("--params", has);
        if (has) throw std::invalid_argument("--params (projection parameter files) is not implemented; use --angles or --angles_md");
        double pad = 2.0, maxFreq = 0.25;
        int degree = 3;
        auto method = values("--method", has);
        if (has) {
            if (method.empty() || method[0] != "fourier")
                throw std::invalid_argument("only --method fourier is implemented (real_space and shears are not)");
            if (method.size() > 1) pad = atof(method[1].c_str());
            if (method.size() > 2) maxFreq = atof(method[2].c_str());
            if (method.size() > 3) {
                if (method[3] == "nearest") degree = 0;
                else if (method[3] == "linear") degree = 1;
                else if (method[3] == "bspline") degree = 3;
                else throw std::invalid_argument("The values for interpolation can be : nearest, linear, bspline");   // project.cpp:59
            }
        } else
            throw std::invalid_argument("the reference's default --method real_space is not implemented: pass --method fourier [pad maxfreq interp]");
        auto ang = values("--angles", has);
        const bool doAngles = has;
        auto angMd = values("--angles_md", has);
        const bool doMd = has;
        if (doAngles == doMd) throw std::invalid_argument("give exactly one of --angles <rot> <tilt> <psi> and --angles_md <md_file>");
        int device = 0;
        auto dv = values("--device", has);
        if (has && !dv.empty()) device = atoi(dv[0].c_str());

        std::vector<double> angles;
        MetaData md;
        if (doAngles) {
            if (ang.size() < 3) throw std::invalid_argument("--angles needs <rot> <tilt> <psi>");
            if (ang.size() > 3 && (atof(ang[3].c_str()) != 0.0 || (ang.size() > 4 && atof(ang[4].c_str()) != 0.0)))
                throw std::invalid_argument("shifted single projections (--angles ... <x> <y>) are not implemented");
            for (int k = 0; k < 3; ++k) angles.push_back(atof(ang[k].c_str()));
        } else {
            if (angMd.empty()) throw std::invalid_argument("--angles_md needs a metadata file");
            md.read(angMd[0]);
            md.removeDisabled();
            if (md.size() == 0) throw std::runtime_error("no rows in " + angMd[0]);
            for (size_t i = 0; i < md.size(); ++i) {
                angles.push_back(md.getValueOrDefault("angleRot", i, 0));
                angles.push_back(md.getValueOrDefault("angleTilt", i, 0));
                angles.push_back(md.getValueOrDefault("anglePsi", i, 0));
            }
        }
        // ---- the volume: nz slices of a cubic object
        ImageInfo info = readImageInfo(fnIn[0]);
        const int N = info.nx;
        if (info.ny != N || info.nz != N) throw std::runtime_error("the volume must be cubic (" + std::to_string(info.nx) + "x" + std::to_string(info.ny) + "x" + std::to_string(info.nz) + ")");
        std::vector<float> vol((size_t)N * N * N);
        for (int k = 0; k < N; ++k) {
            char idx[32];
            snprintf(idx, sizeof idx, "%d@", k + 1);
            readImage2D(idx + fnIn[0], &vol[(size_t)k * N * N], N, N);
        }
        closeImageCache();
        Api api = loadApi();
        rfb200_projector pr = nullptr;
        int rc = api.create(vol.data(), N, pad, maxFreq, degree, device, &pr);
        if (rc != RFB200_OK) throw std::runtime_error(rc == RFB200_ERR_CUDA ? "GPU initialisation failed (no CUDA device? there is no CPU fallback)" : "bad projector parameters");
        const size_t n = angles.size() / 3;
        std::vector<float> img(n * (size_t)N * N);
        rc = api.project(pr, angles.data(), nullptr, (int32_t)n, img.data());
        if (rc != RFB200_OK) {
            std::string msg = api.last_error(pr);
            api.destroy(pr);
            throw std::runtime_error("projection failed: " + msg);
        }
        api.destroy(pr);
        if (doAngles) {
            writeVolume(fnOut[0], img.data(), N, N, 1);
        } else {
            writeStack(fnOut[0], img.data(), N, N, n);
            MetaData out;
            for (const char* l : {"image", "enabled", "angleRot", "angleTilt", "anglePsi", "shiftX", "shiftY"}) out.addLabel(l);
            size_t slash = fnOut[0].rfind('/');
            const std::string base = slash == std::string::npos ? fnOut[0] : fnOut[0].substr(slash + 1);
            for (size_t i = 0; i < n; ++i) {
                size_t r = out.addRow();
                char name[64];
                snprintf(name, sizeof name, "%06zu@", i + 1);
                out.setValue("image", r, name + base);
                out.setValue("enabled", r, 1.0);
                out.setValue("angleRot", r, angles[3 * i]);
                out.setValue("angleTilt", r, angles[3 * i + 1]);
                out.setValue("anglePsi", r, angles[3 * i + 2]);
                out.setValue("shiftX", r, 0.0);
                out.setValue("shiftY", r, 0.0);
            }
            size_t dot = fnOut[0].rfind('.');
            out.write((dot == std::string::npos || (slash != std::string::npos && dot < slash) ? fnOut[0] : fnOut[0].substr(0, dot)) + ".xmd");
        }
        return 0;
    } catch (const std::invalid_argument& e) {
        std::cerr << e.what() << std::endl;
        return 2;
    } catch (const std::exception& e) {
        std::cerr << "XMIPP_ERROR: " << e.what() << std::endl;
        return 1;
    }
}
