// xmipp_phantom_project_b200 — the Fourier mode of xmipp_phantom_project (libraries/reconstruction/project.cpp:35-86,
// 89-140, 992: FourierProjector over the requested orientations) on the GPU projector of librecfourier_b200.so.
//
//   xmipp_phantom_project_b200 -i volume.vol -o image.xmp --angles <rot> <tilt> <psi> [--method fourier <pad=2> <maxfreq=0.25> <interp=bspline>]
//   xmipp_phantom_project_b200 -i volume.vol -o stack.stk --angles_md angles.xmd [...]      (extension: one projection per row of
//                              angleRot / angleTilt / anglePsi; writes the stack and <stack root>.xmd)
//
//   xmipp_phantom_project_b200 -i volume.vol -o stack.stk --params proj.param [...]         (project.cpp:62-72, 220-388: metadata
//                              flavour of the parameter file, angle file or deterministic ranges)
//
// Only --method fourier is implemented (the reference defaults to real_space ray tracing; asking for it, for random angle
// ranges / noise in --params, or for shifted projections, is refused with a message).  Like the reconstruction program the CUDA library is bound at
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
    "  [--params <parameters_file>]                : projection parameter file (metadata flavour): _dimensions2D, and\n"
    "                                                _projAngleFile or deterministic _projRotRange / _projTiltRange /\n"
    "                                                _projPsiRange; writes a stack and <output root>.xmd\n"
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
        auto fnParams = values("--params", has);
        const bool doParams = has;
        if (doParams && fnParams.empty()) throw std::invalid_argument("--params needs a parameter file");
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
        if ((int)doAngles + (int)doMd + (int)doParams != 1)
            throw std::invalid_argument(doAngles && doParams ? "--params and --angles are mutually exclusive"        // project.cpp:65-66
                                                             : "You should provide --params or --angles (or --angles_md)");
        int device = 0;
        auto dv = values("--device", has);
        if (has && !dv.empty()) device = atoi(dv[0].c_str());

        std::vector<double> angles;
        MetaData md;
        int paramXdim = -1, paramYdim = -1;
        if (doAngles) {
            if (ang.size() < 3) throw std::invalid_argument("--angles needs <rot> <tilt> <psi>");
            if (ang.size() > 3 && (atof(ang[3].c_str()) != 0.0 || (ang.size() > 4 && atof(ang[4].c_str()) != 0.0)))
                throw std::invalid_argument("shifted single projections (--angles ... <x> <y>) are not implemented");
            for (int k = 0; k < 3; ++k) angles.push_back(atof(ang[k].c_str()));
        } else if (doParams) {
            // ParametersProjection::read, metadata flavour (project.cpp:220-388): one non-loop block with _dimensions2D,
            // then either _projAngleFile (a metadata with the angles) or the three deterministic ranges
            // _projRotRange / _projTiltRange / _projPsiRange '<ang0> [<angF> <samples>]'; projection number
            // (i_rot * Ntilt + i_tilt) * Npsi + i_psi (generate_angles, project.cpp:560-703).  Random ranges, angular and pixel
            // noise need the reference's random stream and are refused unless they are zero / absent.
            MetaData pm;
            pm.read(fnParams[0]);
            if (pm.size() == 0) throw std::runtime_error("Prog_Project_Parameters::read: There is a problem opening the file " + fnParams[0]);
            auto vec = [&](const char* label) {
                std::vector<double> out;
                std::string sv;
                if (pm.getValue(label, 0, sv)) {
                    char* c = const_cast<char*>(sv.c_str());
                    for (;;) {
                        char* end = nullptr;
                        double x = strtod(c, &end);
                        if (end == c) break;
                        out.push_back(x);
                        c = end;
                    }
                }
                return out;
            };
            for (const char* l : {"noisePixelLevel", "noiseCoord", "projRotNoise", "projTiltNoise", "projPsiNoise"}) {
                auto nv = vec(l);
                for (double x : nv)
                    if (x != 0.0) throw std::runtime_error(std::string("--params: _") + l + " with a non-zero value (random noise) is not implemented");
            }
            for (const char* l : {"projRotRandomness", "projTiltRandomness", "projPsiRandomness"}) {
                std::string rs;
                if (pm.getValue(l, 0, rs) && !rs.empty() && rs != "NULL" && rs != "even")
                    throw std::runtime_error(std::string("--params: _") + l + " (random angle ranges) is not implemented");
            }
            auto dims = vec("dimensions2D");
            if (dims.size() >= 2) paramXdim = (int)dims[0], paramYdim = (int)dims[1];
            std::string fnAng;
            if (pm.getValue("projAngleFile", 0, fnAng) && !fnAng.empty() && fnAng != "NULL") {
                size_t slash = fnParams[0].rfind('/');
                FILE* probe = fopen(fnAng.c_str(), "rb");
                if (probe) fclose(probe);
                else if (fnAng[0] != '/' && slash != std::string::npos) fnAng = fnParams[0].substr(0, slash + 1) + fnAng;
                md.read(fnAng);
                md.removeDisabled();
                if (md.size() == 0) throw std::runtime_error("Prog_Project_Parameters::read: file " + fnAng + " doesn't exist");
                for (size_t i = 0; i < md.size(); ++i) {
                    if (md.getValueOrDefault("shiftX", i, 0) != 0.0 || md.getValueOrDefault("shiftY", i, 0) != 0.0)
                        throw std::runtime_error("--params: shifted projections (shiftX / shiftY in the angle file) are not implemented");
                    angles.push_back(md.getValueOrDefault("angleRot", i, 0));
                    angles.push_back(md.getValueOrDefault("angleTilt", i, 0));
                    angles.push_back(md.getValueOrDefault("anglePsi", i, 0));
                }
            } else {
                struct Range { double a0 = 0, aF = 0; int n = 1; } rg[3];
                const char* labels[3] = {"projRotRange", "projTiltRange", "projPsiRange"};
                for (int k = 0; k < 3; ++k) {
                    auto rv = vec(labels[k]);
                    if (rv.empty()) throw std::runtime_error(std::string("--params: neither _projAngleFile nor _") + labels[k] + " given");
                    rg[k].a0 = rv[0];
                    if (rv.size() >= 3) { rg[k].aF = rv[1]; rg[k].n = (int)rv[2]; } else { rg[k].aF = rv[0]; rg[k].n = 1; }
                    if (rg[k].a0 == rg[k].aF) rg[k].n = 1;                                  // project.cpp:274-275
                    if (rg[k].n < 1) throw std::runtime_error(std::string("--params: bad sample count in _") + labels[k]);
                }
                auto at = [](const Range& r, int i) { return r.n > 1 ? r.a0 + (r.aF - r.a0) / (double)(r.n - 1) * i : r.a0; };
                for (int i = 0; i < rg[0].n; ++i)
                    for (int j = 0; j < rg[1].n; ++j)
                        for (int k = 0; k < rg[2].n; ++k) {
                            angles.push_back(at(rg[0], i));
                            angles.push_back(at(rg[1], j));
                            angles.push_back(at(rg[2], k));
                        }
            }
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
        if (paramXdim > 0 && (paramXdim != N || paramYdim != N))
            throw std::runtime_error("--params: _dimensions2D " + std::to_string(paramXdim) + " " + std::to_string(paramYdim) +
                                     " differs from the volume size " + std::to_string(N) + " (resizing projections is not implemented)");
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
