// image_io.cpp — see image_io.h
#include "image_io.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <vector>

namespace rfhost {

enum class ImageKind { SPIDER, MRC };

// An open image file: header decoded once, descriptor owned (closed by the last holder: the cache entry or a reader in
// flight).
struct ImageSource {
    ImageSource() = default;
    ImageSource(const ImageSource&) = delete;
    ImageSource& operator=(const ImageSource&) = delete;
    ~ImageSource() { if (fd >= 0) close(fd); }
    std::string path;
    int fd = -1;
    ImageKind kind = ImageKind::SPIDER;
    bool swap = false;
    ImageInfo info;
    // Spider: header bytes of the file and of each stacked image; MRC: data offset and mode
    size_t fileHeader = 0, imgHeader = 0, dataOffset = 0;
    int mrcMode = 2;
    bool isStack = false;
};

namespace {

using Kind = ImageKind;
using OpenFile = ImageSource;

// Open files are shared between the cache and the readers that are using them: evicting an entry only drops the
// cache's reference, so a loader thread in the middle of a pread keeps a valid descriptor (the CLI runs up to 16 loader
// threads, and a dataset stored as one file per particle cycles through far more than kMaxOpen files).
using FilePtr = std::shared_ptr<const OpenFile>;
struct CacheEntry { FilePtr file; uint64_t lastUse = 0; };
constexpr size_t kMaxOpen = 256;
std::mutex g_mutex;
std::map<std::string, CacheEntry> g_cache;
uint64_t g_tick = 0;

inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
inline float swapf(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u = bswap32(u);
    memcpy(&f, &u, 4);
    return f;
}

std::string lower(std::string s) {
    for (auto& c : s) c = (char)tolower(c);
    return s;
}

Kind kindOf(const std::string& fmt) {
    if (fmt == "mrc" || fmt == "mrcs" || fmt == "map" || fmt == "st" || fmt == "ali" || fmt == "rec") return Kind::MRC;
    return Kind::SPIDER;   // spi, xmp, stk, vol and anything unknown
}

void preadAll(int fd, void* buf, size_t n, off_t off, const std::string& what) {
    char* p = (char*)buf;
    while (n) {
        ssize_t r = pread(fd, p, n, off);
        if (r <= 0) throw std::runtime_error("short read in " + what);
        p += r;
        off += r;
        n -= (size_t)r;
    }
}

// IMAGIC stack: <base>.hed holds one 1024-byte record of 256 32-bit words per image (word 0: image number, 1: number of
// following images, 12: lines, 13: pixels per line, 14: type "REAL" / "INTG" / "PACK"), <base>.img the images back to
// back without headers.  Decoded into the MRC reader's terms (headerless data, mode 2 / 1 / 100), so reading an image is
// the same single pread (xmippCore: rwIMAGIC.cpp; fixtures resources/test/image/smallStack.hed, singleImage.hed).
void openImagic(OpenFile& f, const std::string& path) {
    const size_t dot = path.rfind('.');
    const std::string base = dot == std::string::npos ? path : path.substr(0, dot);
    const std::string hed = base + ".hed", img = base + ".img";
    int hfd = open(hed.c_str(), O_RDONLY);
    if (hfd < 0) throw std::runtime_error("cannot open IMAGIC header file " + hed);
    int32_t w[256];
    const bool got = pread(hfd, w, sizeof w, 0) == (ssize_t)sizeof w;
    close(hfd);
    if (!got) throw std::runtime_error(hed + ": not an IMAGIC header file");
    auto sane = [](int32_t v) { return v > 0 && v < 100000; };
    if (!(sane(w[12]) && sane(w[13]))) {
        f.swap = true;
        for (int i = 0; i < 256; ++i)
            if (i != 14) w[i] = (int32_t)bswap32((uint32_t)w[i]);      // word 14 is four characters
    }
    if (!(sane(w[12]) && sane(w[13])) || w[1] < 0) throw std::runtime_error(hed + ": bad IMAGIC header");
    char type[5] = {0, 0, 0, 0, 0};
    memcpy(type, &w[14], 4);
    const std::string t(type);
    if (t == "REAL") f.mrcMode = 2;
    else if (t == "INTG") f.mrcMode = 1;
    else if (t == "PACK") f.mrcMode = 100;        // unsigned bytes
    else throw std::runtime_error(hed + ": unsupported IMAGIC data type '" + t + "'");
    f.fd = open(img.c_str(), O_RDONLY);
    if (f.fd < 0) throw std::runtime_error("cannot open IMAGIC data file " + img);
    f.kind = Kind::MRC;
    f.dataOffset = 0;
    f.isStack = true;
    f.info.nx = w[13];
    f.info.ny = w[12];
    f.info.nz = 1;
    f.info.nImages = (size_t)w[1] + 1;
    struct stat st;
    const size_t bytes = f.mrcMode == 2 ? 4 : f.mrcMode == 1 ? 2 : 1;
    if (fstat(f.fd, &st) != 0 || (size_t)st.st_size < f.info.nImages * (size_t)f.info.nx * f.info.ny * bytes)
        throw std::runtime_error(img + ": shorter than its header file says");
}

FilePtr openFile(const std::string& path, const std::string& fmt) {
    auto fp = std::make_shared<OpenFile>();      // owns the descriptor from here on: every throw below closes it
    OpenFile& f = *fp;
    f.path = path;
    if (fmt == "hed" || fmt == "img") {
        openImagic(f, path);
        return fp;
    }
    f.fd = open(path.c_str(), O_RDONLY);
    if (f.fd < 0) throw std::runtime_error("cannot open image file " + path);
    struct stat st;
    if (fstat(f.fd, &st) != 0) throw std::runtime_error("cannot stat image file " + path);
    f.kind = kindOf(fmt);
    if (f.kind == Kind::MRC) {
        unsigned char h[1024];
        if (st.st_size < 1024) throw std::runtime_error(path + ": not an MRC file");
        preadAll(f.fd, h, 1024, 0, path);
        int32_t w[56];
        memcpy(w, h, sizeof w);
        auto sane = [](int32_t v) { return v > 0 && v < 100000; };
        if (!(sane(w[0]) && sane(w[1]) && sane(w[2]))) {
            f.swap = true;
            for (auto& v : w) v = (int32_t)bswap32((uint32_t)v);
        }
        if (!(sane(w[0]) && sane(w[1]) && sane(w[2]))) throw std::runtime_error(path + ": bad MRC header");
        f.mrcMode = w[3];
        int nx = w[0], ny = w[1], nz = w[2];
        int next = w[23];   // NSYMBT: bytes of extended header
        if (next < 0) next = 0;
        f.dataOffset = 1024 + (size_t)next;
        f.isStack = (fmt == "mrcs" || fmt == "st" || fmt == "ali");
        f.info.nx = nx;
        f.info.ny = ny;
        if (f.isStack) { f.info.nz = 1; f.info.nImages = nz; }
        else { f.info.nz = nz; f.info.nImages = 1; }
    } else {
        float h[64];
        if (st.st_size < 256) throw std::runtime_error(path + ": not a Spider file");
        preadAll(f.fd, h, sizeof h, 0, path);
        auto sane = [](float v) { return v >= 1.f && v < 1e5f && v == std::floor(v); };
        if (!(sane(h[1]) && sane(h[11]))) {
            f.swap = true;
            for (auto& v : h) v = swapf(v);
        }
        if (!(sane(h[1]) && sane(h[11]))) throw std::runtime_error(path + ": bad Spider header");
        int nslice = (int)std::fabs(h[0]), nrow = (int)h[1], nsam = (int)h[11];
        int labrec = (int)h[12], labbyt = (int)h[21], lenbyt = (int)h[22];
        if (lenbyt <= 0) lenbyt = nsam * 4;
        if (labrec <= 0) { labrec = 1024 / lenbyt; if (1024 % lenbyt) labrec++; }
        if (labbyt <= 0) labbyt = labrec * lenbyt;
        int istack = (int)h[23], maxim = (int)h[25];
        f.info.nx = nsam;
        f.info.ny = nrow;
        f.info.nz = std::max(1, nslice);
        f.fileHeader = (size_t)labbyt;
        if (istack > 0) {
            f.isStack = true;
            f.imgHeader = (size_t)labbyt;
            f.info.nImages = (size_t)std::max(1, maxim);
        } else {
            f.isStack = false;
            f.imgHeader = 0;
            f.info.nImages = 1;
        }
    }
    return fp;
}

FilePtr cached(const std::string& path, const std::string& fmt) {
    std::lock_guard<std::mutex> g(g_mutex);
    std::string key = path + ":" + fmt;
    auto it = g_cache.find(key);
    if (it != g_cache.end()) {
        it->second.lastUse = ++g_tick;
        return it->second.file;
    }
    FilePtr fp = openFile(path, fmt);
    if (g_cache.size() >= kMaxOpen) {            // evict the least recently used entry only
        auto lru = g_cache.begin();
        for (auto e = g_cache.begin(); e != g_cache.end(); ++e)
            if (e->second.lastUse < lru->second.lastUse) lru = e;
        g_cache.erase(lru);
    }
    g_cache[key] = CacheEntry{fp, ++g_tick};
    return fp;
}

void writeAll(FILE* f, const void* p, size_t n, const std::string& path) {
    if (fwrite(p, 1, n, f) != n) throw std::runtime_error("write failed: " + path);
}

void spiderHeader(std::vector<float>& h, int nx, int ny, int nz, int istack, int maxim, int imgnum, const float* data, size_t count) {
    int lenbyt = nx * 4;
    int labrec = 1024 / lenbyt;
    if (1024 % lenbyt) labrec++;
    int labbyt = labrec * lenbyt;
    h.assign((size_t)labbyt / 4, 0.f);
    h[0] = (float)nz;
    h[1] = (float)ny;
    h[2] = (float)(ny * nz + labrec);      // irec
    h[4] = (nz > 1) ? 3.f : 1.f;           // iform
    h[11] = (float)nx;
    h[12] = (float)labrec;
    h[21] = (float)labbyt;
    h[22] = (float)lenbyt;
    h[23] = (float)istack;
    h[25] = (float)maxim;
    h[26] = (float)imgnum;
    if (data && count) {
        double mn = data[0], mx = data[0], s = 0, s2 = 0;
        for (size_t i = 0; i < count; ++i) { double v = data[i]; mn = std::min(mn, v); mx = std::max(mx, v); s += v; s2 += v * v; }
        double av = s / count, var = std::max(0.0, s2 / count - av * av);
        h[5] = 1.f;                        // imami: statistics valid
        h[6] = (float)mx;
        h[7] = (float)mn;
        h[8] = (float)av;
        h[9] = (float)std::sqrt(var);
    }
    h[49] = 1.f;                           // scale
}

void mrcHeader(unsigned char* h, int nx, int ny, int nz, const float* data, size_t count) {
    memset(h, 0, 1024);
    int32_t w[56] = {0};
    w[0] = nx; w[1] = ny; w[2] = nz; w[3] = 2;
    w[7] = nx; w[8] = ny; w[9] = nz;
    float cell[3] = {(float)nx, (float)ny, (float)nz}, ang[3] = {90.f, 90.f, 90.f};
    memcpy(&w[10], cell, 12);
    memcpy(&w[13], ang, 12);
    w[16] = 1; w[17] = 2; w[18] = 3;
    float stats[3] = {0, 0, 0};
    if (data && count) {
        double mn = data[0], mx = data[0], s = 0;
        for (size_t i = 0; i < count; ++i) { mn = std::min<double>(mn, data[i]); mx = std::max<double>(mx, data[i]); s += data[i]; }
        stats[0] = (float)mn; stats[1] = (float)mx; stats[2] = (float)(s / count);
    }
    memcpy(&w[19], stats, 12);
    memcpy(h, w, sizeof w);
    memcpy(h + 208, "MAP ", 4);
    h[212] = 0x44; h[213] = 0x44;          // little-endian machine stamp
}

}  // namespace

void parseImageName(const std::string& spec, size_t& index, std::string& path, std::string& fmt) {
    index = 0;
    path = spec;
    size_t at = spec.find('@');
    if (at != std::string::npos) {
        index = (size_t)strtoull(spec.substr(0, at).c_str(), nullptr, 10);
        path = spec.substr(at + 1);
    }
    fmt.clear();
    size_t colon = path.rfind(':');
    size_t slash = path.rfind('/');
    if (colon != std::string::npos && (slash == std::string::npos || colon > slash)) {
        fmt = lower(path.substr(colon + 1));
        path = path.substr(0, colon);
    }
    if (fmt.empty()) {
        size_t dot = path.rfind('.');
        if (dot != std::string::npos && (slash == std::string::npos || dot > slash)) fmt = lower(path.substr(dot + 1));
    }
}

ImageInfo readImageInfo(const std::string& spec) {
    size_t idx;
    std::string path, fmt;
    parseImageName(spec, idx, path, fmt);
    return cached(path, fmt)->info;
}

void readImage2D(const std::string& spec, float* out, int nx, int ny) {
    size_t idx;
    std::string path, fmt;
    parseImageName(spec, idx, path, fmt);
    const FilePtr fp = cached(path, fmt);      // keeps the descriptor open for the duration of the read
    readImage2D(*fp, idx, out, nx, ny);
}

std::shared_ptr<const ImageSource> openImageSource(const std::string& path, const std::string& fmt) { return cached(path, fmt); }

void readImage2D(const ImageSource& f, size_t idx, float* out, int nx, int ny) {
    const std::string& spec = f.path;
    if (f.info.nx != nx || f.info.ny != ny)
        throw std::runtime_error("image " + spec + " is " + std::to_string(f.info.nx) + "x" + std::to_string(f.info.ny) +
                                 ", expected " + std::to_string(nx) + "x" + std::to_string(ny));
    size_t count = (size_t)nx * ny;
    size_t k = idx ? idx - 1 : 0;
    // a ".mrc" volume addressed with an index is treated as a stack of slices
    size_t avail = f.isStack ? f.info.nImages : (size_t)f.info.nz * f.info.nImages;
    if (k >= avail) throw std::runtime_error("image index " + std::to_string(idx) + " out of range in " + spec);
    if (f.kind == Kind::SPIDER) {
        off_t off = f.isStack ? (off_t)(f.fileHeader + k * (f.imgHeader + count * 4) + f.imgHeader) : (off_t)(f.fileHeader + k * count * 4);
        preadAll(f.fd, out, count * 4, off, spec);
        if (f.swap) for (size_t i = 0; i < count; ++i) out[i] = swapf(out[i]);
        return;
    }
    int bytes = (f.mrcMode == 0 || f.mrcMode == 100) ? 1 : (f.mrcMode == 1 || f.mrcMode == 6 || f.mrcMode == 12) ? 2 : 4;
    off_t off = (off_t)(f.dataOffset + k * count * bytes);
    if (f.mrcMode == 2) {
        preadAll(f.fd, out, count * 4, off, spec);
        if (f.swap) for (size_t i = 0; i < count; ++i) out[i] = swapf(out[i]);
        return;
    }
    std::vector<unsigned char> raw(count * bytes);
    preadAll(f.fd, raw.data(), raw.size(), off, spec);
    for (size_t i = 0; i < count; ++i) {
        if (f.mrcMode == 0) out[i] = (float)(int8_t)raw[i];
        else if (f.mrcMode == 100) out[i] = (float)raw[i];
        else if (f.mrcMode == 1 || f.mrcMode == 6) {
            uint16_t v;
            memcpy(&v, &raw[2 * i], 2);
            if (f.swap) v = (uint16_t)((v >> 8) | (v << 8));
            out[i] = (f.mrcMode == 1) ? (float)(int16_t)v : (float)v;
        } else if (f.mrcMode == 12) {
            uint16_t v;
            memcpy(&v, &raw[2 * i], 2);
            if (f.swap) v = (uint16_t)((v >> 8) | (v << 8));
            uint32_t sign = (v >> 15) & 1, ex = (v >> 10) & 31, man = v & 1023;
            float val;
            if (ex == 0) val = std::ldexp((float)man, -24);
            else if (ex == 31) val = man ? NAN : INFINITY;
            else val = std::ldexp((float)(man | 1024), (int)ex - 25);
            out[i] = sign ? -val : val;
        } else
            throw std::runtime_error("unsupported MRC mode in " + spec);
    }
}

void writeVolume(const std::string& spec, const float* data, int nx, int ny, int nz) {
    size_t idx;
    std::string path, fmt;
    parseImageName(spec, idx, path, fmt);
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    size_t count = (size_t)nx * ny * nz;
    try {
        if (kindOf(fmt) == Kind::MRC) {
            unsigned char h[1024];
            mrcHeader(h, nx, ny, nz, data, count);
            writeAll(f, h, 1024, path);
        } else {
            std::vector<float> h;
            spiderHeader(h, nx, ny, nz, 0, 0, 0, data, count);
            writeAll(f, h.data(), h.size() * 4, path);
        }
        writeAll(f, data, count * 4, path);
    } catch (...) {
        fclose(f);
        throw;
    }
    fclose(f);
}

void writeStack(const std::string& spec, const float* data, int nx, int ny, size_t n) {
    size_t idx;
    std::string path, fmt;
    parseImageName(spec, idx, path, fmt);
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    size_t count = (size_t)nx * ny;
    try {
        if (kindOf(fmt) == Kind::MRC) {
            unsigned char h[1024];
            mrcHeader(h, nx, ny, (int)n, data, count * n);
            writeAll(f, h, 1024, path);
            writeAll(f, data, count * n * 4, path);
        } else {
            std::vector<float> h;
            spiderHeader(h, nx, ny, 1, 2, (int)n, 0, nullptr, 0);
            writeAll(f, h.data(), h.size() * 4, path);
            for (size_t k = 0; k < n; ++k) {
                spiderHeader(h, nx, ny, 1, 0, 0, (int)k + 1, data + k * count, count);
                writeAll(f, h.data(), h.size() * 4, path);
                writeAll(f, data + k * count, count * 4, path);
            }
        }
    } catch (...) {
        fclose(f);
        throw;
    }
    fclose(f);
}

void closeImageCache() {
    std::lock_guard<std::mutex> g(g_mutex);
    g_cache.clear();       // descriptors close when their last holder lets go
}

}  // namespace rfhost
