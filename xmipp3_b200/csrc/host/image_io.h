// image_io.h — particle-stack readers and volume writers (Spider and MRC families).
//
// Stands in for xmippCore's Image<T>::read / readApplyGeo(file part) / write as used by the
// reconstruction path (reconstruct_fourier.cpp:199, 362, 1179).  Formats were decoded from the
// reference's fixtures src/xmipp/resources/test/image/{singleImage.spi,smallStack.stk,
// smallVolume.vol,singleImage.mrc,smallStack.mrcs}.
//
// File names follow Xmipp's conventions: "file", "NNNNNN@stack" (1-based image index) and an optional
// ":fmt" suffix that overrides the extension (e.g. "particles.mrc:mrcs").
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>

namespace rfhost {

struct ImageInfo {
    int nx = 0, ny = 0, nz = 0;     // nz: slices of ONE object (1 for 2-D images)
    size_t nImages = 0;             // objects in the file (stack length)
};

// parse "NNNNNN@path:fmt" -> index (0 = none), path, format ("spi", "stk", "vol", "mrc", "mrcs", ...)
void parseImageName(const std::string& spec, size_t& index, std::string& path, std::string& fmt);

ImageInfo readImageInfo(const std::string& spec);

// Read the 2-D image designated by `spec` as float32 (nx*ny values, row-major, y outer).
// Throws std::runtime_error if the file is unreadable or the size differs from (nx, ny).
void readImage2D(const std::string& spec, float* out, int nx, int ny);

// The same in two steps, for loaders that read many images of few files: an ImageSource is an open file with its
// header decoded (shared with the descriptor cache; the descriptor stays valid for as long as the pointer is held), and
// reading from it involves no name parsing, no lock and no copy: one pread per image straight into `out`.
struct ImageSource;
std::shared_ptr<const ImageSource> openImageSource(const std::string& path, const std::string& fmt);
void readImage2D(const ImageSource& src, size_t index /* 1-based, 0 = the only image */, float* out, int nx, int ny);

// Write helpers.  The format is chosen from the extension / ":fmt" suffix:
//   .vol .spi .xmp .stk -> Spider,  .mrc .mrcs .map -> MRC (mode 2)
void writeVolume(const std::string& spec, const float* data, int nx, int ny, int nz);
void writeStack(const std::string& spec, const float* data, int nx, int ny, size_t n);

// close cached file descriptors (the readers keep recently used stacks open)
void closeImageCache();

}  // namespace rfhost
