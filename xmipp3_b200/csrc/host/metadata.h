// metadata.h — minimal Xmipp MetaData (.xmd / STAR) reader and writer.
//
// Stands in for xmippCore's MetaDataVec as used by the reconstruction path
// (SF.read / removeDisabled / containsLabel / getValue, reconstruct_fourier.cpp:191-197,
// 335-336, 362-381; CTF columns data/ctf.cpp:365-419, 1172-1212).  File layout as in
// src/xmipp/resources/test/sampling/experimental_images.xmd:
//     # XMIPP_STAR_1 *
//     data_noname
//     loop_
//      _image
//      _angleRot
//      ...
//      images/proj_sh000001.spi  2.5645  39.456 ...
// Non-loop blocks ("_label value" pairs, used by .ctfparam files) are read as one row.
#pragma once
#include <map>
#include <string>
#include <vector>

namespace rfhost {

class MetaData {
public:
    // "file" or "block@file"; throws std::runtime_error on I/O or syntax errors
    void read(const std::string& spec);
    void write(const std::string& path, const std::string& block = "noname") const;

    size_t size() const { return rows_.size(); }
    bool containsLabel(const std::string& label) const { return index_.count(label) != 0; }
    const std::vector<std::string>& labels() const { return labels_; }

    // value of `label` in row `i`; returns false (and leaves out untouched) if the column is absent
    bool getValue(const std::string& label, size_t i, std::string& out) const;
    bool getValue(const std::string& label, size_t i, double& out) const;
    double getValueOrDefault(const std::string& label, size_t i, double def) const;

    // column access for loops over many rows: resolve the label once (-1 if absent), then read cells by index
    int column(const std::string& label) const {
        auto it = index_.find(label);
        return it == index_.end() ? -1 : (int)it->second;
    }
    const std::string& cell(size_t i, int col) const { return rows_[i][(size_t)col]; }
    double cellOrDefault(size_t i, int col, double def) const;

    // drop rows whose `enabled` column is <= 0 (MetaData::removeDisabled)
    void removeDisabled();

    void addLabel(const std::string& label);
    size_t addRow();
    void setValue(const std::string& label, size_t i, const std::string& v);
    void setValue(const std::string& label, size_t i, double v);

private:
    std::vector<std::string> labels_;
    std::map<std::string, size_t> index_;
    std::vector<std::vector<std::string>> rows_;
};

}  // namespace rfhost
