// metadata.cpp — see metadata.h
#include "metadata.h"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace rfhost {

namespace {

// split a data line into tokens; single or double quotes protect blanks
std::vector<std::string> tokenize(const std::string& line) {
    std::vector<std::string> out;
    size_t i = 0, n = line.size();
    while (i < n) {
        while (i < n && (line[i] == ' ' || line[i] == '\t' || line[i] == '\r')) ++i;
        if (i >= n) break;
        if (line[i] == '\'' || line[i] == '"') {
            char q = line[i++];
            size_t j = i;
            while (j < n && line[j] != q) ++j;
            out.push_back(line.substr(i, j - i));
            i = (j < n) ? j + 1 : j;
        } else {
            size_t j = i;
            while (j < n && line[j] != ' ' && line[j] != '\t' && line[j] != '\r') ++j;
            out.push_back(line.substr(i, j - i));
            i = j;
        }
    }
    return out;
}

std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

}  // namespace

void MetaData::addLabel(const std::string& label) {
    if (index_.count(label)) return;
    index_[label] = labels_.size();
    labels_.push_back(label);
    for (auto& r : rows_) r.resize(labels_.size());
}

size_t MetaData::addRow() {
    rows_.emplace_back(labels_.size());
    return rows_.size() - 1;
}

void MetaData::setValue(const std::string& label, size_t i, const std::string& v) {
    addLabel(label);
    rows_.at(i).resize(labels_.size());
    rows_[i][index_[label]] = v;
}

void MetaData::setValue(const std::string& label, size_t i, double v) {
    char buf[64];
    snprintf(buf, sizeof buf, "%.10g", v);
    setValue(label, i, std::string(buf));
}

bool MetaData::getValue(const std::string& label, size_t i, std::string& out) const {
    auto it = index_.find(label);
    if (it == index_.end() || i >= rows_.size()) return false;
    out = rows_[i][it->second];
    return true;
}

bool MetaData::getValue(const std::string& label, size_t i, double& out) const {
    std::string s;
    if (!getValue(label, i, s) || s.empty()) return false;
    char* end = nullptr;
    double v = strtod(s.c_str(), &end);
    if (end == s.c_str()) return false;
    out = v;
    return true;
}

double MetaData::getValueOrDefault(const std::string& label, size_t i, double def) const {
    double v;
    return getValue(label, i, v) ? v : def;
}

double MetaData::cellOrDefault(size_t i, int col, double def) const {
    if (col < 0) return def;
    const std::string& s = rows_[i][(size_t)col];
    if (s.empty()) return def;
    char* end = nullptr;
    const double v = strtod(s.c_str(), &end);
    return end == s.c_str() ? def : v;
}

void MetaData::removeDisabled() {
    auto it = index_.find("enabled");
    if (it == index_.end()) return;
    std::vector<std::vector<std::string>> keep;
    for (auto& r : rows_) {
        double e = atof(r[it->second].c_str());
        if (e > 0) keep.push_back(std::move(r));
    }
    rows_.swap(keep);
}

void MetaData::read(const std::string& spec) {
    labels_.clear();
    index_.clear();
    rows_.clear();
    std::string block, path = spec;
    size_t at = spec.find('@');
    if (at != std::string::npos) {
        block = spec.substr(0, at);
        path = spec.substr(at + 1);
    }
    std::ifstream in(path);
    if (!in) throw std::runtime_error("MetaData: cannot open " + path);
    std::string line;
    bool inBlock = false, inLoop = false, labelsDone = false, anyBlock = false;
    std::vector<std::string> single;   // values of a non-loop block
    while (std::getline(in, line)) {
        std::string t = trim(line);
        if (t.empty() || t[0] == '#' || t[0] == ';') continue;
        if (t.compare(0, 5, "data_") == 0) {
            if (inBlock) break;   // next block begins: done
            std::string name = t.substr(5);
            anyBlock = true;
            if (block.empty() || name == block) inBlock = true;
            continue;
        }
        if (!anyBlock) {
            // headerless file (old selfile style "name flag"): not supported, be explicit
            throw std::runtime_error("MetaData: " + path + " has no data_ block (only STAR .xmd files are supported)");
        }
        if (!inBlock) continue;
        if (t == "loop_") {
            inLoop = true;
            continue;
        }
        if (t[0] == '_' && !labelsDone) {
            std::vector<std::string> tok = tokenize(t);
            addLabel(tok[0].substr(1));
            if (!inLoop) single.push_back(tok.size() > 1 ? tok[1] : std::string());
            continue;
        }
        if (!inLoop) continue;
        labelsDone = true;
        std::vector<std::string> tok = tokenize(t);
        if (tok.size() < labels_.size())
            throw std::runtime_error("MetaData: short row in " + path + ": '" + t + "'");
        tok.resize(labels_.size());
        rows_.push_back(std::move(tok));
    }
    if (!inBlock) throw std::runtime_error("MetaData: block '" + block + "' not found in " + path);
    if (!inLoop && !labels_.empty()) rows_.push_back(single);
}

void MetaData::write(const std::string& path, const std::string& block) const {
    std::ofstream out(path);
    if (!out) throw std::runtime_error("MetaData: cannot write " + path);
    out << "# XMIPP_STAR_1 * \n# \ndata_" << block << "\nloop_\n";
    for (auto& l : labels_) out << " _" << l << "\n";
    for (auto& r : rows_) {
        for (auto& v : r) {
            bool quote = v.empty() || v.find(' ') != std::string::npos;
            out << " " << (quote ? "'" + v + "'" : v);
        }
        out << "\n";
    }
}

}  // namespace rfhost
