// symmetries.cpp — see symmetries.h
#include "symmetries.h"

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace rfhost {

namespace {

constexpr double kPi = 3.14159265358979323846;

Mat3 identity() { return {1, 0, 0, 0, 1, 0, 0, 0, 1}; }

Mat3 mul(const Mat3& a, const Mat3& b) {
    Mat3 c{};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += a[i * 3 + k] * b[k * 3 + j];
            c[i * 3 + j] = s;
        }
    return c;
}

bool close(const Mat3& a, const Mat3& b) {
    for (int i = 0; i < 9; ++i)
        if (std::fabs(a[i] - b[i]) > 1e-6) return false;
    return true;
}

// rotation by `deg` about `axis` (Rodrigues)
Mat3 axisRotation(double x, double y, double z, double deg) {
    double n = std::sqrt(x * x + y * y + z * z);
    if (n == 0) throw std::runtime_error("symmetry: null rotation axis");
    x /= n; y /= n; z /= n;
    double t = deg * kPi / 180.0, c = std::cos(t), s = std::sin(t), C = 1 - c;
    return {c + x * x * C, x * y * C - z * s, x * z * C + y * s,
            y * x * C + z * s, c + y * y * C, y * z * C - x * s,
            z * x * C - y * s, z * y * C + x * s, c + z * z * C};
}

Mat3 mirror(double x, double y, double z) {
    double n = std::sqrt(x * x + y * y + z * z);
    if (n == 0) throw std::runtime_error("symmetry: null mirror normal");
    x /= n; y /= n; z /= n;
    return {1 - 2 * x * x, -2 * x * y, -2 * x * z, -2 * y * x, 1 - 2 * y * y, -2 * y * z, -2 * z * x, -2 * z * y, 1 - 2 * z * z};
}

Mat3 inversion() { return {-1, 0, 0, 0, -1, 0, 0, 0, -1}; }

std::vector<Mat3> closeGroup(const std::vector<Mat3>& gens) {
    std::vector<Mat3> elems{identity()}, frontier{identity()};
    while (!frontier.empty()) {
        std::vector<Mat3> next;
        for (auto& e : frontier)
            for (auto& g : gens) {
                Mat3 c = mul(g, e);
                bool found = false;
                for (auto& x : elems)
                    if (close(c, x)) { found = true; break; }
                if (!found) {
                    elems.push_back(c);
                    next.push_back(c);
                    if (elems.size() > 480) throw std::runtime_error("symmetry: generators do not close to a finite point group");
                }
            }
        frontier.swap(next);
    }
    elems.erase(elems.begin());   // drop the identity
    return elems;
}

bool generatorsFromName(std::string s, std::vector<Mat3>& g) {
    for (auto& c : s) c = (char)tolower(c);
    auto rot = axisRotation;
    if (s.empty()) return false;
    if (s[0] == 'i') {
        bool h = s.size() > 1 && s.back() == 'h';
        std::string base = h ? s.substr(0, s.size() - 1) : s;
        if (base == "i" || base == "i2")
            g = {rot(0, 0, 1, 180), rot(0.525731114, 0, 0.850650807, 72), rot(0, 0.356822076, 0.934172364, 120)};
        else if (base == "i1")
            g = {rot(1, 0, 0, 180), rot(0.85065080702670, 0, -0.5257311142635, 72), rot(0.9341723640, 0.3568220765, 0, 120)};
        else if (base == "i3")
            g = {rot(-0.5257311143, 0, 0.8506508070, 180), rot(0, 0, 1, 72), rot(-0.4911234778630044, 0.3568220764705179, 0.7946544753759428, 120)};
        else if (base == "i4")
            g = {rot(0.5257311143, 0, 0.8506508070, 180), rot(0.8944271932547096, 0, 0.4472135909903704, 72),
                 rot(0.4911234778630044, 0.3568220764705179, 0.7946544753759428, 120)};
        else
            return false;
        if (h) g.push_back(inversion());
        return true;
    }
    if (s == "t" || s == "td" || s == "th") {
        g = {rot(0, 0, 1, 120), rot(0, 0.816496, 0.577350, 180)};
        if (s == "td") g.push_back(mirror(1.4142136, 2.4494897, 0.0));
        if (s == "th") g.push_back(inversion());
        return true;
    }
    if (s == "o" || s == "oh") {
        g = {rot(.5773502, .5773502, .5773502, 120), rot(0, 0, 1, 90)};
        if (s == "oh") g.push_back(mirror(0, 1, 1));
        return true;
    }
    char kind = s[0];
    if (kind != 'c' && kind != 'd' && kind != 's') return false;
    size_t i = 1;
    while (i < s.size() && isdigit((unsigned char)s[i])) ++i;
    if (i == 1) return false;
    int n = atoi(s.substr(1, i - 1).c_str());
    std::string tail = s.substr(i);
    if (n < 1) return false;
    g.clear();
    if (kind == 's') {
        if (n % 2 || !tail.empty()) return false;
        if (n / 2 > 1) g.push_back(rot(0, 0, 1, 360.0 / (n / 2)));
        g.push_back(inversion());
        return true;
    }
    if (n > 1) g.push_back(rot(0, 0, 1, 360.0 / n));
    if (kind == 'd') g.push_back(rot(1, 0, 0, 180));
    if (tail == "v") g.push_back(kind == 'c' ? mirror(0, 1, 0) : mirror(1, 0, 0));
    else if (tail == "h") g.push_back(mirror(0, 0, 1));
    else if (!tail.empty()) return false;
    return true;
}

bool generatorsFromFile(const std::string& path, std::vector<Mat3>& g) {
    std::ifstream in(path);
    if (!in) return false;
    std::string line;
    g.clear();
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::string key;
        if (!(ss >> key) || key[0] == '#' || key[0] == ';') continue;
        if (key == "rot_axis") {
            int fold;
            double x, y, z;
            if (!(ss >> fold >> x >> y >> z) || fold < 1) throw std::runtime_error("symmetry file " + path + ": bad rot_axis line");
            if (fold > 1) g.push_back(axisRotation(x, y, z, 360.0 / fold));
        } else if (key == "mirror_plane") {
            double x, y, z;
            if (!(ss >> x >> y >> z)) throw std::runtime_error("symmetry file " + path + ": bad mirror_plane line");
            g.push_back(mirror(x, y, z));
        } else if (key == "inversion") {
            g.push_back(inversion());
        } else
            throw std::runtime_error("symmetry file " + path + ": unknown keyword '" + key + "'");
    }
    return true;
}

}  // namespace

std::vector<Mat3> symmetryMatrices(const std::string& nameOrFile) {
    std::vector<Mat3> gens;
    if (!generatorsFromName(nameOrFile, gens) && !generatorsFromFile(nameOrFile, gens))
        throw std::runtime_error("unknown symmetry group or unreadable symmetry file: " + nameOrFile);
    if (gens.empty()) return {};
    return closeGroup(gens);
}

}  // namespace rfhost
