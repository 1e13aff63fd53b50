// vtk_reader.cpp — see vtk_reader.hpp.
#include "vtk_reader.hpp"

#include <algorithm>
#include <cctype>
#include <charconv>
#include <cstring>
#include <stdexcept>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace c5host {

namespace {

struct mapped_file {
    const char* data = nullptr;
    std::size_t size = 0;
    int fd = -1;
    explicit mapped_file(const std::string& name) {
        fd = ::open(name.c_str(), O_RDONLY);
        if (fd < 0) throw std::runtime_error("cannot open " + name);
        struct stat st {};
        if (::fstat(fd, &st) != 0 || st.st_size == 0) {
            ::close(fd);
            throw std::runtime_error("cannot stat (or empty file) " + name);
        }
        size = static_cast<std::size_t>(st.st_size);
        void* p = ::mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED) {
            ::close(fd);
            throw std::runtime_error("cannot map " + name);
        }
        ::madvise(p, size, MADV_SEQUENTIAL);
        data = static_cast<const char*>(p);
    }
    ~mapped_file() {
        if (data) ::munmap(const_cast<char*>(data), size);
        if (fd >= 0) ::close(fd);
    }
    mapped_file(const mapped_file&) = delete;
    mapped_file& operator=(const mapped_file&) = delete;
};

class cursor {
public:
    cursor(const char* b, const char* e, std::string file) : _p(b), _e(e), _file(std::move(file)) {}

    [[noreturn]] void fail(const std::string& what) const {
        throw std::runtime_error(_file + ": " + what);
    }
    void skip_space() {
        while (_p < _e && std::isspace(static_cast<unsigned char>(*_p))) ++_p;
    }
    bool at_end() {
        skip_space();
        return _p >= _e;
    }
    bool has_more() const { return _p < _e; }
    std::string line() {
        const char* s = _p;
        while (_p < _e && *_p != '\n') ++_p;
        std::string out(s, _p);
        if (_p < _e) ++_p;
        while (!out.empty() && (out.back() == '\r' || out.back() == ' ')) out.pop_back();
        return out;
    }
    std::string word() {
        skip_space();
        const char* s = _p;
        while (_p < _e && !std::isspace(static_cast<unsigned char>(*_p))) ++_p;
        return std::string(s, _p);
    }
    std::string upper_word() {
        std::string w = word();
        std::transform(w.begin(), w.end(), w.begin(), [](unsigned char c) { return std::toupper(c); });
        return w;
    }
    long long integer() {
        skip_space();
        long long v = 0;
        auto r = std::from_chars(_p, _e, v);
        if (r.ec != std::errc()) fail("expected an integer");
        _p = r.ptr;
        return v;
    }
    double real() {
        skip_space();
        double v = 0;
        if (_p < _e && *_p == '+') ++_p;
        auto r = std::from_chars(_p, _e, v);
        if (r.ec != std::errc()) fail("expected a number");
        _p = r.ptr;
        return v;
    }
    // after the header line of a BINARY section: exactly one newline, then raw big-endian data
    void eat_newline() {
        while (_p < _e && (*_p == ' ' || *_p == '\r')) ++_p;
        if (_p < _e && *_p == '\n') ++_p;
    }
    const char* take(std::size_t bytes) {
        if (static_cast<std::size_t>(_e - _p) < bytes) fail("file ends inside a binary section");
        const char* s = _p;
        _p += bytes;
        return s;
    }
    // n elements of w bytes each (n * w must not wrap before it is compared with what is left)
    const char* take(std::size_t n, std::size_t w) {
        if (w != 0 && n > static_cast<std::size_t>(_e - _p) / w) fail("file ends inside a binary section");
        return take(n * w);
    }
    // a count read from the file: non-negative, and no larger than the bytes that are left could hold
    // (every value takes at least one byte in either format) — before anything is allocated for it
    void require_values(std::size_t n) const {
        if (n > static_cast<std::size_t>(_e - _p)) fail("a section declares more values than the file has bytes left");
    }
    std::size_t count(long long v, const char* what) {
        if (v < 0 || static_cast<unsigned long long>(v) > static_cast<unsigned long long>(_e - _p)) {
            fail(std::string("implausible ") + what + " (" + std::to_string(v) + ")");
        }
        return static_cast<std::size_t>(v);
    }

private:
    const char* _p;
    const char* _e;
    std::string _file;
};

std::size_t type_size(const std::string& t, const cursor& c) {
    if (t == "double" || t == "long" || t == "unsigned_long" || t == "vtktypeint64" || t == "vtktypeuint64" ||
        t == "vtkidtype")
        return 8;
    if (t == "float" || t == "int" || t == "unsigned_int" || t == "vtktypeint32" || t == "vtktypeuint32") return 4;
    if (t == "short" || t == "unsigned_short") return 2;
    if (t == "char" || t == "unsigned_char" || t == "bit") return 1;
    c.fail("unsupported data type '" + t + "'");
}

bool is_float_type(const std::string& t) {
    return t == "float" || t == "double";
}

template <class T>
T load_be(const char* p) {
    unsigned char b[sizeof(T)];
    for (std::size_t i = 0; i < sizeof(T); i++) b[i] = static_cast<unsigned char>(p[sizeof(T) - 1 - i]);
    T v;
    std::memcpy(&v, b, sizeof(T));
    return v;
}

// n values of VTK type `type` as doubles
void read_reals(cursor& c, bool binary, std::string type, std::size_t n, std::vector<double>& out) {
    std::transform(type.begin(), type.end(), type.begin(), [](unsigned char ch) { return std::tolower(ch); });
    c.require_values(n);
    out.resize(n);
    if (!binary) {
        for (auto& v : out) v = c.real();
        return;
    }
    const std::size_t w = type_size(type, c);
    c.eat_newline();
    const char* p = c.take(n, w);
    for (std::size_t i = 0; i < n; i++, p += w) {
        if (is_float_type(type)) {
            out[i] = (w == 8) ? load_be<double>(p) : static_cast<double>(load_be<float>(p));
        } else if (w == 8) {
            out[i] = static_cast<double>(load_be<int64_t>(p));
        } else if (w == 4) {
            out[i] = static_cast<double>(load_be<int32_t>(p));
        } else if (w == 2) {
            out[i] = static_cast<double>(load_be<int16_t>(p));
        } else {
            out[i] = static_cast<double>(static_cast<signed char>(*p));
        }
    }
}

void read_ints(cursor& c, bool binary, std::string type, std::size_t n, std::vector<long long>& out) {
    std::transform(type.begin(), type.end(), type.begin(), [](unsigned char ch) { return std::tolower(ch); });
    c.require_values(n);
    out.resize(n);
    if (!binary) {
        for (auto& v : out) v = c.integer();
        return;
    }
    const std::size_t w = type_size(type, c);
    c.eat_newline();
    const char* p = c.take(n, w);
    for (std::size_t i = 0; i < n; i++, p += w) {
        out[i] = (w == 8)   ? load_be<int64_t>(p)
                 : (w == 4) ? load_be<int32_t>(p)
                 : (w == 2) ? load_be<int16_t>(p)
                            : static_cast<long long>(static_cast<signed char>(*p));
    }
}

int32_t point_id(long long v, const cursor& c) { // before the narrowing cast: a 64-bit id must not alias a valid one
    if (v < 0 || v > 0x7FFFFFFFll) c.fail("cell references point id " + std::to_string(v) + " out of range");
    return static_cast<int32_t>(v);
}

} // namespace

tet_grid read_legacy_vtk(const std::string& filename) {
    mapped_file file(filename);
    cursor c(file.data, file.data + file.size, filename);
    tet_grid grid;

    const std::string magic = c.line();
    if (magic.rfind("# vtk DataFile", 0) != 0) c.fail("not a legacy VTK file (bad first line)");
    c.line(); // title
    const std::string format = c.upper_word();
    if (format != "ASCII" && format != "BINARY") c.fail("expected ASCII or BINARY, got '" + format + "'");
    const bool binary = format == "BINARY";

    std::size_t n_cells = 0;
    bool in_cell_data = false;
    std::vector<long long> ints, offsets;
    while (!c.at_end()) {
        const std::string key = c.upper_word();
        if (key.empty()) break;
        if (key == "DATASET") {
            const std::string kind = c.upper_word();
            if (kind != "UNSTRUCTURED_GRID") c.fail("DATASET " + kind + " is not UNSTRUCTURED_GRID");
        } else if (key == "POINTS") {
            const std::size_t n = c.count(c.integer(), "POINTS count");
            const std::string type = c.word();
            read_reals(c, binary, type, 3 * n, grid.points);
        } else if (key == "CELLS") {
            const long long a = c.integer();
            const long long b = c.integer();
            c.skip_space();
            // VTK >= 9 (file version 5.1): "CELLS n_offsets n_conn" then OFFSETS / CONNECTIVITY arrays
            std::string next;
            {
                cursor probe = c;
                next = probe.upper_word();
            }
            if (next == "OFFSETS") {
                c.upper_word();
                const std::string otype = c.word();
                read_ints(c, binary, otype, c.count(a, "OFFSETS count"), offsets);
                if (c.upper_word() != "CONNECTIVITY") c.fail("expected CONNECTIVITY after OFFSETS");
                const std::string ctype = c.word();
                read_ints(c, binary, ctype, c.count(b, "CONNECTIVITY count"), ints);
                n_cells = offsets.empty() ? 0 : offsets.size() - 1;
                // offsets index `ints`: non-negative, non-decreasing, the last one inside the array
                const long long n_conn = static_cast<long long>(ints.size());
                for (std::size_t k = 0; k < offsets.size(); k++) {
                    if (offsets[k] < 0 || offsets[k] > n_conn || (k > 0 && offsets[k] < offsets[k - 1])) {
                        c.fail("OFFSETS entry " + std::to_string(k) + " (" + std::to_string(offsets[k]) +
                               ") is negative, decreasing or beyond CONNECTIVITY");
                    }
                }
                grid.tets.resize(4 * n_cells);
                for (std::size_t k = 0; k < n_cells; k++) {
                    if (offsets[k + 1] - offsets[k] < 4) c.fail("cell " + std::to_string(k) + " has fewer than 4 points");
                    for (int i = 0; i < 4; i++) {
                        grid.tets[4 * k + i] = point_id(ints[static_cast<std::size_t>(offsets[k]) + i], c);
                    }
                }
            } else {
                n_cells = c.count(a, "CELLS count");
                read_ints(c, binary, "int", c.count(b, "CELLS size"), ints);
                grid.tets.resize(4 * n_cells);
                std::size_t at = 0;
                for (std::size_t k = 0; k < n_cells; k++) {
                    if (at >= ints.size()) c.fail("CELLS section is shorter than declared");
                    const long long m = ints[at];
                    if (m < 4 || at + 1 + static_cast<std::size_t>(m) > ints.size()) {
                        c.fail("cell " + std::to_string(k) + " has fewer than 4 points or is truncated");
                    }
                    for (int i = 0; i < 4; i++) grid.tets[4 * k + i] = point_id(ints[at + 1 + i], c);
                    at += 1 + static_cast<std::size_t>(m);
                }
            }
        } else if (key == "CELL_TYPES") {
            const std::size_t n = c.count(c.integer(), "CELL_TYPES count");
            read_ints(c, binary, "int", n, ints);
        } else if (key == "CELL_DATA") {
            c.integer();
            in_cell_data = true;
        } else if (key == "POINT_DATA") {
            c.integer();
            in_cell_data = false;
        } else if (key == "SCALARS") {
            const std::string name = c.word();
            const std::string type = c.word();
            long long comps = 1;
            {
                cursor probe = c;
                const std::string maybe = probe.upper_word();
                if (maybe != "LOOKUP_TABLE") comps = c.integer();
            }
            if (comps < 1 || comps > 4) c.fail("SCALARS " + name + ": 1 to 4 components");
            if (c.upper_word() != "LOOKUP_TABLE") c.fail("SCALARS " + name + ": expected LOOKUP_TABLE");
            c.word();
            std::vector<double> values;
            const std::size_t tuples = in_cell_data ? n_cells : grid.n_points();
            read_reals(c, binary, type, tuples * static_cast<std::size_t>(comps), values);
            if (in_cell_data) {
                if (comps != 1) { // first component, like *(GetTuple(k)) at object3d_base.cpp:48
                    std::vector<double> first(tuples);
                    for (std::size_t k = 0; k < tuples; k++) first[k] = values[k * static_cast<std::size_t>(comps)];
                    values.swap(first);
                }
                grid.cell_scalars[name] = std::move(values);
            }
        } else if (key == "FIELD") {
            c.word();
            const long long n_arrays = c.integer();
            for (long long a = 0; a < n_arrays; a++) {
                const std::string name = c.word();
                const long long comps = c.integer();
                const long long tuples = c.integer();
                const std::string type = c.word();
                std::vector<double> values;
                const std::size_t n_comp = c.count(comps, "FIELD component count"), n_tup = c.count(tuples, "FIELD tuple count");
                if (n_comp != 0 && n_tup > (static_cast<std::size_t>(-1) / 8) / n_comp) c.fail("implausible FIELD array size");
                read_reals(c, binary, type, n_comp * n_tup, values);
                if (in_cell_data && static_cast<std::size_t>(tuples) == n_cells && comps >= 1) {
                    std::vector<double> first(static_cast<std::size_t>(tuples));
                    for (std::size_t k = 0; k < first.size(); k++) first[k] = values[k * static_cast<std::size_t>(comps)];
                    grid.cell_scalars[name] = std::move(first);
                }
            }
        } else if (key == "LOOKUP_TABLE") {
            c.word();
            const std::size_t n = c.count(c.integer(), "LOOKUP_TABLE size");
            std::vector<double> skip;
            if (binary) {
                c.eat_newline();
                c.take(n, 4);
            } else {
                read_reals(c, false, "float", 4 * n, skip);
            }
        } else if (key == "VECTORS" || key == "NORMALS") {
            c.word();
            const std::string type = c.word();
            std::vector<double> skip;
            read_reals(c, binary, type, 3 * (in_cell_data ? n_cells : grid.n_points()), skip);
        } else if (key == "METADATA") {
            // "METADATA" block (VTK >= 8 writes one after arrays that carry information keys): the rest of
            // this line, then lines up to and including the first empty one. (at_end() would skip that
            // empty line as white space and the block would swallow the next section.)
            c.line();
            while (c.has_more()) {
                if (c.line().empty()) break;
            }
        } else {
            c.fail("unsupported section '" + key + "'");
        }
    }
    if (grid.points.empty() || grid.tets.empty()) c.fail("no POINTS or no CELLS");
    const long long n_pts = static_cast<long long>(grid.n_points());
    for (int32_t v : grid.tets) {
        if (v < 0 || v >= n_pts) c.fail("cell references point id " + std::to_string(v) + " out of range");
    }
    return grid;
}

} // namespace c5host
