// Host mirror of the reference's sortCSV helper (sort_csv.cpp:26-122): reorder the rows and columns of a
// (gzip or plain) Jaccard matrix to the order of a file of file names.  Pure host code, no GPU involved.
#include <getopt.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "comparator.h"
#include "seqio.h"

namespace spsp_host {

// utils.cpp:609-630 split(s, ','): the last field stops at the first non-printable byte.
static std::vector<std::string> split_fields(const std::string &s, char delim)
{
    std::vector<std::string> res;
    size_t pred = 0;
    for (size_t i = 0; i < s.size(); ++i)
        if (s[i] == delim) { res.push_back(s.substr(pred, i - pred)); pred = i + 1; }
    std::string last;
    for (size_t i = pred; i < s.size() && isprint((unsigned char)s[i]); ++i) last += s[i];
    res.push_back(last);
    return res;
}

static std::vector<std::string> lines_of(const std::vector<uint8_t> &raw, bool keep_trailing_empty)
{
    std::vector<std::string> out;
    size_t b = 0;
    while (b <= raw.size()) {
        size_t e = b;
        while (e < raw.size() && raw[e] != '\n') e++;
        if (e < raw.size() || b < raw.size() || keep_trailing_empty) out.emplace_back(raw.begin() + (ptrdiff_t)b, raw.begin() + (ptrdiff_t)e);
        b = e + 1;
    }
    return out;
}

int sort_csv(const std::string &filename, const std::string &outfilename, const std::string &fof_name)
{
    std::vector<uint8_t> raw, fraw;
    if (!read_file_maybe_gz(filename, raw)) {
        std::cout << "cant open file" << std::endl;
        return 0;
    }
    read_file_maybe_gz(fof_name, fraw);
    // `while (not fof.eof()) getline`: the empty string behind the last line end is an entry too (:35-38)
    std::vector<std::string> names_ordered = lines_of(fraw, true);
    std::vector<std::string> lines = lines_of(raw, false);
    if (lines.empty()) lines.emplace_back();
    std::vector<std::string> files_names = split_fields(lines[0], ',');
    const size_t N = files_names.size();
    std::vector<double> matrix(N * N, 0.0);
    std::map<uint32_t, uint32_t> sorted_names, old2new;
    std::map<uint32_t, std::string> names;
    for (uint32_t i = 0; i < N; ++i) {
        const uint32_t pos = (uint32_t)(std::find(names_ordered.begin(), names_ordered.end(), files_names[i]) - names_ordered.begin());
        sorted_names[pos] = i;
        names[pos] = files_names[i];
    }
    uint32_t id = 0;
    for (const auto &kv : sorted_names) old2new[kv.second] = id++;
    std::ofstream out(outfilename);
    id = 0;
    for (const auto &kv : names) {
        out << kv.second;
        id++;
        if (id != N) out << ',';
    }
    out << std::endl;
    uint32_t line_id = 0;
    for (size_t l = 1; l < lines.size(); l++) {
        if (lines[l].size() < N) break;                               // :80 (a containment file stops at its blank line)
        std::vector<std::string> values = split_fields(lines[l], ',');
        for (uint32_t i = 0; i < N && i < values.size(); ++i)
            matrix[(size_t)old2new[i] * N + old2new[line_id]] = std::stod(values[i]);
        line_id++;
    }
    for (size_t i = 0; i < N; ++i) {
        for (size_t j = 0; j < N; ++j) {
            out << matrix[i * N + j];
            if (matrix[i * N + j] != matrix[j * N + i]) std::cout << "bug1 OR you are sorting a containment file" << std::endl;
            if (i == j && matrix[i * N + j] != 1) {
                std::cout << matrix[i * N + j] << std::endl;
                std::cout << "bug2" << std::endl;                     // the reference waits for a key here (:100-104)
            }
            if (j != N - 1) out << ',';
        }
        out << std::endl;
    }
    std::cout << "The end" << std::endl;
    return 0;
}

int sort_csv_main(int argc, char **argv)
{
    if (argc < 4) {
        std::cout << "Need input, output filename and original fof" << std::endl;
        return 0;
    }
    try {
        return sort_csv(argv[1], argv[2], argv[3]);
    } catch (const std::exception &e) {
        std::cerr << "sortCSV: " << e.what() << std::endl;
        return 2;
    }
}

}  // namespace spsp_host
