// libsvm_text.hpp -- readers for libsvm's sparse text rows ("label idx:val idx:val ..."), shared by the two CLI front
// ends.  Two dialects, because the two reference programs tokenise differently:
//   * svm-predict (svm-predict.c:72-128): strtok on " \t\n" / ":" / " \t", strict -- ascending indices, every number must
//     be consumed completely, errno from strtol / strtod must be 0; the first bad line stops the program.
//   * svm-scale (svm-scale.c:131-146, :175-190, :272-285): sscanf("%d:%lf") pairs, lenient -- a row ends at the first
//     thing that is not such a pair.
// Host-side text handling only; all arithmetic happens behind the C ABI.
#pragma once
#include <cctype>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace hafsvmtext {

struct Rows {
    std::vector<double> target;
    std::vector<long long> row_ptr{0};
    std::vector<int> index;
    std::vector<double> value;
    int max_index = 0;
    int n() const { return (int)target.size(); }
};

inline bool read_line(FILE* fp, std::string& line) {
    line.clear();
    char buf[65536];
    while (fgets(buf, sizeof buf, fp)) {
        line += buf;
        if (!line.empty() && line.back() == '\n') return true;
    }
    return !line.empty();
}

inline bool in_set(char c, const char* set) { return c != '\0' && strchr(set, c) != nullptr; }

// One row in svm-predict's dialect.  Returns false on what svm-predict calls "Wrong input format".
inline bool parse_predict_row(std::string& line, Rows& R) {
    char* p = &line[0];
    while (in_set(*p, " \t\n")) p++;
    if (!*p) return false;  // empty line
    char* tok = p;
    while (*p && !in_set(*p, " \t\n")) p++;
    if (*p) *p++ = '\0';
    char* end = nullptr;
    const double target = strtod(tok, &end);
    if (end == tok || *end != '\0') return false;
    const size_t mark_i = R.index.size();
    int prev = -1;
    for (;;) {
        while (*p == ':') p++;           // index token: up to the next ':' (leading blanks stay in the token)
        if (!*p) break;
        char* idx = p;
        while (*p && *p != ':') p++;
        if (!*p) break;                  // no ':' left -> there is no value token
        *p++ = '\0';
        while (in_set(*p, " \t")) p++;   // value token: up to the next blank
        if (!*p) break;
        char* val = p;
        while (*p && !in_set(*p, " \t")) p++;
        if (*p) *p++ = '\0';
        errno = 0;
        const long li = strtol(idx, &end, 10);
        const int i = (int)li;
        if (end == idx || errno != 0 || *end != '\0' || i <= prev) { R.index.resize(mark_i); R.value.resize(mark_i); return false; }
        prev = i;
        errno = 0;
        const double v = strtod(val, &end);
        if (end == val || errno != 0 || (*end != '\0' && !isspace((unsigned char)*end))) { R.index.resize(mark_i); R.value.resize(mark_i); return false; }
        R.index.push_back(i);
        R.value.push_back(v);
        if (i > R.max_index) R.max_index = i;
    }
    R.target.push_back(target);
    R.row_ptr.push_back((long long)R.index.size());
    return true;
}

// One row in svm-scale's dialect: "%lf" target, then "%d:%lf" pairs until one does not match.
inline void parse_scale_row(const std::string& line, Rows& R, long long* n_pairs) {
    const char* p = line.c_str();
    char* end = nullptr;
    double target = strtod(p, &end);   // sscanf(p, "%lf", &target); a failed conversion leaves the reference's variable
    if (end == p) target = 0.0;        // uninitialised -- defined here as 0
    while (isspace((unsigned char)*p)) p++;
    while (*p && !isspace((unsigned char)*p)) p++;
    for (;;) {
        const char* q = p;
        while (isspace((unsigned char)*q)) q++;
        errno = 0;
        const long li = strtol(q, &end, 10);
        if (end == q || *end != ':') break;
        const char* vq = end + 1;
        while (isspace((unsigned char)*vq)) vq++;
        char* vend = nullptr;
        const double v = strtod(vq, &vend);
        if (vend == vq) break;
        R.index.push_back((int)li);
        R.value.push_back(v);
        if ((int)li > R.max_index) R.max_index = (int)li;
        if (n_pairs) (*n_pairs)++;
        p = vend;                        // SKIP_ELEMENT: past ':' and the value's characters
        while (*p && !isspace((unsigned char)*p)) p++;
    }
    R.target.push_back(target);
    R.row_ptr.push_back((long long)R.index.size());
}

}  // namespace hafsvmtext
